// fp64_peak.cu — measured instruction-rate roof of the pruning kernel's arithmetic on this GPU.
//
// k_prune's arithmetic is DMUL and DADD rounded separately (the reference's evaluation order forbids FMA contraction:
// DESIGN.md §3), so its roof is not the data-sheet FMA rate but the rate at which the fp64 pipe retires DMUL/DADD
// instructions.  This microbenchmark issues exactly those — 8 independent chains per thread, alternating
// __dmul_rn / __dadd_rn, no memory traffic — over every SM at full occupancy and reports operations per second
// (one DMUL or DADD = one operation).  bench.py calls it (ctypes) and uses the result as `roofline.peak`.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o tools/libfp64peak.so tools/fp64_peak.cu
#include <cuda_runtime.h>

namespace {
constexpr int CHAINS = 8;
__global__ void __launch_bounds__(256) k_fp64(double* out, double a, double b, int iters) {
  double x[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) x[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += x[i];
  if (s == 12345.678) out[0] = s;               // never true: keeps the chains alive
}
}  // namespace

// operations (DMUL + DADD) per second on `device`, best of `reps` launches; < 0 on a CUDA error
extern "C" double fp64_issue_peak(int device, int reps) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
  double* d = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8, iters = 4096;
  double best = 0.0;
  for (int r = 0; r < reps + 1; r++) {
    cudaEventRecord(e0);
    k_fp64<<<blocks, 256>>>(d, 0.999999, 1e-7, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 2.0 * CHAINS * (double)iters * 256.0 * blocks / (ms * 1e-3);
    if (r > 0 && ops > best) best = ops;        // the first launch warms up
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}
