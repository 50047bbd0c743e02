// kernels.cuh — sm_100a CUDA kernels of the Felsenstein-pruning likelihood engine.
//
// Data layout in HBM (per handle; SURVEY.md §8d "site-major structure-of-arrays"):
//   tipmask  u8     [N][Sp]                 4-bit state masks, Sp = numLoci padded to a multiple of 128
//   part[s]  f64x4  [N-1][R][Sp]            conditional-likelihood 4-vectors, one 32-byte element per
//                                           (internal vertex, rate, site); two slots s = 0,1 so that a
//                                           rejected proposal rolls back by flipping a per-vertex slot bit
//   expo[s]  f32    [N-1][R][Sp]            per-vertex rescale exponents (IntermediateNode.cpp:63: a FLOAT)
//   flag[s]  u8     [R*T][N-1]              1 when any site of the 128-site tile rescaled at that vertex;
//                                           expo tiles whose flag is 0 are never written nor read
// One thread owns one (site, rate) of one vertex: its two child loads are 256-bit (LDG.E.256), consecutive
// threads touch consecutive 32-byte elements, so every warp request is 1 KB fully coalesced.  Per update the
// compulsory traffic is 2 x 32 B read (internal children; 1 B for a tip) + 32 B written: 68-96 B, against
// ~70 fp64 operations — HBM-bound, no tensor cores (a 4x4 mat-vec has nothing to tile).
//
// Arithmetic follows the reference's evaluation order with separately rounded products and sums
// (__dmul_rn/__dadd_rn: never contracted into FMAs), so partials are bit-identical to the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmpc {

constexpr int RMAX = 8;         // rate categories supported (the packaged data uses 3)
constexpr int TILE = 128;       // sites per block of the pruning kernel = flag granularity
constexpr int CHUNK = 256;      // sites per deterministic reduction chunk

struct Task {                   // one dirty vertex of one level
  int32_t out;                  // internal index (vertex id - N) to write
  int32_t c0, c1;               // children: tip index, or internal index
  int32_t flags;
};
enum : int32_t { TF_OUTSLOT = 1, TF_C0TIP = 2, TF_C1TIP = 4, TF_C0SLOT = 8, TF_C1SLOT = 16, TF_WITHIN = 32 };

struct DevView {
  double* part[2];
  float* expo[2];
  uint8_t* flag[2];
  const uint8_t* tipmask;
  uint8_t* cur;                 // [N-1] slot of each internal vertex, as seen by the root reduction
  const double* lut;            // [2][R][16][4]: P * indicator(mask) for every 4-bit mask (tips)
  int N, S, Sp, R, T, nInt;
  // select instead of dynamic indexing: indexing a by-value kernel parameter array would spill it to local memory
  __device__ __forceinline__ double* partOf(int slot) const { return slot ? part[1] : part[0]; }
  __device__ __forceinline__ float* expoOf(int slot) const { return slot ? expo[1] : expo[0]; }
  __device__ __forceinline__ uint8_t* flagOf(int slot) const { return slot ? flag[1] : flag[0]; }
};

struct RootView {
  float* elist;                 // [Sp][Kcap][RP] per-site non-zero exponents in vertex-id order, zero padded
  int* kmax;                    // [Sp] rows used at each site (uncapped)
  double* rootdot;              // [R][Sp] dot(root partial, limProbs)
  int* act;                     // [Sp] active (rescaled) sites, ascending
  int* rowoff;                  // [Sp+1] exclusive prefix of rows over active sites
  int* posIncl;                 // [Sp] number of active sites <= s
  int* chunkStart;              // chain staging chunks
  float* cout;                  // [Sp][RP] normalised exponent vector after each active site
  float* mxo;                   // [Sp] maxExponent of each active site
  float* carryIn;               // [RMAX] exponent vector entering this shard (zeros on a single GPU)
  float* carryOut;              // [RMAX]
  double* sitell;               // [Sp] rateAveragedLogLiks
  double* chunk;                // [nChunks] sums of CHUNK consecutive sites
  unsigned* ticket;
  int* overflow;
  double* result;
  int Kcap, RP, Q;              // Q = rows per chain staging chunk
};

__constant__ double c_P[2][RMAX][16];   // [set][rate][i*4+j], set 0 = between, 1 = within; P(i,j) row-major
__constant__ double c_pi[4];

__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// y = P * x in Armadillo's gemv_emul_tinysq order: ((P(i,0)x0 + P(i,1)x1) + P(i,2)x2) + P(i,3)x3
__device__ __forceinline__ double row4(const double* Pi, double x0, double x1, double x2, double x3) {
  return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(Pi[0], x0), __dmul_rn(Pi[1], x1)), __dmul_rn(Pi[2], x2)),
                   __dmul_rn(Pi[3], x3));
}

// IntermediateNode.cpp:58-64: if max < 1e-150, divide by it and OVERWRITE the exponent with (float)log(max)
__device__ __forceinline__ void rescale(double& s0, double& s1, double& s2, double& s3, float& e) {
  double m = fmax(fmax(s0, s1), fmax(s2, s3));
  if (m < 1e-150) {
    s0 = s0 / m; s1 = s1 / m; s2 = s2 / m; s3 = s3 / m;
    e = (float)log(m);
  }
}

// ------------------------------------------------------------------------------------------------
// K2 prune_level: all dirty vertices of one level x all (rate, site) pairs.
// Replaces IntermediateNode::ComputeSolutions/ComputeSolution (IntermediateNode.cpp:26-71) driven by
// AugTree::TrySolve (AugTree.cpp:114-134).  grid = tasks * R * T blocks of 128 threads.
__global__ void __launch_bounds__(TILE) k_prune_level(DevView v, const Task* __restrict__ tasks) {
  __shared__ double s_lut[64];
  const int RT = v.R * v.T;
  const int task = blockIdx.x / RT;
  const int blk = blockIdx.x - task * RT;
  const int r = blk / v.T;
  const int tile = blk - r * v.T;
  const Task t = tasks[task];
  const int set = (t.flags & TF_WITHIN) ? 1 : 0;
  if (t.flags & (TF_C0TIP | TF_C1TIP)) {
    if (threadIdx.x < 64) s_lut[threadIdx.x] = v.lut[(size_t)(set * v.R + r) * 64 + threadIdx.x];
    __syncthreads();
  }
  const int site = tile * TILE + threadIdx.x;
  const double* P = c_P[set][r];

  // issue both children's loads before any arithmetic
  double a0, a1, a2, a3, b0, b1, b2, b3;
  unsigned ma = 0, mb = 0;
  if (t.flags & TF_C0TIP) ma = v.tipmask[(size_t)t.c0 * v.Sp + site];
  else ld256(v.partOf(t.flags & TF_C0SLOT) + (((size_t)t.c0 * v.R + r) * v.Sp + site) * 4, a0, a1, a2, a3);
  if (t.flags & TF_C1TIP) mb = v.tipmask[(size_t)t.c1 * v.Sp + site];
  else ld256(v.partOf(t.flags & TF_C1SLOT) + (((size_t)t.c1 * v.R + r) * v.Sp + site) * 4, b0, b1, b2, b3);

  double s0, s1, s2, s3;
  float e = 0.0f;
  if (t.flags & TF_C0TIP) { const double* y = s_lut + 4 * ma; s0 = y[0]; s1 = y[1]; s2 = y[2]; s3 = y[3]; }
  else { s0 = row4(P, a0, a1, a2, a3); s1 = row4(P + 4, a0, a1, a2, a3); s2 = row4(P + 8, a0, a1, a2, a3); s3 = row4(P + 12, a0, a1, a2, a3); }
  rescale(s0, s1, s2, s3, e);                                   // after child 0 (ones % y == y exactly)
  if (t.flags & TF_C1TIP) {
    const double* y = s_lut + 4 * mb;
    s0 = __dmul_rn(s0, y[0]); s1 = __dmul_rn(s1, y[1]); s2 = __dmul_rn(s2, y[2]); s3 = __dmul_rn(s3, y[3]);
  } else {
    s0 = __dmul_rn(s0, row4(P, b0, b1, b2, b3)); s1 = __dmul_rn(s1, row4(P + 4, b0, b1, b2, b3));
    s2 = __dmul_rn(s2, row4(P + 8, b0, b1, b2, b3)); s3 = __dmul_rn(s3, row4(P + 12, b0, b1, b2, b3));
  }
  rescale(s0, s1, s2, s3, e);                                   // after child 1: overwrites

  const int os = (t.flags & TF_OUTSLOT) ? 1 : 0;
  const size_t o = ((size_t)t.out * v.R + r) * v.Sp + site;
  st256(v.partOf(os) + o * 4, s0, s1, s2, s3);
  const int any = __syncthreads_or(e != 0.0f);
  if (any) v.expoOf(os)[o] = e;
  if (threadIdx.x == 0) {
    v.flagOf(os)[(size_t)blk * v.nInt + t.out] = (uint8_t)(any != 0);
    if (blk == 0) v.cur[t.out] = (uint8_t)os;
  }
}

// rollback: flip the slot bits of the vertices a rejected proposal touched (IntermediateNode.h:26-33)
__global__ void k_set_cur(uint8_t* cur, const int32_t* idxAndSlot, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cur[idxAndSlot[i] >> 1] = (uint8_t)(idxAndSlot[i] & 1);
}

// pad sites get the all-ones mask; counts cells whose mask is 0 (unknown symbol)
__global__ void k_prepare_masks(uint8_t* mask, int N, int S, int Sp, int* bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * Sp) return;
  int s = (int)(i % Sp);
  if (s >= S) mask[i] = 0xF;
  else if ((mask[i] & 0xF) == 0 || mask[i] > 0xF) atomicAdd(bad, 1);
}

// ------------------------------------------------------------------------------------------------
// K4a root_gather: per site tile, dot(root, limProbs) and the per-site list of non-zero rescale exponents in
// vertex-id order (the order AugTree.cpp:309-315 adds them in).  Only vertices whose tile flag is set are read.
__global__ void __launch_bounds__(TILE) k_root_gather(DevView v, RootView rv) {
  __shared__ unsigned s_ballot[TILE / 32];
  __shared__ int s_cnt[RMAX][TILE];
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int site = tile * TILE + tid;
  const bool valid = site < v.S;
  const int rootSlot = v.cur[0];
  for (int r = 0; r < v.R; r++) {
    double a0, a1, a2, a3;
    ld256(v.partOf(rootSlot) + ((size_t)r * v.Sp + site) * 4, a0, a1, a2, a3);
    // arma::dot (op_dot::direct_dot_arma): two interleaved accumulators
    rv.rootdot[(size_t)r * v.Sp + site] =
        __dadd_rn(__dadd_rn(__dmul_rn(a0, c_pi[0]), __dmul_rn(a2, c_pi[2])), __dadd_rn(__dmul_rn(a1, c_pi[1]), __dmul_rn(a3, c_pi[3])));
  }
  int kmax = 0;
  for (int r = 0; r < v.R; r++) {
    int k = 0;
    const uint8_t* f0 = v.flag[0] ? v.flag[0] + (size_t)(r * v.T + tile) * v.nInt : nullptr;
    const uint8_t* f1 = v.flag[1] ? v.flag[1] + (size_t)(r * v.T + tile) * v.nInt : nullptr;
    for (int base = 0; base < v.nInt; base += TILE) {
      const int nd = base + tid;
      bool f = false;
      if (nd < v.nInt) f = (v.cur[nd] ? f1[nd] : f0[nd]) != 0;
      const unsigned b = __ballot_sync(0xffffffffu, f);
      if ((tid & 31) == 0) s_ballot[tid >> 5] = b;
      __syncthreads();
      for (int w = 0; w < TILE / 32; w++) {
        unsigned bits = s_ballot[w];
        while (bits) {
          const int node = base + 32 * w + __ffs(bits) - 1;
          bits &= bits - 1;
          const float e = v.expoOf(v.cur[node])[((size_t)node * v.R + r) * v.Sp + site];
          if (e != 0.0f) {
            if (k < rv.Kcap && valid) rv.elist[((size_t)site * rv.Kcap + k) * rv.RP + r] = e;
            k++;
          }
        }
      }
      __syncthreads();
    }
    s_cnt[r][tid] = k;
    kmax = max(kmax, k);
  }
  if (valid) {
    const int kk = min(kmax, rv.Kcap);
    for (int r = 0; r < rv.RP; r++) {
      const int kr = r < v.R ? min(s_cnt[r][tid], rv.Kcap) : 0;
      for (int j = kr; j < kk; j++) rv.elist[((size_t)site * rv.Kcap + j) * rv.RP + r] = 0.0f;
    }
    rv.kmax[site] = kmax;
    if (kmax > rv.Kcap) atomicMax(rv.overflow, kmax);
  } else {
    rv.kmax[site] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// K4b chain: AugTree.cpp:306-318 keeps ONE fp32 exponent vector across all loci (it is never reset), so the
// per-site values depend on the site order through a sequential chain of fp32 additions whose rounding has to
// be reproduced bit for bit.  One block: every thread compacts the active sites, then
// thread 0 walks them in order while the other warps stage the next chunk of exponent rows in shared memory.
template <int RP>
__global__ void __launch_bounds__(1024) k_chain(DevView v, RootView rv) {
  extern __shared__ float4 s_dyn[];
  constexpr int V = RP / 4;                       // float4 per row
  __shared__ int s_scan[2][32];
  __shared__ int s_total[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nth = blockDim.x;
  const int S = v.S, Q = rv.Q;

  // phase A: ordered compaction of the sites that rescaled anywhere (kmax > 0) + row prefix
  int baseCnt = 0, baseRows = 0;
  for (int start = 0; start < S; start += nth) {
    const int s = start + tid;
    const int km = s < S ? min(rv.kmax[s], rv.Kcap) : 0;
    const int actv = km > 0;
    int ic = actv, ir = km;                        // inclusive warp scans
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int tc = __shfl_up_sync(0xffffffffu, ic, d), tr = __shfl_up_sync(0xffffffffu, ir, d);
      if (lane >= d) { ic += tc; ir += tr; }
    }
    if (lane == 31) { s_scan[0][wid] = ic; s_scan[1][wid] = ir; }
    __syncthreads();
    if (wid == 0) {
      int wc = lane < (nth >> 5) ? s_scan[0][lane] : 0, wr = lane < (nth >> 5) ? s_scan[1][lane] : 0;
      int xc = wc, xr = wr;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int tc = __shfl_up_sync(0xffffffffu, xc, d), tr = __shfl_up_sync(0xffffffffu, xr, d);
        if (lane >= d) { xc += tc; xr += tr; }
      }
      s_scan[0][lane] = xc - wc; s_scan[1][lane] = xr - wr;   // exclusive warp offsets
      if (lane == 31) { s_total[0] = xc; s_total[1] = xr; }
    }
    __syncthreads();
    const int exC = baseCnt + s_scan[0][wid] + ic - actv, exR = baseRows + s_scan[1][wid] + ir - km;
    if (actv) { rv.act[exC] = s; rv.rowoff[exC] = exR; }
    if (s < S) rv.posIncl[s] = exC + actv;
    baseCnt += s_total[0]; baseRows += s_total[1];
    __syncthreads();
  }
  const int nact = baseCnt;
  if (tid == 0) rv.rowoff[nact] = baseRows;
  __syncthreads();
  // staging chunks: chunk c = active sites whose row offset lies in [cQ, (c+1)Q); Q >= Kcap so no chunk is empty
  const int nchunks = nact ? rv.rowoff[nact - 1] / Q + 1 : 0;
  for (int i = tid; i < nact; i += nth) {
    const int c = rv.rowoff[i] / Q;
    if (i == 0 || rv.rowoff[i - 1] / Q != c) rv.chunkStart[c] = i;
  }
  if (tid == 0) rv.chunkStart[nchunks] = nact;
  __syncthreads();

  // phase B
  float4* const buf0 = s_dyn;                                     // 2 staging buffers of 2Q rows each
  int* const offs0 = reinterpret_cast<int*>(s_dyn + (size_t)4 * Q * V);   // 2 x (Q+1) row offsets
  float c[RP];
#pragma unroll
  for (int r = 0; r < RP; r++) c[r] = r < v.R ? rv.carryIn[r] : 0.0f;

  auto stage = [&](int ch, int ltid, int nld, int b, bool loadersOnly) {
    float4* const bufb = buf0 + (size_t)b * 2 * Q * V;
    int* const offb = offs0 + b * (Q + 1);
    const int i0 = rv.chunkStart[ch], i1 = rv.chunkStart[ch + 1];
    const int r0 = rv.rowoff[i0], nrows = rv.rowoff[i1] - r0, ns = i1 - i0;
    for (int k = ltid; k <= ns; k += nld) offb[k] = rv.rowoff[i0 + k] - r0;
    if (loadersOnly) asm volatile("bar.sync 1, %0;" ::"r"(nld)); else __syncthreads();
    for (int q = ltid; q < nrows; q += nld) {
      int lo = 0, hi = ns;                        // last k with offs[k] <= q
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offb[mid] <= q) lo = mid; else hi = mid; }
      const int site = rv.act[i0 + lo], j = q - offb[lo];
      const float4* src = reinterpret_cast<const float4*>(rv.elist + ((size_t)site * rv.Kcap + j) * RP);
#pragma unroll
      for (int u = 0; u < V; u++) bufb[(size_t)q * V + u] = src[u];
    }
  };

  if (nchunks > 0) stage(0, tid, nth, 0, false);
  __syncthreads();
  for (int ch = 0; ch < nchunks; ch++) {
    const int b = ch & 1;
    if (wid > 0) {
      if (ch + 1 < nchunks) stage(ch + 1, tid - 32, nth - 32, b ^ 1, true);
    } else if (tid == 0) {
      const int i0 = rv.chunkStart[ch], ns = rv.chunkStart[ch + 1] - i0;
      const float4* const bufb = buf0 + (size_t)b * 2 * Q * V;
      const int* const offb = offs0 + b * (Q + 1);
      for (int k = 0; k < ns; k++) {
        const int o0 = offb[k], o1 = offb[k + 1];
        for (int q = o0; q < o1; q++) {
#pragma unroll
          for (int u = 0; u < V; u++) {
            const float4 x = bufb[(size_t)q * V + u];
            c[4 * u + 0] = __fadd_rn(c[4 * u + 0], x.x); c[4 * u + 1] = __fadd_rn(c[4 * u + 1], x.y);
            c[4 * u + 2] = __fadd_rn(c[4 * u + 2], x.z); c[4 * u + 3] = __fadd_rn(c[4 * u + 3], x.w);
          }
        }
        float m = c[0];
#pragma unroll
        for (int r = 1; r < RP; r++) if (r < v.R) m = fmaxf(m, c[r]);
#pragma unroll
        for (int r = 0; r < RP; r++) c[r] = r < v.R ? __fsub_rn(c[r], m) : 0.0f;
        float4* co = reinterpret_cast<float4*>(rv.cout + (size_t)(i0 + k) * RP);
#pragma unroll
        for (int u = 0; u < V; u++) co[u] = make_float4(c[4 * u], c[4 * u + 1], c[4 * u + 2], c[4 * u + 3]);
        rv.mxo[i0 + k] = m;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
#pragma unroll
    for (int r = 0; r < RP; r++) rv.carryOut[r] = c[r];
  }
}

// ------------------------------------------------------------------------------------------------
// K4c final: rateAveragedLogLiks (AugTree.cpp:317-323) per site, then a deterministic sum: fixed tree inside
// each 256-site chunk, chunk sums combined by reduce_chunks order (identical on host for the multi-GPU path).
// mode 1: exponent vectors come from the chain; mode 0: each site folds its own list from zero (numRateCats == 1,
// where the carry is always exactly zero, or the textbook reduction).
__device__ __forceinline__ double reduce_chunks_device(const double* chunk, int n, double* s_red) {
  double acc = 0.0;
  for (int cI = threadIdx.x; cI < n; cI += CHUNK) acc = __dadd_rn(acc, chunk[cI]);
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int d = CHUNK / 2; d > 0; d >>= 1) {
    if (threadIdx.x < d) s_red[threadIdx.x] = __dadd_rn(s_red[threadIdx.x], s_red[threadIdx.x + d]);
    __syncthreads();
  }
  return s_red[0];
}

__global__ void __launch_bounds__(CHUNK) k_final(DevView v, RootView rv, int mode, int reduceAll) {
  __shared__ double s_red[CHUNK];
  __shared__ bool s_last;
  const int s = blockIdx.x * CHUNK + threadIdx.x;
  double ll = 0.0;
  if (s < v.S) {
    float ev[RMAX];
    float mx = 0.0f;
    const int km = min(rv.kmax[s], rv.Kcap);
    if (mode == 1) {
      const int p = rv.posIncl[s];
#pragma unroll
      for (int r = 0; r < RMAX; r++) if (r < v.R) ev[r] = p ? rv.cout[(size_t)(p - 1) * rv.RP + r] : rv.carryIn[r];
      if (km > 0) mx = rv.mxo[p - 1];
    } else {
#pragma unroll
      for (int r = 0; r < RMAX; r++) ev[r] = 0.0f;
      for (int j = 0; j < km; j++)
#pragma unroll
        for (int r = 0; r < RMAX; r++) if (r < v.R) ev[r] = __fadd_rn(ev[r], rv.elist[((size_t)s * rv.Kcap + j) * rv.RP + r]);
      mx = ev[0];
#pragma unroll
      for (int r = 1; r < RMAX; r++) if (r < v.R) mx = fmaxf(mx, ev[r]);
#pragma unroll
      for (int r = 0; r < RMAX; r++) if (r < v.R) ev[r] = __fsub_rn(ev[r], mx);
    }
    // mean(likProp % exp(ev)): expf on the fp32 vector, promoted to double; two interleaved accumulators
    double a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int r = 0; r < RMAX; r++)
      if (r < v.R) {
        const double w = (double)(float)exp((double)ev[r]);   // correctly rounded expf
        const double x = __dmul_rn(rv.rootdot[(size_t)r * v.Sp + s], w);
        if (r & 1) a2 = __dadd_rn(a2, x); else a1 = __dadd_rn(a1, x);
      }
    ll = __dadd_rn(log(__dadd_rn(a1, a2) / (double)v.R), (double)mx);
    rv.sitell[s] = ll;
  }
  s_red[threadIdx.x] = ll;
  __syncthreads();
  for (int d = CHUNK / 2; d > 0; d >>= 1) {
    if (threadIdx.x < d) s_red[threadIdx.x] = __dadd_rn(s_red[threadIdx.x], s_red[threadIdx.x + d]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    rv.chunk[blockIdx.x] = s_red[0];
    __threadfence();
    s_last = atomicAdd(rv.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && reduceAll) {
    __threadfence();
    const double tot = reduce_chunks_device(rv.chunk, gridDim.x, s_red);
    if (threadIdx.x == 0) { *rv.result = tot; *rv.ticket = 0; }
  } else if (s_last && threadIdx.x == 0) {
    *rv.ticket = 0;
  }
}

}  // namespace dmpc
