// kernels.cuh — sm_100a CUDA kernels of the Felsenstein-pruning likelihood engine.
//
// Data layout in HBM (per handle):
//   tipmask  u8     [N][Sp]                 4-bit state masks, Sp = numLoci padded to a multiple of 256
//   part[s]  f64x4  [N][R][Sp]              conditional-likelihood 4-vectors, one 32-byte element per (internal
//                                           vertex, rate, site); row N-1 is scratch.  Two slots s = 0,1 so that a
//                                           rejected proposal rolls back by flipping a per-vertex slot bit.  Rows
//                                           of FUSED vertices (small clades, see k_prune) are never written.
//   expo[s]  f32    [N][R][Sp]              per-vertex rescale exponents (IntermediateNode.cpp:63: a FLOAT)
//   flag[s]  u8     [R][T][N up to 16]      1 when any site of the 256-site tile rescaled at that vertex for that
//                                           rate; expo tiles whose flag is 0 are never written nor read
// Consecutive threads touch consecutive 32-byte elements (LDG.E.256 / STG.E.256), so every warp request is 1 KB
// fully coalesced.  There is no GEMM shape here (a 4x4 mat-vec per thread): no tensor cores, no TMA tiles.
//
// Arithmetic follows the reference's evaluation order with separately rounded products and sums
// (__dmul_rn/__dadd_rn: never contracted into FMAs), so partials are bit-identical to the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <limits.h>

#include "task.hpp"

namespace dmpc {

constexpr int RMAX = 8;         // rate categories supported (the packaged data uses 3)
constexpr int TILE = 256;       // sites per block of the pruning kernel = flag granularity
constexpr int PT = 128;         // threads of a pruning block
constexpr int SPT = TILE / PT;  // sites per pruning thread
constexpr int RCH = 64;         // sites per deterministic reduction chunk: site blocks of a sharded run are multiples of it,
                                // so that chunk boundaries — and the summation tree — do not depend on the number of ranks
constexpr int CPT = TILE / RCH; // reduction chunks per site tile
constexpr int RED = 256;        // strided accumulators of the final sum over the chunk sums (reduce_chunks)
static_assert(TILE % RCH == 0 && RED == TILE, "the root kernels reduce CPT chunks per block with TILE threads");

struct DevView {
  double* part[2];
  float* expo[2];
  uint8_t* flag[2];
  const uint8_t* tipmask;
  uint8_t* cur;                 // [N] slot of each internal vertex, as seen by the root reduction
  const double* lut;            // [2][R][16][4]: P * indicator(mask) for every 4-bit mask (tips)
  int N, S, Sp, R, T, nInt;      // nInt = rows of the per-vertex arrays: the N-1 internal vertices + 1 scratch row
  int nReal;                    // N-1: what the root reduction scans
  int nFlag;                    // bytes of one (rate, tile) row of the flag arrays: nInt rounded up to 16, so that the root
                                // reduction scans 16 vertices per 128-bit load
  int tileOff;                  // first site tile of a pruning launch (0, except while logLikCpp prunes the alignment slab by
                                // slab behind its upload)
  // select instead of dynamic indexing: indexing a by-value kernel parameter array would spill it to local memory
  __device__ __forceinline__ double* partOf(int slot) const { return slot ? part[1] : part[0]; }
  __device__ __forceinline__ float* expoOf(int slot) const { return slot ? expo[1] : expo[0]; }
  __device__ __forceinline__ uint8_t* flagOf(int slot) const { return slot ? flag[1] : flag[0]; }
};

struct DevStatus {
  double ll;                    // the log-likelihood (valid when the kernel that owns the last ticket reduced)
  int overflow;                 // longest per-site exponent list seen, when it exceeds the capacity
  int badCells;                 // alignment cells without any state
  int violation;                // a fused clade may have rescaled (k_prune, fast mode)
  int nact;                     // site tiles holding a rescaled site (set by k_root_gather)
  unsigned ticket, ticket2;     // last-block detection of k_final / k_root_gather
  int xok;                      // site-sharded runs: 1 = the in-kernel peer exchange produced ll, 0 = host protocol needed,
                                //   2 = a peer never answered, -1 = no exchange attempted
  int cert;                     // chain mode: CERT_* — how certify() resolved the cross-locus exponent chain
  int chainSites, chainRows;    // what the literal walker (k_chain, or certify()'s prefix) walked: rescaled sites and rows
  unsigned pubSeq;              // number of times a root kernel has published this block to the host mirror
  int certL0;                   // rescaled sites of the literal prefix (chunks 0 .. certTile)
  int certRate;                 // the dominant rate category of the certified regime
  int certTile;                 // last (local) RCH-site chunk of the literal prefix; every later chunk is certified
  // %globaltimer stamps (ns) of the root stage, for the latency breakdown dmpc_get_root_timeline reports:
  //   0 first gather block starts    1 gather's last block found    2 exchange / quiet reduction done    3 certificate done
  //   4 first final block starts     5 final's last block found     6 final's exchange / reduction done
  unsigned long long stamp[8];
};
static_assert(sizeof(DevStatus) == 128, "status block = 32 words");
// Programmatic dependent launch: a kernel launched with the stream-serialisation attribute may be SCHEDULED while its
// predecessor still runs; nothing the predecessor writes may be touched before pdl_wait() (a no-op for an ordinary
// launch).  Every block signals at once that its successor may be scheduled: the successor's launch latency and block
// ramp-up then overlap this grid's tail instead of following it.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
enum : int {
  CERT_NONE = 0,                // certify() has not run (or had nothing to do: no site rescaled anywhere)
  CERT_OK = 1,                  // every tile behind the literal prefix is certified: k_final uses the closed form there
  CERT_LITERAL = 2              // the certificate does not hold (or the prefix is too long): k_chain + k_final(mode 1)
};

// Peer mailboxes of a site-sharded run (one process per GPU of the box, NVLink/NVSwitch peer access — or several shards
// of one process): every rank owns a mailbox in its HBM that all peers can store into.  Layout of a mailbox: unsigned
// flag[CW] (sequence number of the last message of each sender), then double payload[2][CW][width] (double-buffered by
// sequence parity): payload[.][sender] = { flags, nChunks, nRecords, 0, chunk sums[nChunks], chunk records[nRecords][XREC] }.
constexpr int CW = 8;            // ranks
constexpr int XHDR = 4;          // doubles in front of the chunk sums of a message
enum : int { XF_ACTIVE = 1 };    // message flags
// record of the chain certificate, one per 64-site chunk (see certify): nAct, ops, absSum, 0, S_r[RP], maxPre[RP][RP]
__host__ __device__ constexpr int xrec(int RP) { return 4 + RP + RP * RP; }
constexpr int XREC_MAX = xrec(RMAX);
struct CommView {
  int world, rank, width;
  unsigned* seq;                // this rank's message counter (device memory)
  unsigned* flag[CW];           // flag arrays of every rank's mailbox (peer pointers)
  double* payload[CW];          // payload areas of every rank's mailbox (peer pointers)
  unsigned* cflag[CW];          // same for the exponent-vector hand-off of the literal cross-locus chain:
  float* carry[CW];             //   carry[2][CW][RMAX], flag value = sequence number of the evaluation
};

struct RootView {
  float* elist;                 // [Sp][Kcap][RP] per-site non-zero exponents in vertex-id order, zero padded
  int* eid;                     // [Sp][Kcap][RP] the vertex (internal index) each of those rows belongs to (committed lists only)
  int* kmax;                    // [Sp] rows used at each site (uncapped)
  const float* baseElist;       // proposal batches: the COMMITTED state's lists (elist / eid / kmax of the handle, refreshed
  const int* baseEid;           //   by dmpc_eval_proposals), into which every proposal merges its path vertices' rows
  const int* baseKmax;
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
  double* chunk;                // [T * CPT] sums of RCH consecutive sites (the first nChunks are real)
  double* tileRec;              // [T * CPT][xrec(RP)] per-chunk records of the chain certificate (k_root_gather -> certify)
  int* certDom;                 // [records of ALL ranks] scratch of certify(): the entry leader of every record
  int* nactCtr;                 // device-only: site tiles with a rescaled site, counted by the gather's blocks; its last block moves
                                // the count into st->nact and re-arms the counter (no memset between evaluations)
  int nChunks;                  // ceil(S / RCH)
  int publishAtGather;          // 1: the gather's last block publishes the status (nothing is queued behind it)
  int certOn;                   // 0: certify() always answers CERT_LITERAL (tests compare both paths)
  DevStatus* st;                // everything the host reads back after an evaluation, in one copy
  DevStatus* hostSt;            // mapped pinned mirror (or null): the last root kernel stores the finished block there
                                // and bumps its pubSeq, so that the host can pick the result up by polling instead of
                                // queueing a copy and synchronising the stream
  int Kcap, RP, Q;              // Q = rows per chain staging chunk
  CommView comm;                // comm.world > 1: the root kernels also perform the cross-rank exchange
  int exactMode;                // the prune kernels ran in exact mode (a raised violation flag is then stale)
};

// Proposal batches (dmpc_eval_proposals): proposal b differs from the committed state only at the stored vertices of
// its path to the root.  Their new exponents / tile flags / the new dot(root, pi) live in per-proposal scratch; the
// batched root kernels (BATCH = true) read those for path vertices and the committed arrays for everything else, and
// write per-proposal copies of the reduction's work arrays.
struct PathView {
  const int32_t* ids;           // [B][JMAX] internal indices of the path vertices, ascending, padded with INT_MAX
  const int32_t* pos;           // [B][JMAX] position in the run of ids[.]
  float* pe;                    // [B][JMAX][R][Sp] new exponents (written where the tile flag is set)
  uint8_t* pf;                  // [B][JMAX][R][T] tile flags
  double* rootdot;              // [B][R][Sp]
  int JMAX;
};
__device__ __forceinline__ RootView forProposal(RootView rv, int b, int Sp, int T) {
  const size_t sp = (size_t)Sp;
  rv.elist += (size_t)b * sp * rv.Kcap * rv.RP;
  rv.kmax += (size_t)b * sp;
  rv.act += (size_t)b * sp;
  rv.rowoff += (size_t)b * (sp + 1);
  rv.posIncl += (size_t)b * sp;
  rv.chunkStart += (size_t)b * (sp + 2);
  rv.cout += (size_t)b * sp * rv.RP;
  rv.mxo += (size_t)b * sp;
  rv.carryOut += (size_t)b * RMAX;
  rv.sitell += (size_t)b * sp;
  rv.chunk += (size_t)b * T * CPT;
  rv.tileRec += (size_t)b * T * CPT * xrec(rv.RP);
  rv.certDom += (size_t)b * T * CPT;
  rv.st += b;
  rv.comm.world = 1;
  return rv;
}

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

// ------------------------------------------------------------------------------------------------
// K2 prune: the dirty STORED vertices of one level x all sites x all rate categories.
// Replaces IntermediateNode::ComputeSolutions/ComputeSolution (IntermediateNode.cpp:26-71) driven by
// AugTree::TrySolve (AugTree.cpp:114-134).  grid = (R, 256-site tiles, tasks of the level), 128 threads; one
// thread owns two sites of one rate of its task: 256-bit child loads (or mask bytes), `sol = (P c0) % (P c1)` with
// products and sums rounded separately in Armadillo's order, the reference's two rescale checks, 256-bit stores.
//
// Small-clade fusion.  An operand of a task may be a FUSED clade (<= 15 tips): its vertices are never stored;
// the thread recomputes them for its (site, rate) from the tip masks by running the clade's token program
// (task.hpp) on a register accumulator + a shared-memory stack, then feeds the result into the stored vertex's
// update.  The masks of all tips a task's programs read, and the tokens, are staged in shared memory in one go.  HBM traffic drops
// from ~68 B per update to the stored vertices only (about a fifth of the vertices of a random tree for
// fuseTips = 8), which moves the kernel off the HBM roof towards the fp64 pipe.
//
// Rescaling inside a fused clade is legal but needs per-vertex exponents and tile flags.  FAST mode (EXACT =
// false) assumes it does not happen, clears the fused vertices' flags, and raises *violation when a thread
// sees a product that could have triggered it (conservative high-word test); the host then re-runs the
// evaluation with EXACT = true, where every fused vertex gets the reference's full two-check treatment, a
// block vote, and exponent rows — and keeps that mode for the handle.
//
// A launch covers one ROUND of the plan (host_tree.hpp: the dirty tree split into vertical paths, Strahler-numbered):
// block column z walks the tasks of path z bottom-up in order — a thread only ever reads partials of its own (site,
// rate), so the child -> parent dependencies inside a path are plain program order (its own store, then its own load).
__device__ __forceinline__ void matvec4(const double* __restrict__ P, const double (&x)[4], double (&y)[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) y[i] = row4(P + 4 * i, x[0], x[1], x[2], x[3]);
}
__device__ __forceinline__ bool below4(const double (&s)[4]) {            // exact: max(s) < 1e-150 (no NaNs here)
  return s[0] < 1e-150 && s[1] < 1e-150 && s[2] < 1e-150 && s[3] < 1e-150;
}
__device__ __forceinline__ bool maybeBelow4(const double (&s)[4]) {       // conservative, integer pipe only
  constexpr int H = 0x20CA2FE8;                                           // > high word of 1e-150 (0x20CA2FE7)
  return __double2hiint(s[0]) < H && __double2hiint(s[1]) < H && __double2hiint(s[2]) < H && __double2hiint(s[3]) < H;
}
// IntermediateNode.cpp:53-66 for one category: first child, check, second child, check (the second OVERWRITES)
// The exact test is guarded by the conservative high-word one: four independent integer compares instead of a chain of
// four dependent fp64 compares on the common (no rescale) path.  below4 implies maybeBelow4 for every non-NaN input.
__device__ __forceinline__ void combineExact(const double (&first)[4], const double (&second)[4], double (&s)[4], float& e) {
#pragma unroll
  for (int i = 0; i < 4; i++) s[i] = first[i];                  // ones % y == y exactly
  if (maybeBelow4(s) && below4(s)) {
    const double m = fmax(fmax(s[0], s[1]), fmax(s[2], s[3]));
    s[0] = s[0] / m; s[1] = s[1] / m; s[2] = s[2] / m; s[3] = s[3] / m;
    e = (float)log(m);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) s[i] = __dmul_rn(s[i], second[i]);
  if (maybeBelow4(s) && below4(s)) {
    const double m = fmax(fmax(s[0], s[1]), fmax(s[2], s[3]));
    s[0] = s[0] / m; s[1] = s[1] / m; s[2] = s[2] / m; s[3] = s[3] / m;
    e = (float)log(m);
  }
}

// PATHS = true is K3, the proposal-batch variant (dmpc_eval_proposals): blockIdx.z is a PROPOSAL, whose run of tasks —
// the stored vertices on the path from the vertex it changes up to the root — the block walks bottom-up with the path
// partial in registers (TF_CxFWD operands), reading only the off-path siblings (stored partials, tip masks, fused
// clades) and writing nothing to the committed state: the path vertices' new rescale exponents and tile flags go to
// per-proposal scratch for the batched root reduction, and dot(root, pi) for its (proposal, rate, site).  A possible
// rescale inside a FUSED clade marks the proposal invalid (the host evaluates those one by one).
struct PruneArgs {
  const Task* tasks;
  int nTasks;
  const Token* prog;
  const int32_t* tipList;
  int* violation;               // fast mode: raised when a fused clade may have rescaled
  int flip;                     // 0/1, XORed into every slot bit of the tasks and tokens: a cached full-evaluation plan is
                                // replayed with all slots toggled instead of being rebuilt and uploaded again
  const int* runStart;          // PATHS: [proposals + 1] task ranges
  const int32_t* curList;       // !PATHS, first launch of an evaluation: slot bits of the last rollback the device has not seen yet
  int curN;                     //   ((internal index << 1) | slot; disjoint from what this plan writes), applied by block (0, 0, 0)
  const int32_t* segStart;      // !PATHS: block column z walks the tasks [segStart[z], segStart[z + 1]) in order (null: task z,
                                // stride gridDim.z)
  int* invalid;                 // PATHS: [proposals]
  PathView pv;                  // PATHS: per-proposal outputs
};

// SPT_ = sites per thread (the block has TILE / SPT_ threads): what is uniform across the block — token decoding, the P
// matrix entries (LDCU), addresses — is paid once per SPT_ updates, and the SPT_ sites are independent instruction streams
// for the scheduler; MINB = resident blocks per SM the register allocation aims at.
template <bool EXACT, bool PATHS, int SPT_, int MINB>
__global__ void __launch_bounds__(TILE / SPT_, MINB) k_prune(DevView v, PruneArgs pa_) {
  constexpr int PT_ = TILE / SPT_;
  static_assert(!(EXACT && PATHS), "proposal batches run in fast mode only");
  pdl_wait();
  const Task* __restrict__ tasks = pa_.tasks;
  const Token* __restrict__ prog = pa_.prog;
  const int32_t* __restrict__ tipList = pa_.tipList;
  const int flip = PATHS ? 0 : pa_.flip;
  // one thread = SPT_ sites (tid, tid + PT_, ...) of one rate: what is uniform across the block — token decoding, the
  // P matrix elements, addresses — is paid once per SPT_ updates
  __shared__ __align__(16) double s_lut[2 * 64];                 // [set][mask][4] of this block's rate
  __shared__ __align__(16) double s_stack[FUSE_DEPTH * 4 * TILE];
  __shared__ __align__(16) unsigned char s_mask[2 * FUSE_MAX_TIPS * TILE];
  __shared__ __align__(16) int4 s_tok[3 * FUSE_MAX_TIPS + 4];
  const int tid = threadIdx.x;
  const int tile = blockIdx.y + v.tileOff, r = blockIdx.x, R = v.R;
  const int site0 = tile * TILE + tid;
  const size_t Sp = (size_t)v.Sp;
  const size_t flagRow = ((size_t)r * v.T + tile) * v.nFlag;
  bool lutLoaded = false, trig = false;
  double prev[PATHS ? SPT_ : 1][4];               // PATHS: the previous task's output = the on-path operand of this one
  if (!PATHS && pa_.curN > 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    for (int i = threadIdx.x; i < pa_.curN; i += PT_) v.cur[pa_.curList[i] >> 1] = (uint8_t)(pa_.curList[i] & 1);
  const bool segs = !PATHS && pa_.segStart != nullptr;
  const int tiBegin = PATHS ? pa_.runStart[blockIdx.z] : segs ? pa_.segStart[blockIdx.z] : (int)blockIdx.z;
  const int tiEnd = PATHS ? pa_.runStart[blockIdx.z + 1] : segs ? pa_.segStart[blockIdx.z + 1] : pa_.nTasks;
  const int tiStep = (PATHS || segs) ? 1 : (int)gridDim.z;

  for (int ti = tiBegin; ti < tiEnd; ti += tiStep) {
    const int4 tk0 = __ldg(reinterpret_cast<const int4*>(tasks + ti));
    const int4 tk1 = __ldg(reinterpret_cast<const int4*>(tasks + ti) + 1);
    const int fl = tk0.w ^ (flip ? (TF_OUTSLOT | TF_C0SLOT | TF_C1SLOT) : 0);
    const int set = (fl & TF_WITHIN) ? 1 : 0;
    const bool tip0 = (fl & TF_C0TIP) != 0, tip1 = (fl & TF_C1TIP) != 0;
    const bool fus0 = (fl & TF_C0FUSED) != 0, fus1 = (fl & TF_C1FUSED) != 0;
    const bool fwd0 = PATHS && (fl & TF_C0FWD) != 0, fwd1 = PATHS && (fl & TF_C1FWD) != 0;
    const double* pa = v.partOf(fl & TF_C0SLOT) + (((size_t)tk0.y * R + r) * Sp + site0) * 4;
    const double* pb = v.partOf(fl & TF_C1SLOT) + (((size_t)tk0.z * R + r) * Sp + site0) * 4;
    const bool plain = !(fus0 | fus1);
    const int n0 = tk1.y & 0xffff, n1 = (tk1.y >> 16) & 0xffff, ntok = n0 + n1;
    unsigned ma[SPT_], mb[SPT_];
    double a[SPT_][4], b[SPT_][4];
    // every load this thread needs from HBM is issued here, before any arithmetic
#pragma unroll
    for (int u = 0; u < SPT_; u++) {
      if (tip0) ma[u] = v.tipmask[(size_t)tk0.y * Sp + site0 + u * PT_];
      else if (!fus0 && !fwd0) ld256(pa + (size_t)u * PT_ * 4, a[u][0], a[u][1], a[u][2], a[u][3]);
      if (tip1) mb[u] = v.tipmask[(size_t)tk0.z * Sp + site0 + u * PT_];
      else if (!fus1 && !fwd1) ld256(pb + (size_t)u * PT_ * 4, b[u][0], b[u][1], b[u][2], b[u][3]);
    }
    if (!plain) {
      // the masks of every tip the programs read (4 sites per cp.async) and the tokens themselves, into shared memory
      const int nTips = tk1.w;
      for (int j = tid; j < nTips * (TILE / 4); j += PT_) {
        const int tipIdx = j / (TILE / 4), q = j - tipIdx * (TILE / 4);
        const unsigned char* src = v.tipmask + (size_t)__ldg(tipList + tk1.z + tipIdx) * Sp + tile * TILE + 4 * q;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s_mask + tipIdx * TILE + 4 * q);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (tid < ntok) s_tok[tid] = __ldg(reinterpret_cast<const int4*>(prog + tk1.x + tid));
    }
    if (!lutLoaded && (!plain || tip0 || tip1)) {
      for (int i = tid; i < 2 * 64; i += PT_) s_lut[i] = v.lut[(size_t)((i >> 6) * R + r) * 64 + (i & 63)];
      lutLoaded = true;
      if (plain) __syncthreads();
    }
    if (!plain) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }

    double reg[SPT_][4];
    // ---- the fused operands' programs
    auto run = [&](int first, int n, int sp) {
      for (int i = first; i < first + n; i++) {
        const int4 t4 = s_tok[i];
        const int op = t4.x & 7;
        if (op == OP_SPILL) {
#pragma unroll
          for (int u = 0; u < SPT_; u++)
#pragma unroll
            for (int k = 0; k < 4; k++) s_stack[(sp * 4 + k) * TILE + tid + u * PT_] = reg[u][k];
          sp++;
          continue;
        }
        const int tset = (t4.x & TOK_SET) ? 1 : 0;
        const double* P = c_P[tset][r];
        const double* lutr = s_lut + tset * 64;
        if (op == OP_SLOTREG) sp--;
        float e[SPT_];
#pragma unroll
        for (int u = 0; u < SPT_; u++) {
          double x[4], y[4];                     // x = term of the operand the token lists first, y = the other
          if (op == OP_CHERRY) {
            const unsigned m0 = s_mask[t4.y * TILE + tid + u * PT_], m1 = s_mask[t4.z * TILE + tid + u * PT_];
#pragma unroll
            for (int k = 0; k < 4; k++) { x[k] = lutr[4 * m0 + k]; y[k] = lutr[4 * m1 + k]; }
          } else if (op == OP_REGTIP) {
            const unsigned m1 = s_mask[t4.z * TILE + tid + u * PT_];
            matvec4(P, reg[u], x);
#pragma unroll
            for (int k = 0; k < 4; k++) y[k] = lutr[4 * m1 + k];
          } else {
            double z[4];
#pragma unroll
            for (int k = 0; k < 4; k++) z[k] = s_stack[(sp * 4 + k) * TILE + tid + u * PT_];
            matvec4(P, z, x);
            matvec4(P, reg[u], y);
          }
          if (EXACT) {
            // child order: CHERRY lists the first child first; REGTIP's flag says the tip (y) is the reference's
            // first child, SLOTREG's that the popped operand (x) is
            const bool firstFlag = (t4.x & TOK_FIRST) != 0;
            const bool xFirst = op == OP_CHERRY ? true : op == OP_REGTIP ? !firstFlag : firstFlag;
            e[u] = 0.0f;
            double s4[4];
            if (xFirst) combineExact(x, y, s4, e[u]); else combineExact(y, x, s4, e[u]);
#pragma unroll
            for (int k = 0; k < 4; k++) reg[u][k] = s4[k];
          } else {
            // products commute bit for bit, so the child order only matters for the rescale checks.  With transition
            // probabilities <= 1 (the host verified it, else the handle starts in exact mode) no term exceeds 1, so a
            // first term below 1e-150 drags the product below it too: one conservative test of the product suffices
#pragma unroll
            for (int k = 0; k < 4; k++) reg[u][k] = __dmul_rn(x[k], y[k]);
            trig |= maybeBelow4(reg[u]);
          }
        }
        if (EXACT) {
          const int tslot = ((t4.x & TOK_SLOT) ? 1 : 0) ^ flip;
          bool resc = false;
#pragma unroll
          for (int u = 0; u < SPT_; u++) resc |= e[u] != 0.0f;
          const int any = __syncthreads_or(resc);
          if (any) {
#pragma unroll
            for (int u = 0; u < SPT_; u++) v.expoOf(tslot)[((size_t)t4.w * R + r) * Sp + site0 + u * PT_] = e[u];
          }
          if (tid == 0) v.flagOf(tslot)[flagRow + t4.w] = (uint8_t)(any != 0);
        }
      }
    };
    if (fus0) {
      run(0, n0, 0);
      if (fus1) {
#pragma unroll
        for (int u = 0; u < SPT_; u++)
#pragma unroll
          for (int k = 0; k < 4; k++) s_stack[k * TILE + tid + u * PT_] = reg[u][k];
      }
    }
    if (fus1) run(n0, n1, 1);
    if (!PATHS && !plain && tid < ntok) {
      // bookkeeping of the fused vertices, one token per thread: current slot, and (fast mode) "nothing rescaled"
      const int4 t4 = s_tok[tid];
      if ((t4.x & 7) != OP_SPILL) {
        const int tslot = ((t4.x & TOK_SLOT) ? 1 : 0) ^ flip;
        if (!EXACT) v.flagOf(tslot)[flagRow + t4.w] = 0;
        if (tile == 0 && r == 0) v.cur[t4.w] = (uint8_t)tslot;
      }
    }

    // ---- the stored vertex itself: exact semantics always
    const double* P = c_P[set][r];
    const double* lutr = s_lut + set * 64;
    const int os = (fl & TF_OUTSLOT) ? 1 : 0;
    const size_t o = ((size_t)tk0.x * R + r) * Sp + site0;
    float e[SPT_];
    bool rescaled = false;
#pragma unroll
    for (int u = 0; u < SPT_; u++) {
      double first[4], second[4], s4[4];
      if (tip0) {
#pragma unroll
        for (int k = 0; k < 4; k++) first[k] = lutr[4 * ma[u] + k];
      } else if (fus0) {
        if (fus1) {
          double z[4];
#pragma unroll
          for (int k = 0; k < 4; k++) z[k] = s_stack[k * TILE + tid + u * PT_];
          matvec4(P, z, first);
        } else {
          matvec4(P, reg[u], first);
        }
      } else if (fwd0) {
        matvec4(P, prev[PATHS ? u : 0], first);
      } else {
        matvec4(P, a[u], first);
      }
      if (tip1) {
#pragma unroll
        for (int k = 0; k < 4; k++) second[k] = lutr[4 * mb[u] + k];
      } else if (fus1) {
        matvec4(P, reg[u], second);
      } else if (fwd1) {
        matvec4(P, prev[PATHS ? u : 0], second);
      } else {
        matvec4(P, b[u], second);
      }
      e[u] = 0.0f;
      combineExact(first, second, s4, e[u]);
      if (PATHS) {
#pragma unroll
        for (int k = 0; k < 4; k++) prev[PATHS ? u : 0][k] = s4[k];
      } else {
        st256(v.partOf(os) + (o + (size_t)u * PT_) * 4, s4[0], s4[1], s4[2], s4[3]);
      }
      rescaled |= e[u] != 0.0f;
    }
    const int any = __syncthreads_or(rescaled);      // also fences s_mask / s_tok / s_stack reuse by the next task of a run
    if (PATHS) {
      const PathView& pv = pa_.pv;
      const size_t row = ((size_t)blockIdx.z * pv.JMAX + (ti - tiBegin)) * R + r;
      if (any) {
#pragma unroll
        for (int u = 0; u < SPT_; u++) pv.pe[row * Sp + site0 + u * PT_] = e[u];
      }
      if (tid == 0) pv.pf[row * v.T + tile] = (uint8_t)(any != 0);
    } else {
      if (any) {
#pragma unroll
        for (int u = 0; u < SPT_; u++) v.expoOf(os)[o + (size_t)u * PT_] = e[u];
      }
      if (tid == 0) {
        v.flagOf(os)[flagRow + tk0.x] = (uint8_t)(any != 0);
        if (tile == 0 && r == 0) v.cur[tk0.x] = (uint8_t)os;
      }
    }
  }
  if (PATHS) {
#pragma unroll
    for (int u = 0; u < SPT_; u++)                  // arma::dot: two interleaved accumulators
      pa_.pv.rootdot[((size_t)blockIdx.z * R + r) * Sp + site0 + u * PT_] =
          __dadd_rn(__dadd_rn(__dmul_rn(prev[PATHS ? u : 0][0], c_pi[0]), __dmul_rn(prev[PATHS ? u : 0][2], c_pi[2])),
                    __dadd_rn(__dmul_rn(prev[PATHS ? u : 0][1], c_pi[1]), __dmul_rn(prev[PATHS ? u : 0][3], c_pi[3])));
    if (trig) pa_.invalid[blockIdx.z] = 1;
  } else if (!EXACT && trig) {
    *pa_.violation = 1;
  }
}



// ------------------------------------------------------------------------------------------------
// K2w walk: the trailing CHAIN of a plan — levels that hold one task each, every task consuming the previous one's
// output (a root path after a split / merge / NNI proposal; the spine of a caterpillar).  Evaluated level by level,
// such a chain costs one global-memory round trip per vertex (store, dependent load of the next task's operands,
// ~1.3 us measured); here one thread owns one (site, rate) for the whole chain, keeps the running partial in
// registers, and everything that does NOT depend on the chain — the task descriptors and the off-chain operands
// (stored partials, tip mask bytes; fused clades were materialised by the plan builder) — is fetched WALK_DEPTH steps
// ahead with cp.async into a per-thread ring in shared memory, so that a step costs
// its dependent arithmetic only: sol = (P prev) % (P sib) with the reference's two rescale checks
// (IntermediateNode.cpp:53-66, same combineExact as k_prune), one 256-bit store of the vertex's partial.
// grid = (R, site tiles), TILE threads.  Rescale bookkeeping keeps k_prune's format (exponent rows only for tiles
// where some site rescaled, one flag byte per (rate, tile, vertex)) without a barrier per step: a thread that
// rescales writes its own exponent at once and remembers the step in a bit mask; every WALK_WIN steps the block ORs
// the masks, the other threads fill in their zeros for the steps that rescaled somewhere, and the flags are written.
constexpr int WALK_DEPTH = 8;     // cp.async stages in flight per thread
constexpr int WALK_CHUNK = 256;   // steps whose descriptors are staged in shared memory per pass
constexpr int WALK_WIN = 32;      // steps per flag window (bits of the mask)
constexpr int WALK_SMEM = WALK_DEPTH * TILE * 32;   // dynamic shared memory: the ring, 32 B per thread and stage
static_assert((WALK_DEPTH & (WALK_DEPTH - 1)) == 0 && WALK_CHUNK % WALK_WIN == 0 && WALK_WIN == 32, "walk geometry");
// a step as the walking threads see it: everything that is uniform across the block is worked out ONCE per step while
// the chunk is staged (one thread per step), so that the per-step instruction stream of a walking warp — which is what
// bounds a chain: a lone warp issues one dependent instruction every ~5 cycles — stays short.
// x = output row (vertex * R + this rate), y = row of the off-chain operand (stored: vertex * R + rate; tip: the tip),
// z = flags, w = output vertex (for the flag / slot bookkeeping).  Rows, not element offsets: the 64-bit product with
// the row pitch is formed at the point of use, so no handle is too large for the walk
enum : int { WF_OUTSLOT = 1, WF_SIBTIP = 2, WF_SIBSLOT = 4, WF_SET = 8, WF_FWD0 = 16 };

__global__ void __launch_bounds__(TILE, 2) k_walk(DevView v, const Task* __restrict__ tasks, int nSteps, int flip, const int32_t* curList, int curN) {
  pdl_wait();
  extern __shared__ __align__(32) double s_ring[];                 // [WALK_DEPTH][TILE][4]
  __shared__ uint4 s_desc[WALK_CHUNK];
  __shared__ __align__(16) double s_lut[2 * 64];                   // [set][mask][4] of this block's rate
  __shared__ __align__(16) double s_P[2 * 16];                     // both matrix sets of this block's rate
  __shared__ unsigned s_win[3];
  const int tid = threadIdx.x, r = blockIdx.x, tile = blockIdx.y + v.tileOff, R = v.R;
  const unsigned site = (unsigned)(tile * TILE + tid);
  const size_t Sp = (size_t)v.Sp;
  const size_t flagRow = ((size_t)r * v.T + tile) * v.nFlag;
  const int slotMask = flip ? (TF_OUTSLOT | TF_C0SLOT | TF_C1SLOT) : 0;
  if (curN > 0 && blockIdx.x == 0 && blockIdx.y == 0)        // pending rollback slot bits (see PruneArgs::curList)
    for (int i = tid; i < curN; i += TILE) v.cur[curList[i] >> 1] = (uint8_t)(curList[i] & 1);
  if (tid < 2 * 64) s_lut[tid] = v.lut[(size_t)((tid >> 6) * R + r) * 64 + (tid & 63)];
  if (tid < 2 * 16) s_P[tid] = c_P[tid >> 4][r][tid & 15];
  if (tid < 3) s_win[tid] = 0;
  const unsigned ringMine = (unsigned)__cvta_generic_to_shared(s_ring) + (unsigned)tid * 32u;
  double* const part0 = v.part[0];
  double* const part1 = v.part[1];
  const uint8_t* const maskMine = v.tipmask + (site & ~3u);        // the word holding this site's mask byte, row 0

  // the off-chain operand of step k of the staged chunk -> this thread's 32 bytes of ring stage k % WALK_DEPTH
  auto issue = [&](int k) {
    const uint4 d = s_desc[k];
    const unsigned dst = ringMine + (unsigned)(k & (WALK_DEPTH - 1)) * (unsigned)(TILE * 32);
    if (d.z & WF_SIBTIP) {
      const uint8_t* src = maskMine + (size_t)d.y * Sp;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
    } else {
      const double* src = ((d.z & WF_SIBSLOT) ? part1 : part0) + ((size_t)d.y * Sp + site) * 4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u), "l"(src + 2) : "memory");
    }
  };

  // the head: child 0 of the first task comes from memory and then plays the forwarded operand
  double prev[4];
  {
    const int4 d = __ldg(reinterpret_cast<const int4*>(tasks));
    const int fl = d.w ^ slotMask;
    if (fl & TF_C0TIP) {
      const unsigned m = v.tipmask[(size_t)d.y * Sp + site];
#pragma unroll
      for (int k = 0; k < 4; k++) prev[k] = (m >> k) & 1u ? 1.0 : 0.0;   // P * indicator == the lookup table, bit for bit
    } else {
      ld256(v.partOf(fl & TF_C0SLOT) + (((size_t)d.y * R + r) * Sp + site) * 4, prev[0], prev[1], prev[2], prev[3]);
    }
  }

  double P[16];                     // the matrix set of the current step, in registers; reloaded only when the set changes
  int curSet = -1;
  unsigned myMask = 0;              // steps of the current window at which this thread's site rescaled
  int win = 0;
  for (int c0 = 0; c0 < nSteps; c0 += WALK_CHUNK) {
    const int cn = min(WALK_CHUNK, nSteps - c0);
    __syncthreads();                // the previous chunk's descriptors are no longer needed (first pass: s_lut, s_P, s_win)
    for (int j = tid; j < cn; j += TILE) {
      const int4 d = __ldg(reinterpret_cast<const int4*>(tasks + c0 + j));
      const int fl = d.w ^ slotMask;
      const bool fwd0 = (fl & (TF_C0FWD | TF_HEAD)) != 0;
      const int sib = fwd0 ? d.z : d.y;
      const bool sibTip = (fl & (fwd0 ? TF_C1TIP : TF_C0TIP)) != 0;
      uint4 w;
      w.x = (unsigned)(d.x * R + r);
      w.y = sibTip ? (unsigned)sib : (unsigned)(sib * R + r);
      w.z = ((fl & TF_OUTSLOT) ? WF_OUTSLOT : 0) | (sibTip ? WF_SIBTIP : 0) | ((fl & (fwd0 ? TF_C1SLOT : TF_C0SLOT)) ? WF_SIBSLOT : 0) |
            ((fl & TF_WITHIN) ? WF_SET : 0) | (fwd0 ? WF_FWD0 : 0);
      w.w = (unsigned)d.x;
      s_desc[j] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WALK_DEPTH - 1; k++) {
      if (k < cn) issue(k);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int k = 0; k < cn; k++) {
      if (k + WALK_DEPTH - 1 < cn) issue(k + WALK_DEPTH - 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(WALK_DEPTH - 1) : "memory");   // the group of step k has landed

      const uint4 d = s_desc[k];
      const int set = (d.z & WF_SET) ? 1 : 0;
      if (set != curSet) {          // uniform; rare (a root path crosses from within-cluster to between-cluster matrices once)
        curSet = set;
#pragma unroll
        for (int q = 0; q < 16; q++) P[q] = s_P[set * 16 + q];
      }
      const double* mine = s_ring + ((size_t)(k & (WALK_DEPTH - 1)) * TILE + tid) * 4;
      double ct[4], sb[4], s4[4];
      if (d.z & WF_SIBTIP) {
        const unsigned w = *reinterpret_cast<const unsigned*>(mine);
        const unsigned m = (w >> (8 * (tid & 3))) & 0xffu;
        const double* lutr = s_lut + set * 64;
#pragma unroll
        for (int q = 0; q < 4; q++) sb[q] = lutr[4 * m + q];
      } else {
        const double x[4] = {mine[0], mine[1], mine[2], mine[3]};
        matvec4(P, x, sb);
      }
      matvec4(P, prev, ct);
      float e = 0.0f;
      if (d.z & WF_FWD0) combineExact(ct, sb, s4, e); else combineExact(sb, ct, s4, e);
      double* const outp = ((d.z & WF_OUTSLOT) ? part1 : part0) + ((size_t)d.x * Sp + site) * 4;
      st256(outp, s4[0], s4[1], s4[2], s4[3]);
#pragma unroll
      for (int q = 0; q < 4; q++) prev[q] = s4[q];
      if (e != 0.0f) { v.expoOf(d.z & WF_OUTSLOT)[(size_t)d.x * Sp + site] = e; myMask |= 1u << (k & (WALK_WIN - 1)); }

      if ((k & (WALK_WIN - 1)) == WALK_WIN - 1 || k == cn - 1) {
        // end of a window: three rotating masks, so that one barrier per window orders set -> read -> reset
        unsigned* sw = &s_win[win % 3];
        if (myMask) atomicOr(sw, myMask);
        __syncthreads();
        const unsigned bm = *(volatile unsigned*)sw;
        if (tid == 0) s_win[(win + 2) % 3] = 0;
        const int w0 = k & ~(WALK_WIN - 1), cnt = k - w0 + 1;
        unsigned z = bm & ~myMask;      // steps that rescaled somewhere in the tile but not at this site: exponent 0
        while (z) {
          const int b = __ffs(z) - 1;
          z &= z - 1;
          const uint4 dz = s_desc[w0 + b];
          v.expoOf(dz.z & WF_OUTSLOT)[(size_t)dz.x * Sp + site] = 0.0f;
        }
        if (tid < cnt) {
          const uint4 dz = s_desc[w0 + tid];
          const int oz = dz.z & WF_OUTSLOT;
          v.flagOf(oz)[flagRow + dz.w] = (uint8_t)((bm >> tid) & 1u);
          if (tile == 0 && r == 0) v.cur[dz.w] = (uint8_t)oz;
        }
        myMask = 0;
        win++;
      }
    }
  }
}

// rollback: flip the slot bits of the vertices a rejected proposal touched (IntermediateNode.h:26-33)
__global__ void k_set_cur(uint8_t* cur, const int32_t* idxAndSlot, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cur[idxAndSlot[i] >> 1] = (uint8_t)(idxAndSlot[i] & 1);
}

// the uploaded alignment ([N][S] bytes, contiguous) -> the padded device layout [N][Sp]; pad sites get the all-ones
// mask; cells without any state (unknown symbol) or with stray high bits are counted and neutralised
// (columns [c0, c1) of the padded layout: logLikCpp uploads and prepares the alignment slab by slab)
__global__ void k_prepare_masks(const uint8_t* __restrict__ in, uint8_t* __restrict__ mask, int N, int S, int Sp, int c0, int c1, int* bad) {
  const size_t w = (size_t)(c1 - c0);
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (size_t)N * w) return;
  const int s = c0 + (int)(j % w);
  const size_t row = j / w;
  const size_t i = row * Sp + s;
  uint8_t m = 0xF;
  if (s < S) {
    m = in[row * S + s];
    if ((m & 0xF) == 0 || m > 0xF) { atomicAdd(bad, 1); m = 0xF; }
  }
  mask[i] = m;
}

// sum of n chunk sums in a fixed order (RED strided accumulators, then a binary tree); RED threads
__device__ __forceinline__ double reduce_chunks_device(const double* chunk, int n, double* s_red) {
  double acc = 0.0;
  for (int cI = threadIdx.x; cI < n; cI += RED) acc = __dadd_rn(acc, chunk[cI]);
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int d = RED / 2; d > 0; d >>= 1) {
    if (threadIdx.x < d) s_red[threadIdx.x] = __dadd_rn(s_red[threadIdx.x], s_red[threadIdx.x + d]);
    __syncthreads();
  }
  return s_red[0];
}

// per-site terms of one tile (TILE threads) -> its CPT chunk sums: a fixed binary tree inside each run of RCH sites
__device__ __forceinline__ void tile_chunk_sums(double ll, double* s_red, double* chunkOut) {
  const int tid = threadIdx.x;
  s_red[tid] = ll;
  __syncthreads();
  for (int d = RCH / 2; d > 0; d >>= 1) {
    if ((tid & (RCH - 1)) < d) s_red[tid] = __dadd_rn(s_red[tid], s_red[tid + d]);
    __syncthreads();
  }
  if ((tid & (RCH - 1)) == 0) chunkOut[tid / RCH] = s_red[tid];
  __syncthreads();          // every chunk sum is written before the caller's thread 0 fences and takes its ticket
}

// The fused collective of a site-sharded run (called by the LAST block of k_root_gather / k_final, TILE threads).
// This rank's {flags, nChunks, nRecords, chunk sums, chunk records} go straight into every peer's mailbox over NVLink;
// once every peer's message of the same sequence number has arrived, and no rank reports a rescaled site, each rank adds
// ALL chunk sums in the fixed global order (the order of reduce_chunks_device over the concatenated list) and has the
// same log-likelihood, bit for bit, as a single GPU would — without any host-side collective.  The chunk records
// (recs != nullptr: nTiles x XR doubles) are what certify() needs from every rank when some site did rescale.
// Returns 1 = ll written, 0 = some rank holds a rescaled site (certify + k_final must run), 2 = timeout.
__device__ __forceinline__ int exchangeAndReduce(const RootView& rv, int nCh, bool quiet, double* s_red, const double* recs,
                                                 int nTiles, int XR) {
  __shared__ unsigned s_seq;
  __shared__ int s_bad;
  const CommView& cm = rv.comm;
  const int tid = threadIdx.x;
  if (tid == 0) { s_seq = *cm.seq + 1; s_bad = 0; }
  __syncthreads();
  const unsigned seq = s_seq;
  const int nRec = recs ? nTiles * XR : 0;
  const int buf = seq & 1, msg = XHDR + nCh + nRec;
  for (int idx = tid; idx < cm.world * msg; idx += TILE) {
    const int p = idx / msg, j = idx - p * msg;
    const double val = j == 0 ? (quiet ? 0.0 : (double)XF_ACTIVE) : j == 1 ? (double)nCh : j == 2 ? (double)(recs ? nTiles : 0) : j == 3 ? 0.0
                       : j < XHDR + nCh ? rv.chunk[j - XHDR] : recs[j - XHDR - nCh];
    *((volatile double*)cm.payload[p] + ((size_t)buf * CW + cm.rank) * cm.width + j) = val;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < cm.world) {
    *((volatile unsigned*)cm.flag[tid] + cm.rank) = seq;                  // "message seq from me is complete"
    const volatile unsigned* mine = (volatile unsigned*)cm.flag[cm.rank] + tid;
    const long long t0 = clock64();
    while ((int)(*mine - seq) < 0) {
      if (clock64() - t0 > 40000000000LL) { s_bad = 1; break; }            // ~20 s: a peer died; do not hang the GPU
      __nanosleep(100);
    }
  }
  __syncthreads();
  __threadfence_system();
  const volatile double* in = (volatile double*)cm.payload[cm.rank] + (size_t)buf * CW * cm.width;
  bool anyActive = false;
  for (int p = 0; p < cm.world; p++) anyActive |= in[(size_t)p * cm.width] != 0.0;
  int res = s_bad ? 2 : 0;
  if (!s_bad && !anyActive) {
    double acc = 0.0;
    int base = 0;                                                          // global index of rank p's first chunk
    for (int p = 0; p < cm.world; p++) {
      const int cnt = (int)in[(size_t)p * cm.width + 1];
      for (int c = ((tid - base) % RED + RED) % RED; c < cnt; c += RED) acc = __dadd_rn(acc, in[(size_t)p * cm.width + XHDR + c]);
      base += cnt;
    }
    s_red[tid] = acc;
    __syncthreads();
    for (int d = RED / 2; d > 0; d >>= 1) {
      if (tid < d) s_red[tid] = __dadd_rn(s_red[tid], s_red[tid + d]);
      __syncthreads();
    }
    if (tid == 0) rv.st->ll = s_red[0];
    res = 1;
  }
  __syncthreads();
  if (tid == 0) *cm.seq = seq;
  return res;
}

// ------------------------------------------------------------------------------------------------
// certify: the chain WITHOUT walking it, wherever that can be proven exact.
//
// The literal chain (k_chain; AugTree.cpp:306-323) keeps an fp32 vector c, adds every site's exponents to it (vertex
// order), subtracts its maximum m, and the site's term is log(mean_r(rootdot_r * expf(c_r))) + m.  In exact arithmetic
// c_r = E_r - max_q E_q with E the running sums of the sites' exponent sums.  Rescale exponents are ~ -345 each and
// the rate categories do not rescale equally often, so after a few rescaled sites ONE category d leads by far more
// than 104 — and from then on expf(c_r) == 0.0f exactly for every r != d, c_d == 0.0f exactly, and the term of a site
// is log(rootdot_d / R) + (fp32 sum, from zero, of that site's own category-d exponents): no history at all.
// The certificate proves, tile by tile, that the literal fp32 chain is in that regime:
//   * since the maximum entry of c is exactly 0, every c_r equals a DIFFERENCE c_r - c_d*, and differences only pick up
//     the roundings of the additions and subtractions themselves — (2K+1) * 2^-24 * |value| per site with K rows —
//     which add up (no amplification).  B(t) = kappa * sum_{t' <= t} 2^-22 * ops(t') * (gap at entry + abs sums + 1)
//     bounds |c_r - (E_r - E_d)| anywhere up to the end of tile t (kappa = exp(2^-22 * total ops) covers B feeding
//     back into the magnitudes);
//   * tile t is certified for d when for every r != d:  (E_r - E_d at entry) + maxPre[d][r](t) + B(t) < -128.
// All tiles behind the last uncertified one take the closed form in k_final; tiles 0 .. certTile (always including
// tile 0: the chain starts from zeros) are walked literally here, by one thread, from rows staged in shared memory —
// typically a few dozen sites.  If the regime never establishes itself (two categories neck and neck, tiny trees) or
// the prefix is long, the answer is CERT_LITERAL and the host runs k_chain as before.  Runs in the LAST block of
// k_root_gather (all chunk records are complete then; in a sharded run, right after the exchange), TILE threads.  In a
// site-sharded run every rank holds every rank's chunk records (the gather's exchange) and reaches the same verdict; the
// literal prefix must then lie inside rank 0's block.
constexpr int CERT_MAXLIT = 2048;    // rescaled sites the literal prefix may hold

// s_rows: 16 KB of shared memory lent by the caller (the literal prefix's rows are staged there)
template <int RP>
__device__ __forceinline__ void certify(const DevView& v, const RootView& rv, float4* s_rows) {
  constexpr int XR = xrec(RP);
  constexpr int V = RP / 4;
  constexpr int CERT_ROWCAP = 1024 / V;    // rows of one literal tile staged in shared memory (more: read from global memory)
  __shared__ int s_base[CW + 1];
  __shared__ const volatile double* s_rec[CW];
  __shared__ int s_cnt[TILE], s_site[TILE];
  __shared__ int s_scan[2][TILE / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, R = v.R;
  const CommView& cm = rv.comm;
  const bool sharded = cm.world > 1;
  DevStatus* const st = rv.st;
  const bool retry = *(volatile int*)&st->overflow > rv.Kcap || (*(volatile int*)&st->violation != 0 && !rv.exactMode);
  if (retry || (!sharded && *(volatile int*)&st->nact == 0) || (sharded && *(volatile int*)&st->xok != 0)) return;   // cert stays CERT_NONE
  if (tid == 0) {
    if (sharded) {
      const unsigned seq = *cm.seq;                                   // the gather's exchange
      const volatile double* in = (volatile double*)cm.payload[cm.rank] + (size_t)(seq & 1) * CW * cm.width;
      int b = 0;
      for (int p = 0; p < cm.world; p++) {
        const volatile double* m = in + (size_t)p * cm.width;
        const int nT = (int)m[2];
        s_base[p] = b;
        // a quiet rank sends no records; its tile count is then unknown here, and irrelevant: zero records change nothing
        s_rec[p] = nT > 0 ? m + XHDR + (int)m[1] : nullptr;
        b += nT;
      }
      for (int p = cm.world; p <= CW; p++) s_base[p] = b;
    } else {
      s_base[0] = 0;
      for (int p = 1; p <= CW; p++) s_base[p] = v.T * CPT;
      s_rec[0] = rv.tileRec;
    }
  }
  __syncthreads();
  const int nRanks = sharded ? cm.world : 1;
  const int Ttot = s_base[nRanks];
  auto recOf = [&](int g) -> const volatile double* {
    int p = 0;
#pragma unroll
    for (int q = 1; q < CW; q++) if (q < nRanks && g >= s_base[q]) p = q;
    return s_rec[p] + (size_t)(g - s_base[p]) * XR;
  };
  // One record per thread and pass, loaded once (ld.global.cg: the mailbox is written by peers) and kept in registers.
  // The leader d of a record is whoever leads at its ENTRY: a certified record ends with everyone else 128 behind d, so
  // the next record's leader is d again — the certified suffix has one leader, and no pass over the totals is needed.
  // Likewise kappa only has to cover the operations up to the record it is applied to.
  __shared__ double s_w[TILE / 32][RP + 3];
  __shared__ long long s_key[TILE / 32];
  __shared__ int s_res[4];
  double carry[RP + 3];                           // running: E_r, rescaled sites, ops, sum of u*g
#pragma unroll
  for (int i = 0; i < RP + 3; i++) carry[i] = 0.0;
  long long best = -1;                            // (last uncertified record) << 32 | rescaled sites up to and including it
  for (int base = 0; base < Ttot; base += TILE) {
    const int g = base + tid;
    const bool in = g < Ttot;
    const double* rec = in ? const_cast<const double*>(recOf(g)) : nullptr;
    double x[RP + 3], Sr[RP], ab = 0.0;           // x: S_r, nAct, ops, (then) u*g — scanned inclusively
#pragma unroll
    for (int r = 0; r < RP; r++) { Sr[r] = in ? __ldcg(rec + 4 + r) : 0.0; x[r] = Sr[r]; }
    x[RP] = in ? __ldcg(rec + 0) : 0.0;
    x[RP + 1] = in ? __ldcg(rec + 1) : 0.0;
    if (in) ab = __ldcg(rec + 2);
    const double opsMine = x[RP + 1];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
#pragma unroll
      for (int i = 0; i < RP + 2; i++) { const double y = __shfl_up_sync(0xffffffffu, x[i], d); if (lane >= d) x[i] += y; }
    if (lane == 31) {
#pragma unroll
      for (int i = 0; i < RP + 2; i++) s_w[wid][i] = x[i];
    }
    __syncthreads();
    double tot[RP + 3];
#pragma unroll
    for (int i = 0; i < RP + 2; i++) {
      tot[i] = 0.0;
#pragma unroll
      for (int w = 0; w < TILE / 32; w++) { const double q = s_w[w][i]; if (w < wid) x[i] += q; tot[i] += q; }
      x[i] += carry[i];
    }
    // what enters this record, who leads there, and this record's share of the rounding bound
    double emax = -1e300, emin = 1e300, edom = 0.0;
    int dom = 0;
#pragma unroll
    for (int r = 0; r < RP; r++)
      if (r < R) {
        const double e = x[r] - Sr[r];
        if (e > emax) { emax = e; dom = r; edom = e; }
        emin = fmin(emin, e);
      }
    double mp[RP];                                // maxPre[dom][.]: issued now, needed after the second scan
#pragma unroll
    for (int r = 0; r < RP; r++) mp[r] = (in && r < R) ? __ldcg(rec + 4 + RP + dom * RP + r) : 0.0;
    if (in) rv.certDom[g] = dom;
    double ug = in ? 0x1p-22 * opsMine * ((emax - emin) + ab + 1.0) : 0.0;
    const double ugMine = ug;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const double y = __shfl_up_sync(0xffffffffu, ug, d); if (lane >= d) ug += y; }
    if (lane == 31) s_w[wid][RP + 2] = ug;
    __syncthreads();
    tot[RP + 2] = 0.0;
#pragma unroll
    for (int w = 0; w < TILE / 32; w++) { const double q = s_w[w][RP + 2]; if (w < wid) ug += q; tot[RP + 2] += q; }
    ug += carry[RP + 2];
    (void)ugMine;
    const double B = exp(x[RP + 1] * 0x1p-22) * (1.0 + 1e-9) * ug;
    bool bad = false;
    if (in) {
#pragma unroll
      for (int r = 0; r < RP; r++)
        if (r < R && r != dom) bad |= !((x[r] - Sr[r]) - edom + mp[r] + B < -128.0);
    }
    long long key = bad ? (((long long)g << 32) | (long long)(unsigned)x[RP]) : -1;
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) { const long long o = __shfl_xor_sync(0xffffffffu, key, q); key = o > key ? o : key; }
    if (lane == 0) s_key[wid] = key;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < TILE / 32; w++) best = s_key[w] > best ? s_key[w] : best;
#pragma unroll
    for (int i = 0; i < RP + 3; i++) carry[i] += tot[i];
    __syncthreads();
  }
  if (tid == 0) {
    const int tb0 = (int)(best >> 32), L00 = (int)(best & 0xffffffffLL);
    __threadfence_block();
    s_res[0] = tb0; s_res[1] = L00;
    s_res[2] = (tb0 + 1 < Ttot) ? *(volatile int*)&rv.certDom[tb0 + 1] : 0;  // the leader of the certified suffix
    s_res[3] = (rv.certOn && best >= 0 && L00 <= CERT_MAXLIT && tb0 < s_base[1]) ? 1 : 0;
  }
  __syncthreads();
  const int tb = s_res[0], L0 = s_res[1], dom = s_res[2];
  const bool ok = s_res[3] != 0;
  const int myBase = sharded ? s_base[cm.rank] : 0;
  if (tid == 0) {
    st->cert = ok ? CERT_OK : CERT_LITERAL;
    st->certRate = dom;
    st->certTile = tb - myBase;                   // local chunk index: every local chunk above it is certified
    st->certL0 = L0;
    if (ok) { st->chainSites = 0; st->chainRows = 0; }
  }
  if (!ok || (sharded && cm.rank != 0)) return;

  // the literal prefix: tiles 0 .. tb of this (first) block of sites, one tile per pass
  float c[RP];
#pragma unroll
  for (int r = 0; r < RP; r++) c[r] = 0.0f;
  int actBase = 0, rowsWalked = 0;
  const int litEnd = min(v.S, (tb + 1) * RCH);    // sites [0, litEnd) are walked
  for (int lt = 0; lt * TILE < litEnd; lt++) {
    const int s = lt * TILE + tid;
    const int km = s < litEnd ? min(rv.kmax[s], rv.Kcap) : 0;
    const int actv = km > 0;
    int ic = actv, ir = km;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int tc = __shfl_up_sync(0xffffffffu, ic, d), tr = __shfl_up_sync(0xffffffffu, ir, d);
      if (lane >= d) { ic += tc; ir += tr; }
    }
    if (lane == 31) { s_scan[0][wid] = ic; s_scan[1][wid] = ir; }
    __syncthreads();
    int nA = 0, nR = 0;
#pragma unroll
    for (int w = 0; w < TILE / 32; w++) { if (w < wid) { ic += s_scan[0][w]; ir += s_scan[1][w]; } nA += s_scan[0][w]; nR += s_scan[1][w]; }
    if (s < litEnd) rv.posIncl[s] = actBase + ic;
    const bool staged = nR <= CERT_ROWCAP;
    if (actv) {
      s_cnt[ic - 1] = km; s_site[ic - 1] = s;
      if (staged) {
        const float4* src = reinterpret_cast<const float4*>(rv.elist + (size_t)s * rv.Kcap * RP);
        for (int j = 0; j < km * V; j++) s_rows[(size_t)(ir - km) * V + j] = src[j];
      }
    }
    __syncthreads();
    if (tid == 0) {
      // the walker: the next site's row count and first row are fetched before the current site's dependent chain
      // (additions, maximum, subtraction) is worked through
      int q = 0;
      int cntN = nA > 0 ? s_cnt[0] : 0;
      float4 first[V];
#pragma unroll
      for (int u = 0; u < V; u++) first[u] = s_rows[u];
      for (int i = 0; i < nA; i++) {
        const int cnt = cntN;
        cntN = s_cnt[min(i + 1, TILE - 1)];
        if (staged) {                               // (two loops: the staged one compiles to shared-memory loads)
          for (int j = 0; j < cnt; j++) {
            float4 w[V];
#pragma unroll
            for (int u = 0; u < V; u++) w[u] = j == 0 ? first[u] : s_rows[(size_t)(q + j) * V + u];
#pragma unroll
            for (int r = 0; r < RP; r++)
              if (r < R) { const float4 x = w[r >> 2]; c[r] = __fadd_rn(c[r], (r & 3) == 0 ? x.x : (r & 3) == 1 ? x.y : (r & 3) == 2 ? x.z : x.w); }
          }
          const int qn = min(q + cnt, CERT_ROWCAP - 1);
#pragma unroll
          for (int u = 0; u < V; u++) first[u] = s_rows[(size_t)qn * V + u];
        } else {
          const float* rows = rv.elist + (size_t)s_site[i] * rv.Kcap * RP;
          for (int j = 0; j < cnt; j++)
#pragma unroll
            for (int r = 0; r < RP; r++) if (r < R) c[r] = __fadd_rn(c[r], rows[j * RP + r]);
        }
        q += cnt;
        float m = c[0];
#pragma unroll
        for (int r = 1; r < RP; r++) if (r < R) m = fmaxf(m, c[r]);
        float* co = rv.cout + (size_t)(actBase + i) * RP;
#pragma unroll
        for (int r = 0; r < RP; r++) if (r < R) { c[r] = __fsub_rn(c[r], m); co[r] = c[r]; }
        rv.mxo[actBase + i] = m;
      }
    }
    actBase += nA; rowsWalked += nR;
    __syncthreads();
  }
  if (tid == 0) { st->chainSites = actBase; st->chainRows = rowsWalked; }
}

// ------------------------------------------------------------------------------------------------
// K4a root_gather: per site tile, dot(root, limProbs) and the per-site list of non-zero rescale exponents in
// vertex-id order (the order AugTree.cpp:309-315 adds them in).  Only vertices whose tile flag is set are read:
// the flagged vertices of each 128-vertex window are compacted (in id order) into shared memory, then visited
// four at a time so that their exponent loads overlap.
// one thread, after every other write of the evaluation to *rv.st has been fenced: status block -> host mirror
__device__ __forceinline__ void publishStatus(const RootView& rv) {
  if (!rv.hostSt) return;
  const unsigned seq = rv.st->pubSeq + 1;
  rv.st->pubSeq = seq;
  const volatile int* src = reinterpret_cast<const volatile int*>(rv.st);
  volatile int* dst = reinterpret_cast<volatile int*>(rv.hostSt);
  constexpr int SEQ_WORD = offsetof(DevStatus, pubSeq) / 4;
#pragma unroll
  for (int i = 0; i < (int)(sizeof(DevStatus) / 4); i++) if (i != SEQ_WORD) dst[i] = src[i];
  __threadfence_system();                 // the payload is visible to the host before the sequence number is
  dst[SEQ_WORD] = (int)seq;
  rv.st->stamp[0] = ~0ull; rv.st->stamp[4] = ~0ull;   // "first block" stamps are atomicMin'ed: re-armed for the next evaluation
}

template <int RP, bool BATCH>
__global__ void __launch_bounds__(TILE) k_root_gather(DevView v, RootView rv0, int mode, PathView pv) {
  pdl_wait();
  const int pb = BATCH ? blockIdx.y : 0;           // proposal of a batch
  const RootView rv = BATCH ? forProposal(rv0, pb, v.Sp, v.T) : rv0;
  constexpr int NV = 16;                           // vertices per thread and pass of the (non-batch) flag scan
  __shared__ __align__(16) int s_list[TILE * NV];   // (also lent to certify() as its 16 KB row stage)
  __shared__ double s_red[TILE];
  __shared__ bool s_last;
  __shared__ int s_wcnt[TILE / 32];
  __shared__ double s_x[TILE / 32][RP + 3];
  __shared__ double s_y[TILE / 32][RP * RP];
  const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int site = tile * TILE + tid;
  const bool valid = site < v.S;
  const int R = v.R;
  const size_t Sp = (size_t)v.Sp;
  if (!BATCH && tid == 0) atomicMin(&rv.st->stamp[0], gtime());
  const int rootSlot = v.cur[0];
  double rd[RP];
#pragma unroll
  for (int r = 0; r < RP; r++)
    if (r < R) {
      if (BATCH) {
        rd[r] = pv.rootdot[((size_t)pb * R + r) * Sp + site];     // k_prune<., PATHS> computed it from the proposal's root
      } else {
        double a0, a1, a2, a3;
        ld256(v.partOf(rootSlot) + ((size_t)r * Sp + site) * 4, a0, a1, a2, a3);
        // arma::dot (op_dot::direct_dot_arma): two interleaved accumulators
        rd[r] = __dadd_rn(__dadd_rn(__dmul_rn(a0, c_pi[0]), __dmul_rn(a2, c_pi[2])), __dadd_rn(__dmul_rn(a1, c_pi[1]), __dmul_rn(a3, c_pi[3])));
        rv.rootdot[(size_t)r * Sp + site] = rd[r];
      }
    }
  int k[RP];
  double sd[RP];                                   // this site's exponent sums, one per rate (for the certificate only)
#pragma unroll
  for (int r = 0; r < RP; r++) { k[r] = 0; sd[r] = 0.0; }
  float* const myList = rv.elist + (size_t)site * rv.Kcap * RP;
  int* const myIds = BATCH ? nullptr : rv.eid + (size_t)site * rv.Kcap * RP;
  const size_t fstride = (size_t)v.T * v.nFlag;                 // between the flag planes of consecutive rates
  const uint8_t* f0 = v.flag[0] ? v.flag[0] + (size_t)tile * v.nFlag : nullptr;
  const uint8_t* f1 = v.flag[1] ? v.flag[1] + (size_t)tile * v.nFlag : nullptr;
  // the flagged vertices of a pass, compacted in vertex-id order into s_list, are visited four at a time so that their
  // exponent loads overlap
  auto consume = [&](int n) {
    for (int i = 0; i < n; i += 4) {
      float e[4][RP];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int ent = s_list[min(i + j, n - 1)];
        const float* src = v.expoOf((ent >> 8) & 1) + (size_t)(ent >> 9) * R * Sp + site;
#pragma unroll
        for (int r = 0; r < RP; r++) e[j][r] = (((ent >> r) & 1) && i + j < n) ? src[r * Sp] : 0.0f;   // unflagged rows were never written
      }
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int r = 0; r < RP; r++)
          if (e[j][r] != 0.0f) {
            if (k[r] < rv.Kcap && valid) {
              myList[(size_t)k[r] * RP + r] = e[j][r];
              if (!BATCH) myIds[(size_t)k[r] * RP + r] = s_list[min(i + j, n - 1)] >> 9;
            }
            k[r]++;
            sd[r] += (double)e[j][r];
          }
    }
  };
  if constexpr (!BATCH) {
    // 16 vertices per thread and pass: the slot bytes and BOTH slots' flag bytes of every rate come in as independent
    // 128-bit loads (one round trip per 4,096 vertices instead of two dependent ones per 256), the current slot's flags
    // are selected bytewise.  A pass without any flag (the usual case) costs nothing more.
    for (int base = 0; base < v.nReal; base += TILE * NV) {
      const int v0 = base + tid * NV;
      unsigned cw[4] = {0u, 0u, 0u, 0u}, any[4] = {0u, 0u, 0u, 0u}, sel[RP][4];
#pragma unroll
      for (int r = 0; r < RP; r++)
#pragma unroll
        for (int q = 0; q < 4; q++) sel[r][q] = 0u;
      if (v0 < v.nReal) {
        const uint4 c4 = *reinterpret_cast<const uint4*>(v.cur + v0);
        cw[0] = c4.x; cw[1] = c4.y; cw[2] = c4.z; cw[3] = c4.w;
        uint4 a4[RP], b4[RP];
#pragma unroll
        for (int r = 0; r < RP; r++)
          if (r < R) {
            a4[r] = *reinterpret_cast<const uint4*>(f0 + (size_t)r * fstride + v0);
            b4[r] = f1 ? *reinterpret_cast<const uint4*>(f1 + (size_t)r * fstride + v0) : make_uint4(0u, 0u, 0u, 0u);
          }
        const int nValid = min(NV, v.nReal - v0);      // the scratch row (and the padding) is not a vertex
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int nb = min(4, max(0, nValid - 4 * q));                    // valid bytes of this word
          const unsigned vm = nb >= 4 ? 0xffffffffu : ((1u << (8 * nb)) - 1u);
          cw[q] &= 0x01010101u & vm;
          const unsigned m = cw[q] * 0xffu;                                  // 0xff in the bytes whose vertex lives in slot 1
#pragma unroll
          for (int r = 0; r < RP; r++)
            if (r < R) {
              const unsigned a = q == 0 ? a4[r].x : q == 1 ? a4[r].y : q == 2 ? a4[r].z : a4[r].w;
              const unsigned b = q == 0 ? b4[r].x : q == 1 ? b4[r].y : q == 2 ? b4[r].z : b4[r].w;
              sel[r][q] = ((a & ~m) | (b & m)) & 0x01010101u & vm;
              any[q] |= sel[r][q];
            }
        }
      }
      const int mine = __popc(any[0]) + __popc(any[1]) + __popc(any[2]) + __popc(any[3]);
      int incl = mine;                              // inclusive scan over the block, in thread (= vertex) order
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int tq = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += tq; }
      if (lane == 31) s_wcnt[wid] = incl;
      __syncthreads();
      int off = incl - mine, n = 0;
#pragma unroll
      for (int w = 0; w < TILE / 32; w++) { if (w < wid) off += s_wcnt[w]; n += s_wcnt[w]; }
      if (mine) {
#pragma unroll
        for (int j = 0; j < NV; j++) {
          const int q = j >> 2, sh = 8 * (j & 3);
          if ((any[q] >> sh) & 1u) {
            int rbits = 0;
#pragma unroll
            for (int r = 0; r < RP; r++) rbits |= (int)((sel[r][q] >> sh) & 1u) << r;
            s_list[off++] = ((v0 + j) << 9) | ((int)((cw[q] >> sh) & 1u) << 8) | rbits;
          }
        }
      }
      __syncthreads();
      consume(n);
      __syncthreads();
    }
  } else {
    // Proposal batches: a proposal differs from the committed state only at the stored vertices of its path, so its
    // per-site lists are the COMMITTED lists (rv.base*: vertex-id order, ids alongside) with the path vertices' rows
    // dropped and their new exponents merged in by id — O(rows + path length) per site instead of a scan of every
    // vertex's flags.
    constexpr int PATH_CAP = 96;                   // = JCAP of dmpc_eval_proposals
    __shared__ int s_pid[PATH_CAP], s_ppos[PATH_CAP];
    __shared__ unsigned char s_pfl[PATH_CAP][RP];
    const int J = min(pv.JMAX, PATH_CAP);
    for (int j = tid; j < J; j += TILE) {
      const int id = pv.ids[(size_t)pb * pv.JMAX + j], ps = pv.pos[(size_t)pb * pv.JMAX + j];
      s_pid[j] = id; s_ppos[j] = ps;
#pragma unroll
      for (int r = 0; r < RP; r++)
        s_pfl[j][r] = (id != INT_MAX && r < R) ? pv.pf[(((size_t)pb * pv.JMAX + ps) * R + r) * v.T + tile] : 0;
    }
    __syncthreads();
    const float* const cl = rv.baseElist + (size_t)site * rv.Kcap * RP;
    const int* const ci = rv.baseEid + (size_t)site * rv.Kcap * RP;
    const int ckAll = rv.baseKmax[site], ck = min(ckAll, rv.Kcap);
    if (valid && ckAll > rv.Kcap) atomicMax(&rv.st->overflow, ckAll);     // truncated committed list: evaluate this one alone
#pragma unroll
    for (int r = 0; r < RP; r++) {
      if (r >= R) continue;
      int i = 0, j = 0;
      for (;;) {
        int idc = INT_MAX;
        float ec = 0.0f;
        while (i < ck) { ec = cl[(size_t)i * RP + r]; if (ec != 0.0f) { idc = ci[(size_t)i * RP + r]; break; } i++; }
        const int idp = j < J ? s_pid[j] : INT_MAX;
        if (idc == INT_MAX && idp == INT_MAX) break;
        float e;
        if (idp <= idc) {                          // a path vertex: its new exponent replaces whatever the committed list held
          if (idp == idc) i++;
          e = s_pfl[j][r] ? pv.pe[(((size_t)pb * pv.JMAX + s_ppos[j]) * R + r) * Sp + site] : 0.0f;
          j++;
        } else {
          e = ec; i++;
        }
        if (e != 0.0f) {
          if (k[r] < rv.Kcap && valid) myList[(size_t)k[r] * RP + r] = e;
          k[r]++;
          sd[r] += (double)e;
        }
      }
    }
  }
  int kmax = 0;
#pragma unroll
  for (int r = 0; r < RP; r++) kmax = max(kmax, k[r]);
  if (valid) {
    const int kk = min(kmax, rv.Kcap);
#pragma unroll
    for (int r = 0; r < RP; r++)
      for (int j = min(k[r], rv.Kcap); j < kk; j++) myList[(size_t)j * RP + r] = 0.0f;   // zero padding: x + 0.0f == x
    rv.kmax[site] = kmax;
    if (kmax > rv.Kcap) atomicMax(&rv.st->overflow, kmax);
  } else {
    rv.kmax[site] = 0;
  }
  const int anyActive = __syncthreads_or(valid && kmax > 0);
  // Optimistic finish: when no site of the whole alignment rescaled, the exponent vector never moves, every site's
  // term is log(mean_r(rootdot_r * expf(ev_r))) with the entry vector ev, and this kernel can finish the evaluation
  // by itself (same arithmetic and the same summation tree as k_final).  The host checks nact before trusting it.
  double ll = 0.0;
  if (valid) {
    double a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int r = 0; r < RP; r++)
      if (r < R) {
        const float ev = mode == 1 ? rv.carryIn[r] : 0.0f;
        const double w = (double)(float)exp((double)ev);
        const double x = __dmul_rn(rd[r], w);
        if (r & 1) a2 = __dadd_rn(a2, x); else a1 = __dadd_rn(a1, x);
      }
    ll = __dadd_rn(log(__dadd_rn(a1, a2) / (double)R), (double)0.0f);
    rv.sitell[site] = ll;
  }
  tile_chunk_sums(ll, s_red, rv.chunk + (size_t)tile * CPT);
  if (mode == 1) {
    // The records of the tile's four 64-site chunks for the chain certificate (certify): with E_r = running sums of the sites' exponent sums,
    //   S_r = the chunk's total, maxPre[d][r] = max over the chunk's sites (and 0) of (E_r - E_d) relative to the chunk's
    //   entry, nAct / ops / absSum = what bounds the fp32 rounding the literal chain would commit inside the chunk.
    constexpr int XR = xrec(RP);
    static_assert(RCH == 64 && TILE / 32 == 2 * CPT, "a chunk record is built from two warps");
    double* const rec = rv.tileRec + (size_t)tile * CPT * XR;
    if (!anyActive) {
      for (int i = tid; i < CPT * XR; i += TILE) rec[i] = 0.0;
    } else {
      const bool actS = valid && kmax > 0;
      double pre[RP], ab = 0.0;
#pragma unroll
      for (int r = 0; r < RP; r++) { pre[r] = (actS && r < R) ? sd[r] : 0.0; ab += fabs(pre[r]); }
      double ops = actS ? (double)(kmax + 1) : 0.0, na = actS ? 1.0 : 0.0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int r = 0; r < RP; r++) { const double y = __shfl_up_sync(0xffffffffu, pre[r], d); if (lane >= d) pre[r] += y; }
        ab += __shfl_xor_sync(0xffffffffu, ab, d); ops += __shfl_xor_sync(0xffffffffu, ops, d); na += __shfl_xor_sync(0xffffffffu, na, d);
      }
      if (lane == 31) {
#pragma unroll
        for (int r = 0; r < RP; r++) s_x[wid][r] = pre[r];
        s_x[wid][RP] = ab; s_x[wid][RP + 1] = ops; s_x[wid][RP + 2] = na;
      }
      __syncthreads();
      if (wid & 1) {                                 // prefix relative to the chunk's first site
#pragma unroll
        for (int r = 0; r < RP; r++) pre[r] += s_x[wid - 1][r];
      }
#pragma unroll
      for (int d = 0; d < RP; d++)
#pragma unroll
        for (int r = 0; r < RP; r++) {
          double mx = (d < R && r < R && d != r) ? fmax(0.0, pre[r] - pre[d]) : 0.0;
#pragma unroll
          for (int q = 16; q > 0; q >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, q));
          if (lane == 0) s_y[wid][d * RP + r] = mx;
        }
      __syncthreads();
      constexpr int PER = RP * RP + RP + 4;          // maxPre[RP][RP], S_r[RP], absSum, ops, nAct, pad
      for (int idx = tid; idx < CPT * PER; idx += TILE) {
        const int q = idx / PER, j = idx - q * PER;
        double* const rq = rec + (size_t)q * XR;
        if (j < RP * RP) rq[4 + RP + j] = fmax(s_y[2 * q][j], s_y[2 * q + 1][j]);
        else if (j == PER - 1) rq[3] = 0.0;
        else {
          const int jj = j - RP * RP;                // 0..RP-1: S_r; RP: absSum; RP+1: ops; RP+2: nAct
          const double acc = s_x[2 * q][jj] + s_x[2 * q + 1][jj];
          if (jj < RP) rq[4 + jj] = acc; else if (jj == RP) rq[2] = acc; else if (jj == RP + 1) rq[1] = acc; else rq[0] = acc;
        }
      }
    }
  }
  __syncthreads();                                 // the tile's records are complete
  if (tid == 0) {
    if (anyActive) atomicAdd(BATCH ? &rv.st->nact : rv.nactCtr, 1);
    __threadfence();
    s_last = atomicAdd(&rv.st->ticket2, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (!BATCH) {
      if (tid == 0) { rv.st->nact = *(volatile int*)rv.nactCtr; *rv.nactCtr = 0; }
      __syncthreads();
    }
    const bool quiet = *(volatile int*)&rv.st->nact == 0;
    const CommView& cm = rv.comm;
    const bool retry = *(volatile int*)&rv.st->overflow > rv.Kcap || (*(volatile int*)&rv.st->violation != 0 && !rv.exactMode);
    const bool pub = !BATCH && rv.publishAtGather;
    if (tid == 0) { rv.st->cert = CERT_NONE; if (!BATCH) rv.st->stamp[1] = gtime(); }
    if (cm.world > 1) {
      // skipped when the host is going to re-run this gather (list overflow, fused-clade violation), so that every
      // rank sends exactly one message per exchange
      if (retry) {
        if (tid == 0) { rv.st->xok = -1; rv.st->ticket2 = 0; __threadfence(); if (pub) publishStatus(rv); }
        return;
      }
      const int res = exchangeAndReduce(rv, rv.nChunks, quiet, s_red, (mode == 1 && !quiet) ? rv.tileRec : nullptr, gridDim.x * CPT, xrec(RP));
      if (tid == 0) { rv.st->xok = res; if (!BATCH) rv.st->stamp[2] = gtime(); }
      __syncthreads();
      if (res == 0 && mode == 1) certify<RP>(v, rv, reinterpret_cast<float4*>(s_list));
      __syncthreads();
      if (tid == 0) { rv.st->ticket2 = 0; if (!BATCH) rv.st->stamp[3] = gtime(); __threadfence(); if (pub) publishStatus(rv); }
      return;
    }
    if (quiet) {
      const double tot = reduce_chunks_device(rv.chunk, rv.nChunks, s_red);
      if (tid == 0) rv.st->ll = tot;
    }
    if (tid == 0 && !BATCH) rv.st->stamp[2] = gtime();
    if (!quiet && mode == 1 && !retry) {
      certify<RP>(v, rv, reinterpret_cast<float4*>(s_list));
      __syncthreads();
    }
    if (tid == 0) { rv.st->ticket2 = 0; if (!BATCH) rv.st->stamp[3] = gtime(); __threadfence(); if (pub) publishStatus(rv); }
  }
}

// ------------------------------------------------------------------------------------------------
// K4b chain: AugTree.cpp:306-318 keeps ONE fp32 exponent vector across all loci (it is never reset), so the
// per-site values depend on the site order through a sequential chain of fp32 additions whose rounding has to
// be reproduced bit for bit.  One block: every thread compacts the active sites (those
// that rescaled anywhere), then thread 0 walks their exponent rows as ONE flat stream while the other warps
// stage the next chunk of rows (and the end-of-site markers) in shared memory.  RC = number of rate categories
// (compile time), RP = floats per row (4 or 8).  When no site rescaled the kernel returns at once.  In a
// site-sharded run the vector entering this shard comes from an earlier rank's mailbox and the vector leaving it is
// handed to the later ranks (CommView).
// G = rows per group (4 for long lists, 1 when most rescaled sites hold a single row): a site's rows are padded to a
// multiple of G with zero rows, the walker tests for the end of a site once per group.
template <int RP, int RC, int G, bool BATCH>
__global__ void __launch_bounds__(1024) k_chain(DevView v, RootView rv0) {
  const RootView rv = BATCH ? forProposal(rv0, blockIdx.x, v.Sp, v.T) : rv0;   // BATCH: one block per proposal
  extern __shared__ float4 s_dyn[];
  constexpr int V = RP / 4;                       // float4 per row
  __shared__ int s_scan[2][32];
  __shared__ int s_total[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nth = blockDim.x;
  const int S = v.S, Q = rv.Q;
  const CommView& cm = rv.comm;
  const bool sharded = cm.world > 1;
  if (sharded) {
    // the chain crosses the site blocks in rank order: take the vector from the nearest earlier rank that holds a
    // rescaled site (every rank learnt who does from the gather's exchange), or zeros when there is none
    __shared__ int s_prev;
    const unsigned seq = *cm.seq;
    const int buf = seq & 1;
    if (tid == 0) {
      const volatile double* in = (volatile double*)cm.payload[cm.rank] + (size_t)buf * CW * cm.width;
      int prev = -1;
      for (int p = 0; p < cm.rank; p++) if (in[(size_t)p * cm.width] != 0.0) prev = p;
      if (prev >= 0) {
        const volatile unsigned* f = (volatile unsigned*)cm.cflag[cm.rank] + prev;
        const long long t0 = clock64();
        while ((int)(*f - seq) < 0) { if (clock64() - t0 > 40000000000LL) { prev = -2; break; } __nanosleep(200); }
      }
      s_prev = prev;
    }
    __syncthreads();
    __threadfence_system();
    if (tid < RMAX) {
      float x = 0.0f;
      if (s_prev >= 0) x = *((volatile float*)cm.carry[cm.rank] + ((size_t)buf * CW + s_prev) * RMAX + tid);
      rv.carryIn[tid] = x;
    }
    if (tid == 0 && s_prev == -2) rv.st->xok = 2;
    __syncthreads();
  }
  if (rv.st->nact == 0) {                         // nothing rescaled here: the vector passes through unchanged
    if (tid < RMAX) rv.carryOut[tid] = rv.carryIn[tid];
    return;
  }

  // phase A: ordered compaction of the sites that rescaled anywhere (kmax > 0) + prefix of their row counts.
  // A site's rows are padded to a multiple of G with zero rows (x + 0.0f == x), so that the walker consumes
  // whole groups and a site can only end at a group boundary.
  int baseCnt = 0, baseRows = 0;
  for (int start = 0; start < S; start += nth) {
    const int s = start + tid;
    const int km = s < S ? (min(rv.kmax[s], rv.Kcap) + G - 1) / G * G : 0;
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
  if (tid == 0) { rv.rowoff[nact] = baseRows; rv.st->chainSites = nact; rv.st->chainRows = baseRows; }
  __syncthreads();
  // staging chunks: chunk c = active sites whose row offset lies in [cQ, (c+1)Q); Q >= padded Kcap so no chunk is empty
  const int nchunks = nact ? rv.rowoff[nact - 1] / Q + 1 : 0;
  for (int i = tid; i < nact; i += nth) {
    const int c = rv.rowoff[i] / Q;
    if (i == 0 || rv.rowoff[i - 1] / Q != c) rv.chunkStart[c] = i;
  }
  if (tid == 0) rv.chunkStart[nchunks] = nact;
  __syncthreads();

  // phase B.  Shared memory: 2 buffers of cap rows (+ one spare group the walker may prefetch), then one
  // end-of-site marker per group of 4 rows, then per-buffer {first active index, rows}
  const int cap = 2 * Q + 8;
  float4* const buf0 = s_dyn;
  unsigned char* const end0 = reinterpret_cast<unsigned char*>(s_dyn + (size_t)2 * cap * V);   // one marker per group
  int* const meta0 = reinterpret_cast<int*>(end0 + 2 * cap);
  float c[RC];
#pragma unroll
  for (int r = 0; r < RC; r++) c[r] = rv.carryIn[r];

  auto stage = [&](int ch, int ltid, int nld, int b) {
    float4* const bufb = buf0 + (size_t)b * cap * V;
    unsigned char* const endb = end0 + (size_t)b * cap;
    const int i0 = rv.chunkStart[ch], i1 = rv.chunkStart[ch + 1];
    const int r0 = rv.rowoff[i0], nrows = rv.rowoff[i1] - r0, ns = i1 - i0;
    if (ltid == 0) { meta0[2 * b] = i0; meta0[2 * b + 1] = nrows; }
    for (int q = ltid; q < nrows + G; q += nld) {
      if (q >= nrows) {                           // the spare group: zeros, never an end
#pragma unroll
        for (int u = 0; u < V; u++) bufb[(size_t)q * V + u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q % G == 0) endb[q / G] = 0;
        continue;
      }
      int lo = 0, hi = ns;                        // last k with rowoff[i0 + k] - r0 <= q
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (rv.rowoff[i0 + mid] - r0 <= q) lo = mid; else hi = mid; }
      const int site = rv.act[i0 + lo], j = q - (rv.rowoff[i0 + lo] - r0);
      const int km = min(rv.kmax[site], rv.Kcap);
      const float4* src = reinterpret_cast<const float4*>(rv.elist + ((size_t)site * rv.Kcap + j) * RP);
#pragma unroll
      for (int u = 0; u < V; u++) bufb[(size_t)q * V + u] = j < km ? src[u] : make_float4(0.f, 0.f, 0.f, 0.f);
      if (q % G == 0) endb[q / G] = (unsigned char)(q + G == rv.rowoff[i0 + lo + 1] - r0);
    }
  };

  if (nchunks > 0) stage(0, tid, nth, 0);
  __syncthreads();
  for (int ch = 0; ch < nchunks; ch++) {
    const int b = ch & 1;
    if (wid > 0) {
      if (ch + 1 < nchunks) stage(ch + 1, tid - 32, nth - 32, b ^ 1);
    } else if (tid == 0) {
      // The walker: ONE thread, because the chain is sequential by definition.  A lone warp retires an instruction
      // every few cycles, so the loop is kept to the loads (explicit shared-space addresses, one running pointer),
      // the RC additions per row and one flag test per group of 4 rows; two groups per trip so that the next group
      // is already on its way from shared memory while the current one is added.
      unsigned rowAddr = (unsigned)__cvta_generic_to_shared(buf0 + (size_t)b * cap * V);
      unsigned endAddr = (unsigned)__cvta_generic_to_shared(end0 + (size_t)b * cap);
      int outIdx = meta0[2 * b];
      const int ngroups = meta0[2 * b + 1] / G;
      float* co = rv.cout + (size_t)outIdx * RP;
      float* mo = rv.mxo + outIdx;
      float4 x[G][V], y[G][V];
      unsigned fx, fy;
      auto loadGroup = [&](float4 (&g)[G][V], unsigned& f, unsigned ra, unsigned ea) {
#pragma unroll
        for (int j = 0; j < G; j++)
#pragma unroll
          for (int u = 0; u < V; u++)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(g[j][u].x), "=f"(g[j][u].y), "=f"(g[j][u].z), "=f"(g[j][u].w)
                         : "r"(ra + (unsigned)((j * V + u) * 16)));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(f) : "r"(ea));
      };
      auto addGroup = [&](const float4 (&g)[G][V], unsigned flag) {
#pragma unroll
        for (int j = 0; j < G; j++)
#pragma unroll
          for (int r = 0; r < RC; r++) {
            const float4 w = g[j][r >> 2];
            c[r] = __fadd_rn(c[r], (r & 3) == 0 ? w.x : (r & 3) == 1 ? w.y : (r & 3) == 2 ? w.z : w.w);
          }
        if (flag) {                               // last group of a site: maxExponent, exponentVec -= maxExponent
          float m = c[0];
#pragma unroll
          for (int r = 1; r < RC; r++) m = fmaxf(m, c[r]);
#pragma unroll
          for (int r = 0; r < RC; r++) c[r] = __fsub_rn(c[r], m);
#pragma unroll
          for (int r = 0; r < RC; r++) co[r] = c[r];
          *mo = m;
          co += RP; mo++;
        }
      };
      constexpr unsigned GB = G * V * 16;         // bytes of one group of rows
      if (ngroups > 0) loadGroup(x, fx, rowAddr, endAddr);
      for (int g = 0; g < ngroups; g += 2) {
        loadGroup(y, fy, rowAddr + GB, endAddr + 1);      // the spare group behind the last one makes this safe
        addGroup(x, fx);
        if (g + 1 >= ngroups) break;
        rowAddr += 2 * GB; endAddr += 2;
        loadGroup(x, fx, rowAddr, endAddr);
        addGroup(y, fy);
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
#pragma unroll
    for (int r = 0; r < RMAX; r++) if (r < RC) rv.carryOut[r] = c[r];
    if (sharded) {                                  // hand the vector to every later rank
      const unsigned seq = *cm.seq;
      const int buf = seq & 1;
      for (int p = cm.rank + 1; p < cm.world; p++) {
#pragma unroll
        for (int r = 0; r < RMAX; r++)
          *((volatile float*)cm.carry[p] + ((size_t)buf * CW + cm.rank) * RMAX + r) = r < RC ? c[r] : 0.0f;
      }
      __threadfence_system();
      for (int p = cm.rank + 1; p < cm.world; p++) *((volatile unsigned*)cm.cflag[p] + cm.rank) = seq;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4c final: rateAveragedLogLiks (AugTree.cpp:317-323) per site, then a deterministic sum: a fixed tree inside each
// RCH-site chunk, chunk sums combined in reduce_chunks order (identical on the host for the multi-GPU path).
// mode 0: each site folds its own list from zero (numRateCats == 1, where the carry is always exactly zero, or the
//         textbook reduction);
// mode 1: exponent vectors come from the literal chain (k_chain);
// mode 2: queued behind k_root_gather (whose last block certifies) without a host round trip: nothing to do unless the verdict is CERT_OK;
//         then the tiles of the literal prefix read certify()'s vectors (as in mode 1) and every later tile takes the closed
//         form of the certified regime (see certify).
// publish: the last thread standing stores the status block into the host mirror (the host polls for it).
template <bool BATCH>
__global__ void __launch_bounds__(TILE) k_final(DevView v, RootView rv0, int mode, int reduceAll, int publish, PathView pv) {
  pdl_wait();
  const int pb = BATCH ? blockIdx.y : 0;
  const RootView rv = BATCH ? forProposal(rv0, pb, v.Sp, v.T) : rv0;
  const double* const rootdot = BATCH ? pv.rootdot + (size_t)pb * v.R * v.Sp : rv.rootdot;
  __shared__ double s_red[TILE];
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const int s = blockIdx.x * TILE + tid;
  int dom = -1;                                      // >= 0: this tile takes the closed form with that category
  if (mode == 2) {
    const bool retry = rv.st->overflow > rv.Kcap || (rv.st->violation != 0 && !rv.exactMode);
    if (retry || rv.st->cert != CERT_OK) {
      if (publish && !BATCH && blockIdx.x == 0 && tid == 0) publishStatus(rv);
      return;
    }
    if (s / RCH > rv.st->certTile) dom = rv.st->certRate;       // (per site: the prefix may end inside this tile)
  }
  if (!BATCH && tid == 0) atomicMin(&rv.st->stamp[4], gtime());
  double ll = 0.0;
  if (s < v.S) {
    float ev[RMAX];
    float mx = 0.0f;
    const int km = min(rv.kmax[s], rv.Kcap);
    double w[RMAX];
    if (dom >= 0) {
      // certified: expf(c_r) == 0 for r != dom, c_dom == 0 when the site begins, so its maximum is the fp32 sum of its own
      // category-dom exponents (zero padding rows add +0.0f)
      for (int j = 0; j < km; j++) mx = __fadd_rn(mx, rv.elist[((size_t)s * rv.Kcap + j) * rv.RP + dom]);
#pragma unroll
      for (int r = 0; r < RMAX; r++) w[r] = r == dom ? 1.0 : 0.0;
    } else {
      if (mode >= 1 && rv.st->nact == 0) {          // the chain kernel returned early: no site rescaled
#pragma unroll
        for (int r = 0; r < RMAX; r++) if (r < v.R) ev[r] = rv.carryIn[r];
      } else if (mode >= 1) {
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
#pragma unroll
      for (int r = 0; r < RMAX; r++) if (r < v.R) w[r] = (double)(float)exp((double)ev[r]);   // correctly rounded expf
    }
    // mean(likProp % exp(ev)): expf on the fp32 vector, promoted to double; two interleaved accumulators
    double a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int r = 0; r < RMAX; r++)
      if (r < v.R) {
        const double x = __dmul_rn(rootdot[(size_t)r * v.Sp + s], w[r]);
        if (r & 1) a2 = __dadd_rn(a2, x); else a1 = __dadd_rn(a1, x);
      }
    ll = __dadd_rn(log(__dadd_rn(a1, a2) / (double)v.R), (double)mx);
    rv.sitell[s] = ll;
  }
  tile_chunk_sums(ll, s_red, rv.chunk + (size_t)blockIdx.x * CPT);
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(&rv.st->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (!BATCH && tid == 0) rv.st->stamp[5] = gtime();
  if (reduceAll == 2) {                             // site-sharded: second exchange, every shard is final now
    const int res = exchangeAndReduce(rv, rv.nChunks, true, s_red, nullptr, 0, 0);
    if (tid == 0) rv.st->xok = res;
  } else if (reduceAll) {
    const double tot = reduce_chunks_device(rv.chunk, rv.nChunks, s_red);
    if (tid == 0) rv.st->ll = tot;
  }
  if (tid == 0) {
    rv.st->ticket = 0;
    if (!BATCH) rv.st->stamp[6] = gtime();
    __threadfence();
    if (publish && !BATCH) publishStatus(rv);
  }
}

}  // namespace dmpc
