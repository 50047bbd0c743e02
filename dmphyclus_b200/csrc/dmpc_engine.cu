// dmpc_engine.cu — C-ABI implementation (include/dmpc.h) of the B200-native likelihood engine.
//
// Each entry point mirrors one Rcpp-exported function of the reference
// (/root/reference/src/clusterFunArmadillo.cpp:36-332): same call sequence
// (NegateAllUpdateFlags -> mutate topology/flags -> ComputeLoglik), but the state lives in flat host arrays
// (host_tree.hpp) plus dense HBM buffers, and ComputeLoglik is a level plan of CUDA launches
// (kernels.cuh).  There is no CPU arithmetic path: without a usable device every call fails.
#include "../../include/dmpc.h"
#include "../../include/dmpc_test.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <thread>
#include <utility>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "host_tree.hpp"
#include "kernels.cuh"

using namespace dmpc;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(DMPC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
  } while (0)

typedef dmpc::DevStatus HostResult;   // pinned mirror of the device status block
// a plan never holds more than one task per internal vertex, two tokens and two tip-list entries per fused vertex
// ... and at most one segment per task (+ the closing entry of the segment table: PLAN_SLACK)
// ... and one pending rollback slot bit per vertex
constexpr size_t PLAN_BYTES = sizeof(Task) + 2 * sizeof(Token) + 2 * sizeof(int32_t) + sizeof(int32_t) + sizeof(int32_t);
constexpr size_t PLAN_SLACK = 64;

// Per-device cache of the host-side resources a handle needs (stream, events, pinned staging buffers).  Creating
// and destroying them costs milliseconds (cudaHostAlloc/cudaFreeHost synchronise the device), which would dwarf
// a logLikFromClusInd-style create/evaluate/destroy cycle; handles borrow a pack and give it back.
constexpr int MAX_SLABS = 16;
struct ResPack {
  cudaStream_t stream = nullptr;
  cudaStream_t copyStream = nullptr;   // the alignment upload of logLikCpp runs here, slab by slab, ahead of the pruning
  cudaEvent_t slabEv[MAX_SLABS + 1] = {};   // [MAX_SLABS]: the plan upload of a slabbed create
  cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  char* h_plan = nullptr;        // pinned: tasks | tokens | tip list of one evaluation, packed, uploaded in ONE copy
  int32_t* h_curList = nullptr;  // pinned
  HostResult* h_res = nullptr;   // pinned + mapped: k_root_gather publishes the status block straight into it
  HostResult* d_res = nullptr;   // ... its device-side address
  double* h_chunks = nullptr;    // pinned
  size_t cap = 0;                // internal vertices the pinned buffers can hold
  size_t chunkCap = 0;
};
struct DeviceCtx {
  bool checked = false;
  std::vector<ResPack> freePacks;
};
std::mutex g_ctxMutex;
DeviceCtx g_ctx[64];

}  // namespace

struct dmpc_comm {
  int device = 0, rank = 0, world = 1, width = 0;
  void* local = nullptr;               // this rank's mailbox (cudaMalloc: IPC-shareable)
  void* peer[CW] = {nullptr};          // every rank's mailbox as seen from here
  unsigned* seq = nullptr;             // message counter (device)
  bool connected = false;
  bool ipc = true;                     // peers were opened through CUDA IPC handles (false: shards of one process)
  int maxChunks = 0;
  size_t flagBytes() const { return 64; }
  size_t payloadBytes() const { return (size_t)2 * CW * width * sizeof(double); }
  size_t carryBytes() const { return (size_t)2 * CW * RMAX * sizeof(float); }
  size_t bytes() const { return 2 * flagBytes() + payloadBytes() + carryBytes(); }   // flags, payload, chain flags, carry
};

struct Group;
struct dmpc_tree {
  Group* grp = nullptr;          // set on the LEADER of a single-process multi-device handle: entry points fan out to its shards
  bool isShard = false;          // a shard of such a handle: the leader holds the device locks
  HostTree ht;
  int N = 0, S = 0, Sp = 0, R = 0, T = 0, nInt = 0, nReal = 0, nChunks = 0, nFlag = 0;
  bool quirk = true, deferred = false, alive = true;
  int device = 0, siteBegin = 0, sitesTotal = 0;
  ResPack pack;
  cudaStream_t stream = nullptr;                       // = pack.stream
  cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  DevView dv{};
  RootView rv{};
  uint8_t* d_tipmask = nullptr;
  double* d_lut = nullptr;
  char* d_plan = nullptr;        // tasks | tokens | tip list of the current evaluation (PLAN_BYTES per internal vertex)
  char* h_plan = nullptr;        // pinned
  int32_t* d_curList = nullptr;
  int32_t* h_curList = nullptr;  // pinned
  HostResult* h_res = nullptr;   // pinned
  double hP[2][RMAX][16];
  double hPi[4];
  double wCol[RMAX * 16], bCol[RMAX * 16];   // the matrix lists as they last came in (column-major), for re-uploads
  bool constantsValid = false;
  std::vector<void*> allocs;
  dmpc_stats st{};
  HostTree::Plan plan;
  // what runPlan launches: the plan just built (d_tasks ...) or the cached full-evaluation plan replayed with a slot flip
  const HostTree::Plan* runPlanMeta = nullptr;
  const Task* runTasks = nullptr; const Token* runTokens = nullptr; const int32_t* runTips = nullptr; const int32_t* runSegs = nullptr;
  const int32_t* runCur = nullptr; int runCurN = 0;   // pending rollback slot bits riding with the plan (first launch applies them)
  int runFlip = 0;
  struct {
    bool valid = false;
    int fuse = 0;
    PooledVec<int32_t> kids;             // (recycled storage: a create / destroy cycle re-uses the blocks, host_tree.hpp)
    PooledVec<uint8_t> within, cur;
    HostTree::Plan meta;                 // level structure and counters (the big vectors stay empty)
    char* dPlan = nullptr;               // device copy of the packed plan
    size_t offTok = 0, offTip = 0, offSeg = 0, bytes = 0;
    int metaTasks = 0;
  } cache;
  unsigned long long planCacheHits = 0;
  // the cached full-evaluation plan as CUDA graphs (one per slot flip): level launches + gather in ONE submission
  struct { cudaGraphExec_t exec = nullptr; unsigned long long version = 0; bool predicted = false; } graphs[2];
  unsigned long long graphVersion = 1;   // bumped whenever something baked into the captured kernel parameters changes
  bool capturing = false, gatherPrelaunched = false;
  unsigned long long graphReplays = 0;
  bool exact = false;            // fused clades evaluated with per-vertex rescale bookkeeping (see k_prune_fused)
  size_t planOffTok = 0, planOffTip = 0, planOffSeg = 0, planBytes = 0;   // layout of the plan last packed into d_plan
  unsigned pubExpected = 0;      // status publications k_root_gather has been asked for so far (see DevStatus::pubSeq)
  bool pdl = true;               // launch the kernels of an evaluation as programmatic dependents of one another (DMPC_PDL=0: off)
  bool timingOn = false;         // dmpc_get_stats has been called: keep the per-evaluation CUDA events of proposals too
  bool timeThis = true;          // ... whether the evaluation in flight records them (full evaluations always do)
  bool evalTimePending = false;  // last_eval_ms still to be read from the events (done lazily: the result is polled)
  // logLikCpp only: the alignment is uploaded in slabs of site tiles on the copy stream while the slabs already on the
  // device are pruned (one-shot; the first evaluation consumes it)
  struct {
    int n = 0; const uint8_t* src = nullptr; size_t pitch = 0; uint8_t* stage = nullptr;
    int tile[MAX_SLABS + 2] = {0};           // slab k = site tiles [tile[k], tile[k + 1])
    // Equal slabs.  Tried: sizes shrinking by 0.75 from slab to slab, so that the pruning left when the last byte has
    // landed is that of a small slab — measured WORSE on 10,000 x 30,000 (6.96 against 6.44 ms end to end): the plan
    // queues behind a slab 0 of 1.55 ms instead of 0.7, and the small late slabs prune at a fraction of the full rate.
    void cut(int T, int nSlab) {
      n = nSlab;
      const double q = 1.0;
      double sum = 0.0, w = 1.0, acc = 0.0;
      for (int k = 0; k < n; k++, w *= q) sum += w;
      w = 1.0;
      for (int k = 0; k < n; k++, w *= q) {
        tile[k] = k == 0 ? 0 : std::min(T - (n - k), std::max(tile[k - 1] + 1, (int)std::lround(T * acc / sum)));
        acc += w;
      }
      tile[n] = T;
    }
  } slabs;
  bool planOnCopyStream = false;   // the next uploadPlan goes through the copy stream (slabbed create: see createImpl)
  double* h_chunks = nullptr;    // pinned: chunk sums of the last deferred evaluation
  bool chunksOnHost = false;     // ... and whether they are final (no rescaled site, zero entry vector)
  bool carryIsZero = true;
  int pendingCur = 0;            // slot bits of a rollback that the device has not seen yet (h_curList)
  dmpc_comm* comm = nullptr;
  int chainGroup = 4;            // rows per group of the chain walker: 4 for long exponent lists, 1 for sparse ones
  int lastNact = 0;
  bool predictActive = false;    // the last evaluation saw a rescaled site: queue gather + cert + final without waiting for the gather
  bool seqPredicted = false;     // what the root sequence in flight was launched with
  int lastCert = 0, lastCertRate = -1, lastCertSites = 0;
  std::vector<int> lastTouched;  // internal indices the pending proposal flipped

  template <class Tp>
  int alloc(Tp** p, size_t n) {
    void* q = nullptr;
    CUDA_TRY(cudaMallocAsync(&q, n * sizeof(Tp), stream));
    allocs.push_back(q);
    st.device_bytes += n * sizeof(Tp);
    *p = (Tp*)q;
    return DMPC_OK;
  }
};

namespace {

int acquirePack(dmpc_tree* t) {
  DeviceCtx& cx = g_ctx[t->device & 63];
  {
    std::lock_guard<std::mutex> lk(g_ctxMutex);
    if (!cx.checked) {
      int major = 0;
      CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, t->device));
      if (major < 10) return fail(DMPC_ERR_CUDA, "device is not sm_100-class; kernels are built for sm_100a only");
      cudaMemPool_t pool;
      CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, t->device));
      unsigned long long keep = ~0ULL;                    // keep freed HBM cached in the pool between handles
      CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
      CUDA_TRY(cudaFuncSetAttribute(k_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, WALK_SMEM));
      cx.checked = true;
    }
    if (!cx.freePacks.empty()) { t->pack = cx.freePacks.back(); cx.freePacks.pop_back(); }
  }
  if (t->pack.h_res) std::memset(t->pack.h_res, 0, sizeof(HostResult));   // an idle pack: nothing is in flight; pubSeq restarts at 0
  ResPack& p = t->pack;
  if (!p.stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&p.copyStream, cudaStreamNonBlocking));
    for (auto& e : p.ev) CUDA_TRY(cudaEventCreate(&e));
    for (auto& e : p.slabEv) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_TRY(cudaHostAlloc((void**)&p.h_res, sizeof(HostResult), cudaHostAllocMapped));
    std::memset(p.h_res, 0, sizeof(HostResult));
    CUDA_TRY(cudaHostGetDevicePointer((void**)&p.d_res, p.h_res, 0));
  }
  if (p.cap < (size_t)t->nInt) {
    if (p.h_plan) cudaFreeHost(p.h_plan);
    if (p.h_curList) cudaFreeHost(p.h_curList);
    p.h_plan = nullptr; p.h_curList = nullptr; p.cap = 0;
    const size_t cap = std::max<size_t>(1024, (size_t)t->nInt * 5 / 4);
    CUDA_TRY(cudaHostAlloc((void**)&p.h_plan, cap * PLAN_BYTES + PLAN_SLACK, cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc((void**)&p.h_curList, cap * sizeof(int32_t), cudaHostAllocDefault));
    p.cap = cap;
  }
  if (p.chunkCap < (size_t)t->nChunks) {
    if (p.h_chunks) cudaFreeHost(p.h_chunks);
    p.h_chunks = nullptr; p.chunkCap = 0;
    const size_t cc = std::max<size_t>(256, (size_t)t->nChunks * 5 / 4);
    CUDA_TRY(cudaHostAlloc((void**)&p.h_chunks, cc * sizeof(double), cudaHostAllocDefault));
    p.chunkCap = cc;
  }
  t->h_chunks = p.h_chunks;
  t->stream = p.stream;
  for (int i = 0; i < 7; i++) t->ev[i] = p.ev[i];
  t->h_plan = p.h_plan; t->h_curList = p.h_curList; t->h_res = p.h_res;
  t->rv.hostSt = p.d_res;
  return DMPC_OK;
}

void releasePack(dmpc_tree* t) {
  if (!t->pack.stream) return;
  std::lock_guard<std::mutex> lk(g_ctxMutex);
  g_ctx[t->device & 63].freePacks.push_back(t->pack);
  t->pack = ResPack{};
  t->stream = nullptr;
}

// RestoreIterVecAndExp (IntermediateNode.h:26-33): the vertices the last proposal touched go back to their previous slot.
// The host state flips at once; the device copy of the slot bits (only the root reduction reads it) is brought up to date
// by the next call that launches kernels, so that a reject costs no launch and no synchronisation of its own.
int flushCur(dmpc_tree* t) {
  if (!t->pendingCur) return DMPC_OK;
  const int n = t->pendingCur;
  t->pendingCur = 0;
  CUDA_TRY(cudaMemcpyAsync(t->d_curList, t->h_curList, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, t->stream));
  k_set_cur<<<(n + 255) / 256, 256, 0, t->stream>>>(t->dv.cur, t->d_curList, n);
  t->st.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  return DMPC_OK;
}

int ensureDevice(dmpc_tree* t) {
  CUDA_TRY(cudaSetDevice(t->device));
  return DMPC_OK;
}

int allocSlot(dmpc_tree* t, int slot) {
  if (t->dv.part[slot]) return DMPC_OK;
  int rc;
  t->graphVersion++;
  const size_t n = (size_t)t->nInt * t->R * t->Sp;
  if ((rc = t->alloc(&t->dv.part[slot], n * 4))) return rc;
  if ((rc = t->alloc(&t->dv.expo[slot], n))) return rc;
  if ((rc = t->alloc(&t->dv.flag[slot], (size_t)t->R * t->T * t->nFlag))) return rc;
  return DMPC_OK;
}

void setElistGeometry(dmpc_tree* t, int kcap) {
  t->rv.Kcap = kcap;
  t->graphVersion++;
  t->rv.Q = std::max(512, (kcap + 3) & ~3);       // >= the padded length of any site's list, multiple of both group sizes
  t->st.exponent_capacity = kcap;
}

int allocElist(dmpc_tree* t, int kcap) {
  // the previous list stays allocated until the handle dies; growth is rare and geometric
  setElistGeometry(t, kcap);
  int rc = t->alloc(&t->rv.elist, (size_t)t->Sp * kcap * t->rv.RP);
  if (rc) return rc;
  return t->alloc(&t->rv.eid, (size_t)t->Sp * kcap * t->rv.RP);
}

// matrices arrive column-major from R (P(i,j) at [i + 4j]); constants hold them row-major
int uploadConstants(dmpc_tree* t, const double* within, const double* between, const double* pi) {
  double P[2][RMAX][16];
  std::memset(P, 0, sizeof P);
  for (int r = 0; r < t->R; r++)
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        P[0][r][i * 4 + j] = between[16 * r + i + 4 * j];
        P[1][r][i * 4 + j] = within[16 * r + i + 4 * j];
      }
  if (t->constantsValid && !std::memcmp(P, t->hP, sizeof P) && !std::memcmp(pi, t->hPi, sizeof t->hPi)) return DMPC_OK;
  // the fast fused path tests only the product of the two child terms for a possible rescale, which is sound when no
  // term can exceed 1: entries in [0, 1] and row sums <= 1 (up to rounding).  Anything else -> exact mode for good.
  for (int set = 0; set < 2 && !t->exact; set++)
    for (int r = 0; r < t->R && !t->exact; r++)
      for (int i = 0; i < 4; i++) {
        double row = 0.0;
        for (int j = 0; j < 4; j++) { const double x = P[set][r][i * 4 + j]; if (!(x >= 0.0 && x <= 1.0)) t->exact = true; row += x; }
        if (!(row <= 1.0 + 1e-9)) t->exact = true;
      }
  std::memcpy(t->hP, P, sizeof P);
  std::memmove(t->hPi, pi, sizeof t->hPi);
  std::memmove(t->wCol, within, (size_t)t->R * 16 * sizeof(double));
  std::memmove(t->bCol, between, (size_t)t->R * 16 * sizeof(double));
  // tip lookup: P * indicator(mask), summed in gemv_emul_tinysq order (adding +0.0 terms is exact)
  std::vector<double> lut((size_t)2 * t->R * 64);
  for (int set = 0; set < 2; set++)
    for (int r = 0; r < t->R; r++)
      for (int m = 0; m < 16; m++)
        for (int i = 0; i < 4; i++) {
          const double* Pi = &P[set][r][i * 4];
          const double x0 = m & 1 ? 1.0 : 0.0, x1 = m & 2 ? 1.0 : 0.0, x2 = m & 4 ? 1.0 : 0.0, x3 = m & 8 ? 1.0 : 0.0;
          volatile double a = Pi[0] * x0, b = Pi[1] * x1, c = Pi[2] * x2, d = Pi[3] * x3;
          volatile double s = a + b; s = s + c; s = s + d;
          lut[((size_t)(set * t->R + r) * 16 + m) * 4 + i] = s;
        }
  // stream-ordered: kernels of earlier evaluations on this stream have been issued before these copies
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_P, P, sizeof P, 0, cudaMemcpyHostToDevice, t->stream));
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_pi, pi, 4 * sizeof(double), 0, cudaMemcpyHostToDevice, t->stream));
  CUDA_TRY(cudaMemcpyAsync(t->d_lut, lut.data(), lut.size() * sizeof(double), cudaMemcpyHostToDevice, t->stream));
  CUDA_TRY(cudaStreamSynchronize(t->stream));   // lut / P are stack/heap temporaries
  t->constantsValid = true;
  return DMPC_OK;
}

// constants are per-module: a second handle on the same device may have overwritten them
std::mutex g_constMutex;
dmpc_tree* g_constOwner[64] = {nullptr};
// ... and because they are, evaluations of different handles on one device must not overlap: every call that launches
// kernels holds the device's lock until its results are back (calls are synchronous anyway)
std::mutex g_deviceMutex[64];
void claimConstants(dmpc_tree* t) {
  std::lock_guard<std::mutex> lk(g_constMutex);
  if (g_constOwner[t->device & 63] != t) { t->constantsValid = false; g_constOwner[t->device & 63] = t; }
}

// kernel launch, optionally as a programmatic dependent of the previous kernel in the stream (kernels.cuh: pdl_wait)
template <class... KArgs, class... Args>
cudaError_t launchK(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

int rootStage(dmpc_tree* t, double* ll);
int launchRootSequence(dmpc_tree* t);


// timing events: inside a stream capture they become event-record nodes that record the (external) event on every replay
int recordEv(dmpc_tree* t, int i) {
  if (!t->capturing && !t->timeThis) return DMPC_OK;      // small plans of a handle nobody asks for timings: no event records
  if (t->capturing) CUDA_TRY(cudaEventRecordWithFlags(t->ev[i], t->stream, cudaEventRecordExternal));
  else CUDA_TRY(cudaEventRecord(t->ev[i], t->stream));
  return DMPC_OK;
}

// segs != nullptr: block column z walks the tasks [segs[z], segs[z + 1]) of `lt` in order; else task z (stride gridDim.z)
int launchPrune(dmpc_tree* t, dim3 grid, const Task* lt, int nTasks, const int32_t* segs = nullptr) {
  PruneArgs a{};
  a.tasks = lt; a.nTasks = nTasks; a.prog = t->runTokens; a.tipList = t->runTips; a.violation = &t->rv.st->violation; a.flip = t->runFlip;
  a.segStart = segs;
  a.curList = t->runCur; a.curN = t->runCurN;
  t->runCurN = 0;                                  // the first launch of the evaluation takes them
  // 2 sites per thread, 5 blocks per SM (96 registers): measured best on B200 — 4 sites per thread (168-212 registers,
  // 12-16 warps per SM) is 1.3-1.6x slower, 6 blocks per SM (80 registers, spills) 1.05x (profiles/r02_prune_variants.md)
  if (t->exact) CUDA_TRY(launchK(k_prune<true, false, 2, 5>, grid, dim3(TILE / 2), 0, t->stream, t->pdl, t->dv, a));
  else CUDA_TRY(launchK(k_prune<false, false, 2, 5>, grid, dim3(TILE / 2), 0, t->stream, t->pdl, t->dv, a));
  return DMPC_OK;
}

// tasks | tokens | tip list of a plan, packed into the pinned staging buffer and uploaded in ONE copy; sets what
// runPlan / launchPrune read
int uploadPlan(dmpc_tree* t, const HostTree::Plan& p, bool withPending = false) {
  const size_t nT = p.tasks.size(), nK = p.tokens.size(), nL = p.tipList.size(), nS = p.segStart.size();
  if (nT > (size_t)t->nInt || nK > 2 * (size_t)t->nInt || nL > 2 * (size_t)t->nInt || nS > nT + 1)
    return fail(DMPC_ERR_STATE, "plan larger than its staging buffer");
  t->planOffTok = nT * sizeof(Task);
  t->planOffTip = t->planOffTok + nK * sizeof(Token);
  t->planOffSeg = t->planOffTip + nL * sizeof(int32_t);
  t->planBytes = t->planOffSeg + nS * sizeof(int32_t);
  if (nT) std::memcpy(t->h_plan, p.tasks.data(), nT * sizeof(Task));
  if (nK) std::memcpy(t->h_plan + t->planOffTok, p.tokens.data(), nK * sizeof(Token));
  if (nL) std::memcpy(t->h_plan + t->planOffTip, p.tipList.data(), nL * sizeof(int32_t));
  if (nS) std::memcpy(t->h_plan + t->planOffSeg, p.segStart.data(), nS * sizeof(int32_t));
  // the slot bits of the last rollback that the device has not seen yet ride along, minus the vertices this plan rewrites
  // (their new slot bit is written by the plan's own tasks): the first pruning launch applies them — no copy and no
  // launch of their own
  int nC = 0;
  if (withPending && t->pendingCur && nT) {
    int32_t* dst = reinterpret_cast<int32_t*>(t->h_plan + t->planBytes);
    for (int i = 0; i < t->pendingCur; i++)
      if (!t->ht.touched[(t->h_curList[i] >> 1) + t->ht.numTips]) dst[nC++] = t->h_curList[i];
    t->pendingCur = 0;
  }
  const size_t offCur = t->planBytes;
  t->planBytes += (size_t)nC * sizeof(int32_t);
  if (t->planBytes && t->planOnCopyStream) {
    // slab 0 of the alignment is in flight on the copy stream, and a copy queued on another stream is not served until
    // that stream's copies run dry (measured): the plan takes its turn IN the copy stream, between slab 0 and slab 1
    CUDA_TRY(cudaMemcpyAsync(t->d_plan, t->h_plan, t->planBytes, cudaMemcpyHostToDevice, t->pack.copyStream));
    CUDA_TRY(cudaEventRecord(t->pack.slabEv[MAX_SLABS], t->pack.copyStream));
    CUDA_TRY(cudaStreamWaitEvent(t->stream, t->pack.slabEv[MAX_SLABS], 0));
  } else if (t->planBytes) {
    CUDA_TRY(cudaMemcpyAsync(t->d_plan, t->h_plan, t->planBytes, cudaMemcpyHostToDevice, t->stream));
  }
  t->planOnCopyStream = false;
  t->runTasks = reinterpret_cast<const Task*>(t->d_plan);
  t->runTokens = reinterpret_cast<const Token*>(t->d_plan + t->planOffTok);
  t->runTips = reinterpret_cast<const int32_t*>(t->d_plan + t->planOffTip);
  t->runSegs = reinterpret_cast<const int32_t*>(t->d_plan + t->planOffSeg);
  t->runCur = reinterpret_cast<const int32_t*>(t->d_plan + offCur); t->runCurN = nC;
  t->runFlip = 0;
  return DMPC_OK;
}

// the pruning launches of the current plan over the site tiles [dv.tileOff, dv.tileOff + tiles): one per round, then the
// walked chain (tasks, tokens, tip lists and segments are already on the device)
int launchRounds(dmpc_tree* t, int tiles, int* launchesOut) {
  const HostTree::Plan& p = *t->runPlanMeta;
  const int nAll = (int)p.levelStart.size() - 1;
  const int nLevels = nAll - p.chainLevels;        // the trailing single-task levels go to k_walk as one chain
  if (tiles > 65535) return fail(DMPC_ERR_ARG, "too many site tiles for one launch (numLoci > 16.7M)");
  if ((int)p.unitSeg.size() != nLevels + 1) return fail(DMPC_ERR_STATE, "plan: launch units and segment table disagree");
  int launches = 0;
  for (int l = 0; l < nLevels; l++) {
    // one launch per unit (the materialised clades, then the rounds): grid = (rate, site tile, segment) — blocks that
    // share a task's masks and tokens are scheduled next to each other; gridDim.z is limited to 65535, wider units go
    // out in slices.  Segment boundaries are task indices of the whole plan, so every slice gets the plan's task array.
    const int s0 = p.unitSeg[l], nSeg = p.unitSeg[l + 1] - s0;
    for (int off = 0; off < nSeg; off += 65535) {
      const int m = std::min(65535, nSeg - off);
      const dim3 grid((unsigned)t->R, (unsigned)tiles, (unsigned)m);
      int rc = launchPrune(t, grid, t->runTasks, 0, t->runSegs + s0 + off);
      if (rc) return rc;
      t->st.kernel_launches++;
      launches++;
    }
  }
  if (p.chainLevels) {
    const int first = p.levelStart[nLevels], n = p.levelStart[nAll] - first;     // one task per level: n == chainLevels
    CUDA_TRY(launchK(k_walk, dim3((unsigned)t->R, (unsigned)tiles), dim3(TILE), (size_t)WALK_SMEM, t->stream, t->pdl, t->dv, t->runTasks + first, n,
                     t->runFlip, t->runCur, t->runCurN));
    t->runCurN = 0;
    t->st.kernel_launches++;
    launches++;
  }
  if (!t->capturing) CUDA_TRY(cudaGetLastError());
  *launchesOut += launches;
  return DMPC_OK;
}

// slab k of the alignment upload on the copy stream, and its event
int queueSlab(dmpc_tree* t, int k) {
  const int t0 = t->slabs.tile[k], t1 = t->slabs.tile[k + 1];
  const int c0 = t0 * TILE, c1 = std::min(t->S, t1 * TILE);      // real columns of the slab (pad columns have no source)
  if (c1 > c0)
    CUDA_TRY(cudaMemcpy2DAsync(t->slabs.stage + c0, (size_t)t->S, t->slabs.src + c0, t->slabs.pitch, (size_t)(c1 - c0), (size_t)t->N,
                               cudaMemcpyHostToDevice, t->pack.copyStream));
  CUDA_TRY(cudaEventRecord(t->pack.slabEv[k], t->pack.copyStream));
  return DMPC_OK;
}

int runPlan(dmpc_tree* t) {
  const HostTree::Plan& p = *t->runPlanMeta;
  int rcE, launches = 0;
  if ((rcE = recordEv(t, 0))) return rcE;
  if (t->slabs.n > 1 && !t->capturing) {
    // logLikCpp: slab k of the alignment is copied (copy stream), re-pitched and pruned while slab k + 1 is on its way
    const int nS = t->slabs.n;
    for (int k = 0; k < nS; k++) {
      const int t0 = t->slabs.tile[k], t1 = t->slabs.tile[k + 1];
      const int c0 = t0 * TILE;
      if (k > 0) { const int rq = queueSlab(t, k); if (rq) return rq; }     // (slab 0 was queued by createImpl, ahead of the planning)
      if (k == nS - 1) CUDA_TRY(cudaEventRecord(t->ev[6], t->pack.copyStream));
      CUDA_TRY(cudaStreamWaitEvent(t->stream, t->pack.slabEv[k], 0));
      {
        const int p1 = t1 * TILE;                  // the padded layout is filled up to the tile boundary
        const size_t n = (size_t)t->N * (p1 - c0);
        k_prepare_masks<<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>(t->slabs.stage, t->d_tipmask, t->N, t->S, t->Sp, c0, p1, &t->rv.st->badCells);
        t->st.kernel_launches++;
      }
      t->dv.tileOff = t0;
      const int rc = launchRounds(t, t1 - t0, &launches);
      t->dv.tileOff = 0;
      if (rc) return rc;
    }
    t->slabs.n = 0;                                // one-shot
  } else {
    const int rc = launchRounds(t, t->T, &launches);
    if (rc) return rc;
  }
  if ((rcE = recordEv(t, 1))) return rcE;
  t->st.last_levels = (int)p.levelStart.size() - 1;
  t->st.last_prune_launches = launches;
  t->st.last_walk_steps = p.chainLevels;
  t->st.last_materialised = p.materialised;
  return DMPC_OK;
}

// ComputeLoglik (AugTree.cpp:289-326): prune the dirty levels, then reduce at the root.
int evaluate(dmpc_tree* t, const double* within, const double* between, const double* pi, double* ll, bool lockHeld = false) {
  int rc;
  std::unique_lock<std::mutex> deviceLock(g_deviceMutex[t->device & 63], std::defer_lock);
  if (!t->isShard && !lockHeld) deviceLock.lock();            // (the leader of a multi-device handle holds the locks of all its devices)
  if ((rc = ensureDevice(t))) return rc;
  claimConstants(t);
  const auto h0 = std::chrono::steady_clock::now();
  auto usSince = [](std::chrono::steady_clock::time_point a) {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count();
  };
  if ((rc = uploadConstants(t, within, between, pi))) return rc;
  HostTree& ht = t->ht;
  HostTree::Plan& p = t->plan;
  const bool full = ht.allInvalid;
  // A full re-evaluation of an unchanged tree (every MCMC iteration proposes new within-cluster matrices ->
  // InvalidateAll) needs the same plan as the last one, with every slot toggled: replay the device-resident copy.
  bool hit = false;
  int flip = 0;
  auto& cc = t->cache;
  if (full && cc.valid && cc.fuse == ht.fuseTips && cc.kids == ht.kids && cc.within == ht.within) {
    flip = ht.cur[ht.root()] ^ cc.cur[ht.root()];
    // every internal vertex's slot bit must differ from the cached plan's by the same flip: eight vertices per compare
    const uint8_t* a = ht.cur.data() + ht.numTips;
    const uint8_t* b = cc.cur.data() + ht.numTips;
    const size_t n = (size_t)(ht.numVerts - ht.numTips);
    const uint64_t want = flip ? 0x0101010101010101ull : 0ull;
    uint64_t diff = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t x, y; std::memcpy(&x, a + i, 8); std::memcpy(&y, b + i, 8); diff |= (x ^ y) ^ want; }
    for (; i < n; i++) diff |= (uint64_t)((a[i] ^ b[i]) ^ flip);
    hit = diff == 0;
  }
  const HostTree::Plan* meta = &p;
  if (hit) {
    {
      uint8_t* c = ht.cur.data() + ht.numTips;
      const size_t n = (size_t)(ht.numVerts - ht.numTips);
      size_t i = 0;
      for (; i + 8 <= n; i += 8) { uint64_t x; std::memcpy(&x, c + i, 8); x ^= 0x0101010101010101ull; std::memcpy(c + i, &x, 8); }
      for (; i < n; i++) c[i] ^= 1;
      std::fill(ht.touched.begin() + ht.numTips, ht.touched.end(), (uint8_t)1);
      std::fill(ht.solved.begin() + ht.numTips, ht.solved.end(), (uint8_t)1);
    }
    ht.allInvalid = false;
    if ((rc = allocSlot(t, 0)) || (rc = allocSlot(t, 1))) return rc;
    meta = &cc.meta;
    t->runTasks = reinterpret_cast<const Task*>(cc.dPlan);
    t->runTokens = reinterpret_cast<const Token*>(cc.dPlan + cc.offTok);
    t->runTips = reinterpret_cast<const int32_t*>(cc.dPlan + cc.offTip);
    t->runSegs = reinterpret_cast<const int32_t*>(cc.dPlan + cc.offSeg);
    t->runCurN = 0; t->pendingCur = 0;             // every slot bit is rewritten by this plan: a pending rollback is moot
    t->runFlip = flip;
    t->planCacheHits++;
  } else {
    PooledVec<uint8_t> curBefore;
    if (full) curBefore = ht.cur;
    ht.plan(p);
    const int nT = (int)p.tasks.size();
    if (nT) {
      for (int sl = 0; sl < 2; sl++)
        if (p.uses[sl] && (rc = allocSlot(t, sl))) return rc;
    }
    if ((rc = uploadPlan(t, p, true))) return rc;
    if (t->pendingCur && (rc = flushCur(t))) return rc;       // (an empty plan launches nothing that could carry the slot bits)
    if (full && nT) {                                // remember it (one device-to-device copy, stream-ordered)
      int r2;
      if (!cc.dPlan && (r2 = t->alloc(&cc.dPlan, (size_t)t->nInt * PLAN_BYTES + PLAN_SLACK))) return r2;
      CUDA_TRY(cudaMemcpyAsync(cc.dPlan, t->d_plan, t->planBytes, cudaMemcpyDeviceToDevice, t->stream));
      cc.offTok = t->planOffTok; cc.offTip = t->planOffTip; cc.offSeg = t->planOffSeg; cc.bytes = t->planBytes;
      cc.kids = ht.kids; cc.within = ht.within; cc.cur = curBefore; cc.fuse = ht.fuseTips;
      cc.meta.levelStart = p.levelStart; cc.meta.unitSeg = p.unitSeg;
      cc.meta.chainLevels = p.chainLevels; cc.meta.materialised = p.materialised;
      cc.meta.storedReads = p.storedReads; cc.meta.tipReads = p.tipReads; cc.meta.fusedOps = p.fusedOps; cc.meta.fp64Ops = p.fp64Ops;
      cc.meta.tasks.clear(); cc.metaTasks = nT;
      cc.valid = true;
      t->graphVersion++;
    }
  }
  t->runPlanMeta = meta;
  const int nTasks = hit ? cc.metaTasks : (int)p.tasks.size();
  // a root-path proposal is a handful of microsecond kernels: three event records are a measurable share of its host
  // time, so they are taken only when someone reads them (dmpc_get_stats) — full evaluations always carry them
  t->timeThis = t->timingOn || full || nTasks > 256;
  const long long storedReads = meta->storedReads, tipReads = meta->tipReads, fusedOps = meta->fusedOps;
  t->st.last_host_plan_us = usSince(h0);
  const bool chainMode = t->quirk && t->R > 1;
  if (hit) {
    // replay (or first capture) of the whole device-side sequence of this cached plan: levels + gather
    auto& g = t->graphs[flip];
    const bool predicted = chainMode && t->predictActive;
    if (!g.exec || g.version != t->graphVersion || g.predicted != predicted) {
      if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
      cudaGraph_t graph = nullptr;
      CUDA_TRY(cudaStreamBeginCapture(t->stream, cudaStreamCaptureModeThreadLocal));
      t->capturing = true;
      int rcC = runPlan(t);
      if (!rcC) rcC = launchRootSequence(t);
      if (!rcC) rcC = recordEv(t, 2);
      t->capturing = false;
      const cudaError_t ce = cudaStreamEndCapture(t->stream, &graph);
      if (rcC) { if (graph) cudaGraphDestroy(graph); return rcC; }
      if (ce != cudaSuccess) return fail(DMPC_ERR_CUDA, std::string("stream capture: ") + cudaGetErrorString(ce));
      const cudaError_t ci = cudaGraphInstantiate(&g.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ci != cudaSuccess) return fail(DMPC_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ci));
      g.version = t->graphVersion;
      g.predicted = predicted;
    }
    t->seqPredicted = predicted;
    CUDA_TRY(cudaGraphLaunch(g.exec, t->stream));
    t->gatherPrelaunched = true;
    t->pubExpected++;
    t->graphReplays++;
    t->st.kernel_launches += t->st.last_prune_launches + (predicted ? 2 : 1);
  } else if ((rc = runPlan(t))) {
    return rc;
  }
  t->st.last_updates = (unsigned long long)(nTasks + fusedOps) * t->S * t->R;
  t->st.updates += t->st.last_updates;
  t->st.evaluations++;
  // compulsory HBM traffic of this evaluation: stored partials written (32 B) and read (32 B per stored operand),
  // one mask byte per tip operand; exponents move only where something rescaled (not counted)
  t->st.last_algorithmic_bytes = (unsigned long long)t->S * t->R * 32ull * (unsigned long long)(nTasks + storedReads) +
                                 (unsigned long long)t->S * (unsigned long long)tipReads;
  t->st.last_fp64_ops = (unsigned long long)meta->fp64Ops * t->S * t->R;
  t->st.last_stored_vertices = nTasks;
  t->st.last_fused_vertices = (int)fusedOps;
  const auto h1 = std::chrono::steady_clock::now();
  rc = rootStage(t, ll);
  t->st.last_host_submit_us = std::chrono::duration<double, std::micro>(h1 - h0).count() - t->st.last_host_plan_us;
  t->st.last_host_total_us = usSince(h0);
  if (rc) return rc;
  if (!t->exact && t->h_res->violation) {
    // a fused clade may have rescaled: redo the same plan with the exact kernel (same tasks, same slots) and
    // keep that mode for the life of the handle
    t->exact = true;
    t->graphVersion++;
    CUDA_TRY(cudaMemsetAsync(&t->rv.st->violation, 0, sizeof(int), t->stream));
    if ((rc = runPlan(t))) return rc;
    t->st.exact_mode = 1;
    return rootStage(t, ll);
  }
  return DMPC_OK;
}

// the peer mailboxes of a site-sharded run, as the root kernels see them (world = 1: no exchange)
void attachComm(dmpc_tree* t) {
  t->rv.comm.world = 1;
  if (!(t->deferred && t->comm && t->comm->connected && t->comm->device == t->device)) return;
  const dmpc_comm& c = *t->comm;
  if (XHDR + t->nChunks + t->T * CPT * xrec(t->rv.RP) > c.width) return;       // the message would not fit: host protocol
  CommView& cv = t->rv.comm;
  cv.world = c.world; cv.rank = c.rank; cv.width = c.width; cv.seq = c.seq;
  for (int p = 0; p < c.world; p++) {
    cv.flag[p] = (unsigned*)c.peer[p];
    cv.payload[p] = (double*)((char*)c.peer[p] + c.flagBytes());
    cv.cflag[p] = (unsigned*)((char*)c.peer[p] + c.flagBytes() + c.payloadBytes());
    cv.carry[p] = (float*)((char*)c.peer[p] + 2 * c.flagBytes() + c.payloadBytes());
  }
}

int launchGather(dmpc_tree* t, int mode, bool publish) {
  t->rv.exactMode = t->exact ? 1 : 0;
  t->rv.publishAtGather = publish ? 1 : 0;
  attachComm(t);
  if (t->deferred && mode == 1)                  // site-sharded: the gather's optimistic finish assumes the vector enters as zeros
    CUDA_TRY(cudaMemsetAsync(t->rv.carryIn, 0, RMAX * sizeof(float), t->stream));
  // (in a sharded run the memset above sits between the pruning and the gather: an ordinary dependency there)
  const bool pdl = t->pdl && !(t->deferred && mode == 1);
  if (t->rv.RP == 4) CUDA_TRY(launchK(k_root_gather<4, false>, dim3(t->T), dim3(TILE), 0, t->stream, pdl, t->dv, t->rv, mode, PathView{}));
  else CUDA_TRY(launchK(k_root_gather<8, false>, dim3(t->T), dim3(TILE), 0, t->stream, pdl, t->dv, t->rv, mode, PathView{}));
  t->st.kernel_launches++;
  if (!t->capturing) CUDA_TRY(cudaGetLastError());
  return DMPC_OK;
}

// reduceAll: 0 = chunk sums only (host protocol of a sharded run), 1 = this handle's total, 2 = cross-rank exchange
int launchFinal(dmpc_tree* t, int mode, int reduceAll, bool publish) {
  CUDA_TRY(launchK(k_final<false>, dim3(t->T), dim3(TILE), 0, t->stream, t->pdl, t->dv, t->rv, mode, reduceAll, publish ? 1 : 0, PathView{}));
  t->st.kernel_launches++;
  if (!t->capturing) CUDA_TRY(cudaGetLastError());
  return DMPC_OK;
}

// k_final(mode 2): the per-site terms under the certificate the gather's last block issued; publishes the result
int launchCertFinal(dmpc_tree* t) {
  return launchFinal(t, 2, t->rv.comm.world > 1 ? 2 : 1, true);
}

// What follows the pruning of an evaluation.  Usually no site rescaled last time and the gather alone finishes (and
// publishes); when the last evaluation did see rescaled sites, the certificate and the final kernel are queued behind
// the gather at once — no host round trip in between — and the final kernel publishes.
int launchRootSequence(dmpc_tree* t) {
  const bool chain = t->quirk && t->R > 1;
  attachComm(t);
  const bool hostProtocol = t->deferred && t->rv.comm.world == 1;
  const bool predicted = chain && !hostProtocol && t->predictActive;
  t->seqPredicted = predicted;
  int rc;
  if ((rc = launchGather(t, chain ? 1 : 0, !predicted))) return rc;
  if (predicted && (rc = launchCertFinal(t))) return rc;
  if (!t->capturing) t->pubExpected++;             // (a replayed graph counts its own publication)
  return DMPC_OK;
}

size_t chainSmem(const dmpc_tree* t) {   // 2 buffers of (2Q + 8) rows of RP floats, 1 marker byte per group (<= per row), metadata
  const size_t cap = 2 * (size_t)t->rv.Q + 8;
  return 2 * cap * t->rv.RP * sizeof(float) + 2 * cap + 4 * sizeof(int) + 16;
}

// rvBatch != nullptr: one block per proposal of a batch (dmpc_eval_proposals), working on per-proposal copies
template <int RP, int RC, int G>
int launchChainT(dmpc_tree* t, size_t smem, const RootView* rvBatch, int nBatch) {
  if (rvBatch) {
    CUDA_TRY(cudaFuncSetAttribute(k_chain<RP, RC, G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_chain<RP, RC, G, true><<<nBatch, 1024, smem, t->stream>>>(t->dv, *rvBatch);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(k_chain<RP, RC, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_chain<RP, RC, G, false><<<1, 1024, smem, t->stream>>>(t->dv, t->rv);
  }
  return DMPC_OK;
}
template <int RP, int RC>
int launchChainG(dmpc_tree* t, size_t smem, const RootView* rvBatch, int nBatch) {
  return t->chainGroup == 1 ? launchChainT<RP, RC, 1>(t, smem, rvBatch, nBatch) : launchChainT<RP, RC, 4>(t, smem, rvBatch, nBatch);
}

int launchChain(dmpc_tree* t, const RootView* rvBatch = nullptr, int nBatch = 0) {
  const size_t smem = chainSmem(t);
  if (smem > 220 * 1024) return fail(DMPC_ERR_ARG, "rescale list too deep for the chain kernel's staging buffers");
  int rc = DMPC_OK;
  switch (t->R) {
    case 2: rc = launchChainG<4, 2>(t, smem, rvBatch, nBatch); break;
    case 3: rc = launchChainG<4, 3>(t, smem, rvBatch, nBatch); break;
    case 4: rc = launchChainG<4, 4>(t, smem, rvBatch, nBatch); break;
    case 5: rc = launchChainG<8, 5>(t, smem, rvBatch, nBatch); break;
    case 6: rc = launchChainG<8, 6>(t, smem, rvBatch, nBatch); break;
    case 7: rc = launchChainG<8, 7>(t, smem, rvBatch, nBatch); break;
    default: rc = launchChainG<8, 8>(t, smem, rvBatch, nBatch); break;
  }
  if (rc) return rc;
  t->st.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  return DMPC_OK;
}

// The host's view of the status block after the kernels queued so far.
// polled: the last kernel queued publishes the finished block into the mapped host mirror (DevStatus::pubSeq) and the
// host polls for it — no copy queued behind the kernel, no stream synchronisation, which together cost more than a
// root-path evaluation's kernels.  Otherwise (literal chain fallback, or the chunk sums are wanted too): one D2H copy
// and the only synchronisation point.
int readStatus(dmpc_tree* t, bool withChunks, bool polled) {
  bool got = false;
  if (polled && !withChunks && t->rv.hostSt) {
    const volatile unsigned* seq = &t->h_res->pubSeq;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spin = 1;; spin++) {
      if (*seq == t->pubExpected) { got = true; break; }
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
      if ((spin & 0x3fff) == 0) {
        if (cudaStreamQuery(t->stream) != cudaErrorNotReady) { got = *seq == t->pubExpected; break; }   // finished (or failed)
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 60.0) break;
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  if (!got) {
    CUDA_TRY(cudaMemcpyAsync(t->h_res, t->rv.st, sizeof(DevStatus), cudaMemcpyDeviceToHost, t->stream));
    if (withChunks)
      CUDA_TRY(cudaMemcpyAsync(t->h_chunks, t->rv.chunk, (size_t)t->nChunks * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
  }
  t->st.last_active_tiles = t->h_res->nact;
  t->lastNact = t->h_res->nact;
  return DMPC_OK;
}

// bookkeeping once an evaluation's chain is resolved: what walked what (dmpc_stats), and how k_chain should group rows
void noteChain(dmpc_tree* t, bool anyActive) {
  const HostResult& h = *t->h_res;
  t->lastCert = anyActive ? h.cert : CERT_NONE;
  t->lastCertRate = t->lastCert == CERT_OK ? h.certRate : -1;
  t->lastCertSites = t->lastCert == CERT_OK ? h.certL0 : 0;
  t->st.last_chain_sites = anyActive && h.nact ? h.chainSites : 0;
  t->st.last_chain_rows = anyActive && h.nact ? h.chainRows : 0;
  t->st.last_cert = t->lastCert; t->st.last_cert_rate = t->lastCertRate;
  if (t->lastCert == CERT_LITERAL && h.nact && h.chainSites > 0) {
    // rows per rescaled site as the chain just saw them (padded to the group size): sparse lists walk faster one row at a time
    const double avg = (double)h.chainRows / h.chainSites;
    if (t->chainGroup == 4 && avg < 4.5) t->chainGroup = 1;
    else if (t->chainGroup == 1 && avg > 6.0) t->chainGroup = 4;
  }
}

// ComputeLoglik's reduction (AugTree.cpp:296-325).  k_root_gather finishes the evaluation by itself when no site
// rescaled anywhere; otherwise the gather's last block resolves the cross-locus chain by certificate and k_final adds the terms up
// (queued behind the gather already when the previous evaluation had rescaled sites); only when the certificate
// does not hold do the literal chain kernel and a second final pass run.
// Deferred (site-sharded) mode without peer mailboxes: stops after the gather; the chunk sums it produced — valid when
// this shard holds no rescaled site and the exponent vector enters as zeros — are already on the host for
// dmpc_shard_finish.
int rootStage(dmpc_tree* t, double* ll) {
  int rc;
  const bool chain = t->quirk && t->R > 1;
  if (t->deferred && chain) t->carryIsZero = true;   // launchGather zeroes the entry vector: optimistic pass
  for (int attempt = 0; attempt < 8; attempt++) {
    if (t->gatherPrelaunched) {
      t->gatherPrelaunched = false;               // the replayed graph already ran the root sequence (and recorded ev[2])
    } else {
      if ((rc = launchRootSequence(t))) return rc;
      if (t->timeThis) CUDA_TRY(cudaEventRecord(t->ev[2], t->stream));
    }
    const bool sharded = t->rv.comm.world > 1;
    const bool hostProtocol = t->deferred && !sharded;
    if ((rc = readStatus(t, hostProtocol, true))) return rc;
    if (t->h_res->overflow > t->rv.Kcap) {
      int k = t->rv.Kcap;
      while (k < t->h_res->overflow) k *= 2;
      CUDA_TRY(cudaMemsetAsync(&t->rv.st->overflow, 0, sizeof(int), t->stream));
      if ((rc = allocElist(t, k))) return rc;
      continue;
    }
    if (t->timeThis) {
      float ms = 0;
      cudaEventElapsedTime(&ms, t->ev[0], t->ev[1]); t->st.last_prune_ms = ms; t->st.prune_ms_total += ms;
      t->evalTimePending = true;
    } else {
      t->st.last_prune_ms = 0.0; t->st.last_eval_ms = 0.0;
    }
    if (hostProtocol) {
      t->chunksOnHost = t->h_res->nact == 0;
      *ll = NAN;
      return DMPC_OK;
    }
    if (sharded && t->h_res->xok == -1) { *ll = NAN; return DMPC_OK; }     // fused-clade violation: evaluate() re-runs in exact mode
    if (!sharded && !t->exact && t->h_res->violation) { *ll = NAN; return DMPC_OK; }   // same, single handle
    // did any site (of any rank) rescale?  Sharded: the gather's exchange answered 1 when nobody did.
    bool anyActive = sharded ? !(t->h_res->xok == 1 && t->h_res->cert == CERT_NONE) : t->h_res->nact != 0;
    if (sharded && t->h_res->xok == 2) return fail(DMPC_ERR_CUDA, "peer exchange failed: a rank of the site-sharded run never answered");
    if (!chain && anyActive) {
      // textbook reduction / one rate category: every site folds its own exponent list
      if ((rc = launchFinal(t, 0, sharded ? 2 : 1, false))) return rc;
      if (t->timeThis) CUDA_TRY(cudaEventRecord(t->ev[2], t->stream));
      if ((rc = readStatus(t, false, false))) return rc;
    } else if (anyActive) {
      if (!t->seqPredicted) {                       // the gather ran alone: certificate + final now
        if ((rc = launchCertFinal(t))) return rc;
        t->pubExpected++;
        if (t->timeThis) CUDA_TRY(cudaEventRecord(t->ev[2], t->stream));
        if ((rc = readStatus(t, false, true))) return rc;
      }
      if (t->h_res->cert != CERT_OK) {              // no certificate: the literal chain (handed from rank to rank when sharded)
        if ((rc = launchChain(t))) return rc;
        if ((rc = launchFinal(t, 1, sharded ? 2 : 1, false))) return rc;
        if (t->timeThis) CUDA_TRY(cudaEventRecord(t->ev[2], t->stream));
        if ((rc = readStatus(t, false, false))) return rc;
        t->carryIsZero = false;
      }
    }
    noteChain(t, chain && anyActive);
    t->predictActive = chain && anyActive;
    t->chunksOnHost = false;
    if (sharded && t->h_res->xok != 1) return fail(DMPC_ERR_CUDA, "peer exchange failed: a rank of the site-sharded run never answered");
    *ll = t->h_res->ll;
    return DMPC_OK;
  }
  return fail(DMPC_ERR_CUDA, "rescale list kept overflowing");
}

void destroy(dmpc_tree* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  for (auto& g : t->graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  if (t->stream) {
    if (t->pack.copyStream) cudaStreamSynchronize(t->pack.copyStream);   // a create that failed may still have slabs in flight
    for (void* p : t->allocs) cudaFreeAsync(p, t->stream);
    cudaStreamSynchronize(t->stream);                  // the pack (stream, pinned buffers) goes back idle
  }
  releasePack(t);
  {
    std::lock_guard<std::mutex> lk(g_constMutex);
    if (g_constOwner[t->device & 63] == t) g_constOwner[t->device & 63] = nullptr;
  }
  delete t;
}

int checkHandle(dmpc_tree* t) {
  if (!t) return fail(DMPC_ERR_ARG, "pointer is null.");   // clusterFunArmadillo.cpp:136
  if (!t->alive) return fail(DMPC_ERR_STATE, "tree handle used after dmpc_final_deallocate");
  return DMPC_OK;
}

}  // namespace

// ================================================================================== C ABI
extern "C" {
static int createImpl(const int32_t* edge, int nEdgeRows, const double* clusterMRCAs, int nClusterMRCAs,
                      const double* limProbs, const double* withinTransMats, const double* betweenTransMats,
                      int numRateCats, const uint8_t* alignmentBin, size_t alignPitch, int numTips, int numLoci,
                      int withinMatListIndex, int betweenMatListIndex, const dmpc_options& o, bool isShard,
                      dmpc_tree** treeOut, double* logLikOut);
}

// ================================================================================== single-process multi-device handles
// One dmpc_tree (the LEADER, which owns no device memory) over n shard handles, each holding a contiguous block of
// sites on its own device: what `DMPC_DEVICES=0,1,...` gives an R session, which is one process (R/DMphyClusInterface.R:71
// stays unchanged).  Every entry point fans out to the shards — one host thread each, because the shards' root kernels
// wait for one another inside the fused NVLink exchange — and all shards return the same log-likelihood, bit for bit
// the single-device value.  The mailboxes are the ones of the process-per-GPU path (dmpc_comm), reached through peer
// access instead of CUDA IPC.
struct Worker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int()> job;
  bool has = false, done = false, quit = false;
  int rc = 0;
  std::string err;
  void loop() {
    for (;;) {
      std::unique_lock<std::mutex> lk(m);
      cv.wait(lk, [&] { return has || quit; });
      if (quit) return;
      auto j = job;
      lk.unlock();
      const int r = j();
      lk.lock();
      rc = r; err = r ? g_err : std::string();
      has = false; done = true;
      cv.notify_all();
    }
  }
};

struct Group {
  std::vector<dmpc_tree*> shards;
  std::vector<dmpc_comm*> comms;
  std::vector<std::unique_ptr<Worker>> workers;    // workers[i] serves shards[i + 1]; shard 0 runs on the calling thread
  std::vector<int> begin, end;                     // site blocks
  std::vector<int> lockDevices;                    // sorted, unique
  int S = 0;
};

namespace {

// fn(shard index, shard) on every shard at once; the device locks of all the group's devices are held meanwhile
int groupRun(Group* g, const std::function<int(int, dmpc_tree*)>& fn) {
  for (int d : g->lockDevices) g_deviceMutex[d & 63].lock();
  const int n = (int)g->shards.size();
  for (int i = 1; i < n; i++) {
    Worker& w = *g->workers[i - 1];
    std::lock_guard<std::mutex> lk(w.m);
    w.job = [&fn, i, g] { return fn(i, g->shards[i]); };
    w.has = true; w.done = false;
    w.cv.notify_all();
  }
  int rc = fn(0, g->shards[0]);
  std::string err = rc ? g_err : std::string();
  for (int i = 1; i < n; i++) {
    Worker& w = *g->workers[i - 1];
    std::unique_lock<std::mutex> lk(w.m);
    w.cv.wait(lk, [&] { return w.done; });
    if (w.rc && !rc) { rc = w.rc; err = w.err; }
  }
  for (auto it = g->lockDevices.rbegin(); it != g->lockDevices.rend(); ++it) g_deviceMutex[*it & 63].unlock();
  if (rc) g_err = err;
  return rc;
}

// a likelihood entry point on every shard; all shards must come back with the same value (they reduce the same chunk
// sums in the same order)
int groupLoglik(Group* g, double* logLikOut, const std::function<int(int, dmpc_tree*, double*)>& fn) {
  std::vector<double> ll(g->shards.size(), NAN);
  const int rc = groupRun(g, [&](int i, dmpc_tree* sh) { return fn(i, sh, &ll[i]); });
  if (rc) return rc;
  for (size_t i = 1; i < ll.size(); i++)
    if (std::memcmp(&ll[i], &ll[0], sizeof(double))) return fail(DMPC_ERR_STATE, "the shards of a multi-device handle disagree on the log-likelihood");
  *logLikOut = ll[0];
  return DMPC_OK;
}

void destroyGroup(dmpc_tree* leader) {
  Group* g = leader->grp;
  if (g) {
    for (auto& w : g->workers) {
      { std::lock_guard<std::mutex> lk(w->m); w->quit = true; }
      w->cv.notify_all();
      if (w->th.joinable()) w->th.join();
    }
    for (dmpc_tree* sh : g->shards) destroy(sh);
    for (dmpc_comm* c : g->comms) dmpc_comm_destroy(c);
    delete g;
  }
  delete leader;
}

// contiguous site block of shard r of n: multiples of RCH sites, so that the chunk boundaries — and the summation
// tree — do not depend on the number of shards (same rule as dmphyclus_b200/sharded.py::shard_range)
void shardRange(int S, int r, int n, int* b, int* e) {
  long per = ((long)S + n - 1) / n;
  per = (per + RCH - 1) / RCH * RCH;
  *b = (int)std::min<long>(S, r * per);
  *e = (int)std::min<long>(S, *b + per);
}

int createGroup(const int32_t* edge, int nEdgeRows, const double* clusterMRCAs, int nClusterMRCAs, const double* limProbs,
                const double* withinTransMats, const double* betweenTransMats, int numRateCats, const uint8_t* alignmentBin,
                int numTips, int numLoci, int withinMatListIndex, int betweenMatListIndex, const dmpc_options& o, int nDev,
                dmpc_tree** treeOut, double* logLikOut) {
  if (o.n_devices > CW) return fail(DMPC_ERR_ARG, "at most 8 devices per handle");
  if (o.deferred || o.comm) return fail(DMPC_ERR_ARG, "n_devices > 1 excludes the process-per-GPU options (deferred, comm)");
  for (int i = 0; i < o.n_devices; i++)
    if (o.devices[i] < 0 || o.devices[i] >= nDev) return fail(DMPC_ERR_ARG, "dmpc_options.devices: no such CUDA device");
  // shards that would own no site are left out (short alignments)
  int n = o.n_devices;
  while (n > 1) { int b, e; shardRange(numLoci, n - 1, n, &b, &e); if (e > b) break; n--; }
  dmpc_tree* leader = new dmpc_tree;
  Group* g = new Group;
  leader->grp = g;
  leader->N = numTips; leader->S = numLoci; leader->R = numRateCats; leader->device = o.devices[0];
  g->S = numLoci;
  auto bail = [&](int rc) { const std::string keep = g_err; destroyGroup(leader); g_err = keep; return rc; };
  int maxChunks = 1;
  for (int r = 0; r < n; r++) {
    int b, e;
    shardRange(numLoci, r, n, &b, &e);
    g->begin.push_back(b); g->end.push_back(e);
    maxChunks = std::max(maxChunks, (e - b + RCH - 1) / RCH);
    if (std::find(g->lockDevices.begin(), g->lockDevices.end(), o.devices[r]) == g->lockDevices.end()) g->lockDevices.push_back(o.devices[r]);
  }
  std::sort(g->lockDevices.begin(), g->lockDevices.end());
  // peer access between the devices, then one mailbox per shard, addressed directly
  for (int a : g->lockDevices)
    for (int b : g->lockDevices) {
      if (a == b) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) return bail(fail(DMPC_ERR_CUDA, "devices of a multi-device handle need peer access to one another"));
      cudaSetDevice(a);
      const cudaError_t pe = cudaDeviceEnablePeerAccess(b, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return bail(fail(DMPC_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe)));
      cudaGetLastError();
    }
  for (int r = 0; r < n; r++) {
    dmpc_comm* c = nullptr;
    char handle[64];
    const int rc = dmpc_comm_create(o.devices[r], r, n, maxChunks, &c, handle);
    if (rc) return bail(rc);
    c->ipc = false;
    g->comms.push_back(c);
  }
  for (int r = 0; r < n; r++) {
    for (int p2 = 0; p2 < n; p2++) g->comms[r]->peer[p2] = g->comms[p2]->local;
    g->comms[r]->connected = true;
  }
  for (int r = 1; r < n; r++) {
    g->workers.emplace_back(new Worker);
    Worker* w = g->workers.back().get();
    w->th = std::thread([w] { w->loop(); });
  }
  g->shards.assign(n, nullptr);
  // the shards are created together as well: their first evaluation already meets in the exchange
  std::vector<double> ll(n, NAN);
  std::vector<dmpc_tree*> made(n, nullptr);
  const int rc = [&]() {
    g->shards.assign(n, leader);                    // placeholders (never dereferenced by the jobs below)
    return groupRun(g, [&](int r, dmpc_tree*) {
      dmpc_options so = o;
      so.n_devices = 0; so.device = o.devices[r]; so.deferred = 1; so.comm = g->comms[r];
      so.site_begin = g->begin[r]; so.sites_total = numLoci;
      return createImpl(edge, nEdgeRows, clusterMRCAs, nClusterMRCAs, limProbs, withinTransMats, betweenTransMats, numRateCats,
                        alignmentBin + g->begin[r], (size_t)numLoci, numTips, g->end[r] - g->begin[r], withinMatListIndex,
                        betweenMatListIndex, so, true, &made[r], &ll[r]);
    });
  }();
  g->shards.clear();
  for (dmpc_tree* sh : made) if (sh) g->shards.push_back(sh);
  if (rc) return bail(rc);
  for (int r = 1; r < n; r++)
    if (std::memcmp(&ll[r], &ll[0], sizeof(double))) return bail(fail(DMPC_ERR_STATE, "the shards of a multi-device handle disagree on the log-likelihood"));
  *logLikOut = ll[0];
  *treeOut = leader;
  return DMPC_OK;
}

}  // namespace

extern "C" {

void dmpc_default_options(dmpc_options* o) {
  std::memset(o, 0, sizeof *o);
  o->device = -1; o->quirk_carry = 1; o->fuse_tips = 15; o->chain_cert = 1;
}

const char* dmpc_last_error(void) { return g_err.c_str(); }
const char* dmpc_version(void) { return "dmphyclus_b200 0.1 (sm_100a)"; }

long dmpc_convert_alignment(const char* chars, long nCells, const char* states, uint8_t* maskOut) {
  if (!chars || !states || !maskOut) { g_err = "null pointer"; return -1; }
  uint8_t map[256];
  std::memset(map, 0, sizeof map);
  for (int j = 0; j < 4; j++) map[(unsigned char)states[j]] = (uint8_t)(1u << j);
  // IUPAC unions are keyed on the literal labels a,t,c,g (clusterFunArmadillo.cpp:75-91)
  const uint8_t a = map['a'], tt = map['t'], c = map['c'], g = map['g'];
  map['r'] = a | g; map['y'] = c | tt; map['s'] = c | g; map['w'] = a | tt; map['k'] = g | tt; map['m'] = a | c;
  map['b'] = c | g | tt; map['d'] = a | g | tt; map['h'] = a | c | tt; map['v'] = a | c | g;
  map['-'] = map['n'] = map['.'] = a | tt | c | g;
  long bad = 0;
  for (long i = 0; i < nCells; i++) { const uint8_t m = map[(unsigned char)chars[i]]; maskOut[i] = m; bad += m == 0; }
  return bad;
}

// one handle on one device.  alignmentBin holds this handle's numLoci columns of every tip, rows alignPitch bytes apart
// (the whole alignment's row length when the handle is a shard of a multi-device handle)
static int createImpl(const int32_t* edge, int nEdgeRows, const double* clusterMRCAs, int nClusterMRCAs,
                      const double* limProbs, const double* withinTransMats, const double* betweenTransMats,
                      int numRateCats, const uint8_t* alignmentBin, size_t alignPitch, int numTips, int numLoci,
                      int withinMatListIndex, int betweenMatListIndex, const dmpc_options& o, bool isShard,
                      dmpc_tree** treeOut, double* logLikOut) {
  dmpc_tree* t = new dmpc_tree;
  t->isShard = isShard;
  if (o.device < 0) { if (cudaGetDevice(&t->device) != cudaSuccess) t->device = 0; } else t->device = o.device;
  if (nEdgeRows != 2 * numTips - 2) { delete t; return fail(DMPC_ERR_ARG, "edge matrix must have 2*numTips-2 rows"); }
  t->ht.withinIdx = withinMatListIndex; t->ht.betweenIdx = betweenMatListIndex;
  t->N = numTips; t->S = numLoci; t->R = numRateCats;
  t->nReal = numTips - 1; t->nInt = numTips;      // rows of the per-vertex arrays: N-1 internal vertices + 1 scratch row
  t->nFlag = (numTips + 15) & ~15;
  t->ht.fuseTips = std::max(1, std::min(o.fuse_tips, (int)FUSE_MAX_TIPS));
  if (const char* pe = std::getenv("DMPC_PDL")) t->pdl = std::atoi(pe) != 0;
  if (const char* cm = std::getenv("DMPC_CHAIN_MIN")) t->ht.chainMin = std::max((int)CHAIN_MIN, std::atoi(cm));
  if (const char* gc = std::getenv("DMPC_GROUP_COST")) t->ht.groupCost = std::max(0, std::atoi(gc));   // experiments: 0 = one launch per level
  t->exact = o.fuse_exact != 0;
  t->Sp = (numLoci + TILE - 1) / TILE * TILE;
  t->T = t->Sp / TILE;
  t->nChunks = (numLoci + RCH - 1) / RCH;
  t->rv.nChunks = t->nChunks;
  t->rv.certOn = o.chain_cert != 0;
  t->ht.walkChains = true;      // (k_walk addresses rows, with 64-bit products at the point of use: no size limit)
  t->quirk = o.quirk_carry != 0;
  t->deferred = o.deferred != 0;
  t->comm = (dmpc_comm*)o.comm;
  t->siteBegin = o.site_begin; t->sitesTotal = o.sites_total ? o.sites_total : numLoci;
  t->rv.RP = numRateCats <= 4 ? 4 : 8;

  int rc = DMPC_OK;
  const auto c0 = std::chrono::steady_clock::now();
  auto body = [&]() -> int {
    CUDA_TRY(cudaSetDevice(t->device));
    int r2;
    if ((r2 = acquirePack(t))) return r2;
    // every buffer whose size is known now comes out of ONE stream-ordered allocation (a create / evaluate / destroy
    // cycle, i.e. logLikFromClusInd, would otherwise spend more time in ~30 cudaMallocAsync / cudaFreeAsync calls than
    // in the upload): alignment staging, masks, plan, root-reduction work arrays, slot 0 of the partials
    struct { std::vector<std::pair<void*, size_t>> req; size_t total = 0; } ar;
    auto want = [&](auto** p, size_t n) {
      ar.total = (ar.total + 255) & ~(size_t)255;
      ar.req.push_back({(void*)p, ar.total});
      ar.total += n * sizeof(**p);
    };
    uint8_t* stage = nullptr;
    const size_t nPart = (size_t)t->nInt * t->R * t->Sp;
    want(&t->dv.part[0], nPart * 4);
    want(&t->dv.expo[0], nPart);
    want(&t->dv.flag[0], (size_t)t->R * t->T * t->nFlag);
    want(&t->d_tipmask, (size_t)t->N * t->Sp);
    want(&stage, (size_t)t->N * t->S);
    want(&t->d_lut, (size_t)2 * t->R * 64);
    want(&t->d_plan, (size_t)t->nInt * PLAN_BYTES + PLAN_SLACK);
    want(&t->d_curList, (size_t)t->nInt);
    want(&t->dv.cur, (size_t)t->nInt);
    want(&t->rv.kmax, (size_t)t->Sp);
    want(&t->rv.rootdot, (size_t)t->R * t->Sp);
    want(&t->rv.act, (size_t)t->Sp);
    want(&t->rv.rowoff, (size_t)t->Sp + 1);
    want(&t->rv.posIncl, (size_t)t->Sp);
    want(&t->rv.chunkStart, (size_t)t->Sp + 2);
    want(&t->rv.cout, (size_t)t->Sp * t->rv.RP);
    want(&t->rv.mxo, (size_t)t->Sp);
    want(&t->rv.carryIn, (size_t)RMAX);
    want(&t->rv.carryOut, (size_t)RMAX);
    want(&t->rv.sitell, (size_t)t->Sp);
    want(&t->rv.chunk, (size_t)t->T * CPT);
    want(&t->rv.tileRec, (size_t)t->T * CPT * xrec(t->rv.RP));
    want(&t->rv.st, (size_t)1);
    want(&t->rv.nactCtr, (size_t)1);
    // one entry per certificate record of EVERY rank of a sharded run (ranks' blocks differ in size: the last one may be tiny)
    want(&t->rv.certDom, std::max((size_t)t->T * CPT, (size_t)t->sitesTotal / RCH + (size_t)(CPT + 1) * CW + 8));
    setElistGeometry(t, 8);
    want(&t->rv.elist, (size_t)t->Sp * t->rv.Kcap * t->rv.RP);
    want(&t->rv.eid, (size_t)t->Sp * t->rv.Kcap * t->rv.RP);
    {
      void* base = nullptr;
      CUDA_TRY(cudaMallocAsync(&base, ar.total, t->stream));
      t->allocs.push_back(base);
      t->st.device_bytes += ar.total;
      for (auto& rq : ar.req) { void* q = (char*)base + rq.second; std::memcpy(rq.first, &q, sizeof q); }
    }
    CUDA_TRY(cudaMemsetAsync(t->rv.carryIn, 0, RMAX * sizeof(float), t->stream));
    CUDA_TRY(cudaMemsetAsync(t->rv.st, 0, sizeof(DevStatus), t->stream));
    CUDA_TRY(cudaMemsetAsync(t->rv.nactCtr, 0, sizeof(int), t->stream));
    CUDA_TRY(cudaMemsetAsync(&t->rv.st->stamp[0], 0xff, sizeof(unsigned long long), t->stream));
    CUDA_TRY(cudaMemsetAsync(&t->rv.st->stamp[4], 0xff, sizeof(unsigned long long), t->stream));
    t->dv.tipmask = t->d_tipmask; t->dv.lut = t->d_lut;
    t->dv.N = t->N; t->dv.S = t->S; t->dv.Sp = t->Sp; t->dv.R = t->R; t->dv.T = t->T; t->dv.nInt = t->nInt; t->dv.nReal = t->nReal; t->dv.nFlag = t->nFlag;
    // alignment: one contiguous upload, re-pitched to [N][Sp] (pad sites = all-ones mask) and validated on the device;
    // the bad-cell count comes back with the evaluation's status, so the upload costs no synchronisation of its own
    t->st.create_setup_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - c0).count();
    // the upload overlaps the pruning: site tiles go up in slabs on the copy stream, and every slab is re-pitched and
    // pruned (all rounds of the plan) as soon as it has landed — the upload (PCIe) is the longer of the two
    // how many slabs: every slab costs its own round launches (~20 us), and saves 1/n of the pruning from the critical
    // path: n ~ sqrt(pruning time / 20 us), with the pruning estimated at 4.2 ps per node-site-rate update (measured:
    // 2 slabs are best for 1,000 x 10,000, 8 for 10,000 x 30,000)
    const double pruneUs = (double)t->nReal * t->S * t->R * 4.2e-6;
    int nSlab = (int)std::lround(std::sqrt(pruneUs / 20.0));
    nSlab = std::max(1, std::min({nSlab, 8, t->T / 4}));
    if (const char* e = std::getenv("DMPC_UPLOAD_SLABS")) nSlab = std::max(1, std::min({MAX_SLABS, std::atoi(e), t->T}));
    // the matrices go up first: once the alignment is under way, other copies wait for it (see uploadPlan).  They live in
    // the device's constant bank, which this handle now holds until the evaluation's results are back
    std::unique_lock<std::mutex> deviceLock(g_deviceMutex[t->device & 63], std::defer_lock);
    if (!t->isShard) deviceLock.lock();
    claimConstants(t);
    if ((r2 = uploadConstants(t, withinTransMats, betweenTransMats, limProbs))) return r2;
    if (nSlab > 1) {
      t->planOnCopyStream = true;
      // the copy stream must not run ahead of the allocation / memsets queued on the main stream
      CUDA_TRY(cudaEventRecord(t->pack.slabEv[7], t->stream));
      CUDA_TRY(cudaStreamWaitEvent(t->pack.copyStream, t->pack.slabEv[7], 0));
      t->slabs.cut(t->T, nSlab); t->slabs.src = alignmentBin; t->slabs.pitch = alignPitch; t->slabs.stage = stage;
      // slab 0 is queued now, before the host has even linked the tree: PCIe works on it while the host builds the tree
      // and the plan (0.35 + 0.3 ms at 10,000 tips; a slab of 10,000 x 30,000 takes 0.7 ms).  Only slab 0: a copy
      // queued on another stream is not served while the copy stream still has copies queued, and the plan is a copy
      // as well — with every slab queued here the pruning started only after the last one (measured: 10.9 ms, not 7.0)
      CUDA_TRY(cudaEventRecord(t->ev[5], t->pack.copyStream));
      if ((r2 = queueSlab(t, 0))) return r2;
    } else {
      CUDA_TRY(cudaEventRecord(t->ev[5], t->stream));
      // (a 2-D copy is programmed row by row: 100,000 rows of 1,500 bytes ran at 11.7 GB/s, 13.1 ms for 150 MB)
      if (alignPitch == (size_t)t->S)
        CUDA_TRY(cudaMemcpyAsync(stage, alignmentBin, (size_t)t->N * t->S, cudaMemcpyHostToDevice, t->stream));
      else
        CUDA_TRY(cudaMemcpy2DAsync(stage, (size_t)t->S, alignmentBin, alignPitch, (size_t)t->S, (size_t)t->N, cudaMemcpyHostToDevice, t->stream));
      {
        const size_t n = (size_t)t->N * t->Sp;
        k_prepare_masks<<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>(stage, t->d_tipmask, t->N, t->S, t->Sp, 0, t->Sp, &t->rv.st->badCells);
        t->st.kernel_launches++;
      }
      CUDA_TRY(cudaEventRecord(t->ev[6], t->stream));
    }
    // the tree is linked and validated while the alignment is in flight
    if (!t->ht.build(edge, nEdgeRows, numTips)) return fail(DMPC_ERR_ARG, t->ht.err);
    t->ht.associate(clusterMRCAs, nClusterMRCAs);
    int rcEval = evaluate(t, withinTransMats, betweenTransMats, limProbs, logLikOut, true);
    if (rcEval == DMPC_OK) {                     // (the upload's closing event sits on the copy stream: it has long fired)
      float ms = 0;
      if (cudaEventSynchronize(t->ev[6]) == cudaSuccess && cudaEventElapsedTime(&ms, t->ev[5], t->ev[6]) == cudaSuccess) t->st.create_upload_ms = ms;
    }
    if (rcEval == DMPC_OK && t->h_res->badCells)
      return fail(DMPC_ERR_ARG, std::to_string(t->h_res->badCells) +
                                    " alignment cells hold no state (unknown symbol; clusterFunArmadillo.cpp:109)");
    return rcEval;
  };
  rc = body();
  if (rc != DMPC_OK) { std::string keepMsg = g_err; destroy(t); g_err = keepMsg; return rc; }
  t->ht.negateAllUpdateFlags();   // nothing to roll back to before the first proposal
  *treeOut = t;
  return DMPC_OK;
}

int dmpc_loglik_create(const int32_t* edge, int nEdgeRows, const double* clusterMRCAs, int nClusterMRCAs,
                       const double* limProbs, const double* withinTransMats, const double* betweenTransMats,
                       int numRateCats, const uint8_t* alignmentBin, int numTips, int numLoci,
                       int withinMatListIndex, int betweenMatListIndex, const dmpc_options* opts,
                       dmpc_tree** treeOut, double* logLikOut) {
  if (!edge || !limProbs || !withinTransMats || !betweenTransMats || !alignmentBin || !treeOut || !logLikOut ||
      (nClusterMRCAs > 0 && !clusterMRCAs))
    return fail(DMPC_ERR_ARG, "null pointer");
  if (numRateCats < 1 || numRateCats > RMAX) return fail(DMPC_ERR_ARG, "numRateCats must be in 1..8");
  if (numTips < 2 || numLoci < 1) return fail(DMPC_ERR_ARG, "need at least 2 tips and 1 locus");
  dmpc_options o;
  dmpc_default_options(&o);
  if (opts) o = *opts;
  int nDev = 0;
  if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
    return fail(DMPC_ERR_CUDA, "no CUDA device: this engine has no CPU path");
  if (o.n_devices > 1)
    return createGroup(edge, nEdgeRows, clusterMRCAs, nClusterMRCAs, limProbs, withinTransMats, betweenTransMats, numRateCats,
                       alignmentBin, numTips, numLoci, withinMatListIndex, betweenMatListIndex, o, nDev, treeOut, logLikOut);
  return createImpl(edge, nEdgeRows, clusterMRCAs, nClusterMRCAs, limProbs, withinTransMats, betweenTransMats, numRateCats,
                    alignmentBin, (size_t)numLoci, numTips, numLoci, withinMatListIndex, betweenMatListIndex, o, false, treeOut,
                    logLikOut);
}

int dmpc_new_between_transprobs_loglik(dmpc_tree* t, const double* withinTransMats, const double* newBetweenTransMats,
                                       int newBetweenMatListIndex, const double* limProbs, double* logLikOut) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!withinTransMats || !newBetweenTransMats || !limProbs || !logLikOut) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->grp)
    return groupLoglik(t->grp, logLikOut, [&](int, dmpc_tree* sh, double* ll) {
      return dmpc_new_between_transprobs_loglik(sh, withinTransMats, newBetweenTransMats, newBetweenMatListIndex, limProbs, ll);
    });
  t->ht.negateAllUpdateFlags();
  t->ht.betweenIdx = newBetweenMatListIndex;
  t->ht.invalidateBetween();
  return evaluate(t, withinTransMats, newBetweenTransMats, limProbs, logLikOut);
}

int dmpc_new_within_transprobs_loglik(dmpc_tree* t, const double* newWithinTransMats, const double* betweenTransMats,
                                      int newWithinMatListIndex, const double* limProbs, double* logLikOut) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!newWithinTransMats || !betweenTransMats || !limProbs || !logLikOut) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->grp)
    return groupLoglik(t->grp, logLikOut, [&](int, dmpc_tree* sh, double* ll) {
      return dmpc_new_within_transprobs_loglik(sh, newWithinTransMats, betweenTransMats, newWithinMatListIndex, limProbs, ll);
    });
  t->ht.negateAllUpdateFlags();
  t->ht.withinIdx = newWithinMatListIndex;
  t->ht.invalidateAll();
  return evaluate(t, newWithinTransMats, betweenTransMats, limProbs, logLikOut);
}

static int nniCommon(dmpc_tree* t, const double* w, const double* b, int origin, const std::vector<long>* mrcas,
                     int numMoves, const double* pi, double* ll, int32_t* edgeOut) {
  t->ht.negateAllUpdateFlags();
  std::vector<int> cand;
  if (mrcas) t->ht.markMrcas(*mrcas);
  struct Unmark { HostTree& h; ~Unmark() { h.unmarkMrcas(); } } unmark{t->ht};
  t->ht.nniCandidates(origin, mrcas, cand);
  if (cand.empty())
    return fail(DMPC_ERR_NO_MOVE, "no vertex admits a nearest-neighbour interchange here (the reference calls "
                                  "gsl_rng_uniform_int(r, 0), clusterFunArmadillo.cpp:187)");
  t->ht.beginNni();
  int lastRoot = -1;
  for (int m = 0; m < numMoves; m++) {
    const int root = cand[t->ht.rng.uniform_int(cand.size())];
    int two[2];
    t->ht.twoVerticesForNNI(root, mrcas, two);
    t->ht.rearrangeNNI(two[0], two[1]);
    lastRoot = root;
  }
  int rc = evaluate(t, w, b, pi, ll);
  if (rc == DMPC_OK) t->ht.exportProposed(edgeOut, numMoves == 1 ? lastRoot : -1);
  return rc;
}

int dmpc_within_clus_nni_loglik(dmpc_tree* t, const double* withinTransMats, const double* betweenTransMats,
                                int mrcaOfClusForNNI, int numMovesNNI, const double* limProbs, double* logLikOut,
                                int32_t* edgeOut) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!withinTransMats || !betweenTransMats || !limProbs || !logLikOut || !edgeOut) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->grp) {
    // every shard draws the same NNI picks from its own Tausworthe stream and exports the same edge matrix
    std::vector<std::vector<int32_t>> e2(t->grp->shards.size());
    return groupLoglik(t->grp, logLikOut, [&](int i, dmpc_tree* sh, double* ll) {
      if (i) e2[i].resize((size_t)4 * (sh->N - 1));
      return dmpc_within_clus_nni_loglik(sh, withinTransMats, betweenTransMats, mrcaOfClusForNNI, numMovesNNI, limProbs, ll, i ? e2[i].data() : edgeOut);
    });
  }
  if (mrcaOfClusForNNI < 1 || mrcaOfClusForNNI > t->ht.numVerts) return fail(DMPC_ERR_ARG, "MRCAofClusForNNI out of range");
  return nniCommon(t, withinTransMats, betweenTransMats, mrcaOfClusForNNI - 1, nullptr, numMovesNNI, limProbs, logLikOut, edgeOut);
}

int dmpc_between_clus_nni_loglik(dmpc_tree* t, const double* withinTransMats, const double* betweenTransMats,
                                 const double* clusterMRCAs, int nClusterMRCAs, int numMovesNNI,
                                 const double* limProbs, double* logLikOut, int32_t* edgeOut) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!withinTransMats || !betweenTransMats || !limProbs || !logLikOut || !edgeOut || (nClusterMRCAs > 0 && !clusterMRCAs))
    return fail(DMPC_ERR_ARG, "null pointer");
  if (t->grp) {
    std::vector<std::vector<int32_t>> e2(t->grp->shards.size());
    return groupLoglik(t->grp, logLikOut, [&](int i, dmpc_tree* sh, double* ll) {
      if (i) e2[i].resize((size_t)4 * (sh->N - 1));
      return dmpc_between_clus_nni_loglik(sh, withinTransMats, betweenTransMats, clusterMRCAs, nClusterMRCAs, numMovesNNI, limProbs, ll,
                                          i ? e2[i].data() : edgeOut);
    });
  }
  std::vector<long> m(nClusterMRCAs);
  for (int i = 0; i < nClusterMRCAs; i++) m[i] = (long)clusterMRCAs[i];
  return nniCommon(t, withinTransMats, betweenTransMats, t->ht.root(), &m, numMovesNNI, limProbs, logLikOut, edgeOut);
}

int dmpc_clus_split_merge_loglik(dmpc_tree* t, const double* withinTransMats, const double* betweenTransMats,
                                 const int32_t* mrcas, int n, const double* limProbs, double* logLikOut) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!withinTransMats || !betweenTransMats || !limProbs || !logLikOut || !mrcas || n < 1) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->grp)
    return groupLoglik(t->grp, logLikOut, [&](int, dmpc_tree* sh, double* ll) {
      return dmpc_clus_split_merge_loglik(sh, withinTransMats, betweenTransMats, mrcas, n, limProbs, ll);
    });
  HostTree& ht = t->ht;
  for (int i = 0; i < n; i++)
    if (mrcas[i] < 1 || mrcas[i] > ht.numVerts) return fail(DMPC_ERR_ARG, "cluster MRCA out of range");
  if (n == 1 && mrcas[0] <= ht.numTips)
    return fail(DMPC_ERR_ARG, "cannot split a tip (the reference dereferences a null child, AugTree.cpp:332)");
  if (n > 1 && ht.parent[mrcas[0] - 1] < 0) return fail(DMPC_ERR_ARG, "cannot merge the root");
  ht.negateAllUpdateFlags();
  if (n == 1) {                                   // HandleSplit, AugTree.cpp:328-336
    const int m = mrcas[0] - 1;
    ht.invalidateSolution(m);
    ht.within[ht.kids[2 * m]] = 0; ht.within[ht.kids[2 * m + 1]] = 0;
  } else {                                        // HandleMerge, AugTree.cpp:338-346
    ht.invalidateSolution(ht.parent[mrcas[0] - 1]);
    for (int i = 0; i < n; i++) ht.within[mrcas[i] - 1] = 1;
  }
  return evaluate(t, withinTransMats, betweenTransMats, limProbs, logLikOut);
}

static int rollbackTouched(dmpc_tree* t) {
  HostTree& ht = t->ht;
  if (t->pendingCur) {                            // two rollbacks in a row: the pinned list is still waiting to be uploaded
    int rc = flushCur(t);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(t->stream));
  }
  int n = 0;
  ht.rollbackTouched([&](int v, int slot) { t->h_curList[n++] = ((v - ht.numTips) << 1) | slot; });
  t->pendingCur = n;
  return DMPC_OK;
}

int dmpc_restore_previous_config(dmpc_tree* t, const int32_t* edge, int nEdgeRows, int withinMatListIndex,
                                 int betweenMatListIndex, int nniMove, const int32_t* clusterMRCAs,
                                 int nClusterMRCAs, int splitMergeMove) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (t->grp)
    return groupRun(t->grp, [&](int, dmpc_tree* sh) {
      return dmpc_restore_previous_config(sh, edge, nEdgeRows, withinMatListIndex, betweenMatListIndex, nniMove, clusterMRCAs, nClusterMRCAs,
                                          splitMergeMove);
    });
  if ((rc = ensureDevice(t))) return rc;
  HostTree& ht = t->ht;
  ht.withinIdx = withinMatListIndex; ht.betweenIdx = betweenMatListIndex;
  if (nniMove) {                                  // BuildTreeNoAssign, AugTree.cpp:88-101
    if (!edge || nEdgeRows != ht.numVerts - 1) return fail(DMPC_ERR_ARG, "edge matrix required for an NNI rollback");
    if (!ht.restoreTopology(edge, nEdgeRows)) return fail(DMPC_ERR_ARG, ht.err);
  }
  if ((rc = rollbackTouched(t))) return rc;
  if (splitMergeMove) {
    if (nClusterMRCAs > 0 && !clusterMRCAs) return fail(DMPC_ERR_ARG, "null pointer");
    ht.associate(clusterMRCAs, nClusterMRCAs);
  }
  return DMPC_OK;
}

int dmpc_final_deallocate(dmpc_tree* t) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (t->grp) destroyGroup(t); else destroy(t);
  return DMPC_OK;
}

double dmpc_get_max_map_size_estimate(double allowedSizeInMegs, unsigned numRateCats) {
  // sizeof(S) = 32 and sizeof(mapContentType) = 24 on LP64/libstdc++ (clusterFunArmadillo.cpp:300)
  return std::floor((allowedSizeInMegs * 1e6) / (32.0 + 24.0 * numRateCats + 16.0));
}

int dmpc_check_and_cull_map(dmpc_tree* t, unsigned, float, const double*, const double*, const double*) {
  return checkHandle(t);   // dense HBM storage: nothing to cull, results never depend on culling
}

int dmpc_recompute_all(dmpc_tree* t, const double* withinTransMats, const double* betweenTransMats,
                       const double* limProbs, double* logLikOut) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!withinTransMats || !betweenTransMats || !limProbs || !logLikOut) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->grp)
    return groupLoglik(t->grp, logLikOut, [&](int, dmpc_tree* sh, double* ll) { return dmpc_recompute_all(sh, withinTransMats, betweenTransMats, limProbs, ll); });
  t->ht.negateAllUpdateFlags();
  t->ht.invalidateAll();
  return evaluate(t, withinTransMats, betweenTransMats, limProbs, logLikOut);
}

// Batch of split / merge proposals against the committed state (no reference counterpart).
int dmpc_eval_proposals(dmpc_tree* t, const double* withinTransMats, const double* betweenTransMats, const double* limProbs,
                        const dmpc_proposal* props, int nProps, double* logLikOut, int* nBatchedOut) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (t->grp) return fail(DMPC_ERR_STATE, "dmpc_eval_proposals is not available on a multi-device handle");
  if ((rc = ensureDevice(t))) return rc;
  if (!withinTransMats || !betweenTransMats || !limProbs || !props || !logLikOut || nProps < 0) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->deferred) return fail(DMPC_ERR_STATE, "dmpc_eval_proposals is not available on a site-sharded (deferred) handle");
  HostTree& ht = t->ht;
  if (!ht.solved[ht.root()]) return fail(DMPC_ERR_STATE, "the committed state has pending invalidations");
  for (int i = 0; i < nProps; i++) {
    const dmpc_proposal& q = props[i];
    if (q.kind == DMPC_PROPOSAL_SPLIT) {
      if (q.a <= ht.numTips || q.a > ht.numVerts) return fail(DMPC_ERR_ARG, "split: the cluster MRCA must be an internal vertex");
    } else if (q.kind == DMPC_PROPOSAL_MERGE) {
      if (q.a < 1 || q.a > ht.numVerts || q.b < 1 || q.b > ht.numVerts || q.a == q.b || ht.parent[q.a - 1] < 0 ||
          ht.parent[q.a - 1] != ht.parent[q.b - 1])
        return fail(DMPC_ERR_ARG, "merge: the two cluster MRCAs must be siblings");
    } else {
      return fail(DMPC_ERR_ARG, "unknown proposal kind");
    }
  }
  // the matrices live in module-level __constant__ memory: no other handle may launch on this device meanwhile
  std::unique_lock<std::mutex> deviceLock(g_deviceMutex[t->device & 63]);
  claimConstants(t);
  if ((rc = uploadConstants(t, withinTransMats, betweenTransMats, limProbs))) return rc;
  if ((rc = flushCur(t))) return rc;
  int nBatched = 0;
  std::vector<uint8_t> needSeq(nProps, 1);
  // Batched path: per proposal, the path to the root is walked with the partial in registers (k_prune<false, true>), then the
  // ordinary root reduction runs on per-proposal copies of its work arrays, reading the proposal's exponents for path
  // vertices and the committed ones for everything else.  Exact-mode handles (fused clades that may rescale) and
  // proposals that trip the fused-clade test or overflow the exponent lists go one by one below.
  const bool chainMode = t->quirk && t->R > 1;
  const bool batchable = !t->exact && nProps > 0;
  if (batchable) {
    // the committed state's per-site exponent lists (with the vertex of every row), which the batched gather merges each
    // proposal's path into: the handle's own lists are those of the LAST evaluation — possibly a rejected proposal — so
    // they are rebuilt here (one ordinary gather over the committed state; nothing is published, nothing is read back)
    if ((rc = launchGather(t, chainMode ? 1 : 0, false))) return rc;
    constexpr int JCAP = 96;                       // longest path handled in a batch
    const size_t Sp = (size_t)t->Sp, Kc = (size_t)t->rv.Kcap, RP = (size_t)t->rv.RP, R = (size_t)t->R, T = (size_t)t->T;
    HostTree::Plan pp;
    std::vector<int> runStart;
    std::vector<int32_t> ids, pos;
    std::vector<DevStatus> stat;
    std::vector<int> invalid;
    int b0 = 0;
    while (b0 < nProps) {
      // plan proposals until the batch's scratch would exceed ~3 GB
      pp.tasks.clear(); pp.tokens.clear(); pp.tipList.clear();
      runStart.assign(1, 0);
      std::vector<int> member;                     // indices (into props) of this batch
      int jmax = 1;
      for (int i = b0; i < nProps; i++) {
        const dmpc_proposal& q = props[i];
        const size_t before = pp.tasks.size();
        ht.proposalRun(q.kind == DMPC_PROPOSAL_MERGE, q.a, q.b, pp);
        const int len = (int)(pp.tasks.size() - before);
        if (len > JCAP) { pp.tasks.resize(before); b0 = i + 1; if (member.empty()) continue; else { b0 = i; break; } }
        const int jm = std::max(jmax, len);
        const size_t per = (size_t)jm * R * Sp * 4 + (size_t)jm * R * T + R * Sp * 8 +            // pe, pf, rootdot
                           Sp * Kc * RP * 4 + Sp * (4 * 4 + 4 + 8) + Sp * RP * 4 + T * 8 + 256;    // root work arrays
        if (!member.empty() && (per * (member.size() + 1) > ((size_t)3 << 30) || member.size() >= 65535)) { pp.tasks.resize(before); b0 = i; break; }
        jmax = jm;
        member.push_back(i);
        runStart.push_back((int)pp.tasks.size());
        b0 = i + 1;
      }
      const int nb = (int)member.size();
      if (!nb) continue;
      ids.assign((size_t)nb * jmax, INT32_MAX); pos.assign((size_t)nb * jmax, 0);
      for (int i = 0; i < nb; i++) {
        const int len = runStart[i + 1] - runStart[i];
        std::vector<std::pair<int32_t, int32_t>> pr(len);
        for (int j = 0; j < len; j++) pr[j] = {pp.tasks[runStart[i] + j].out, j};
        std::sort(pr.begin(), pr.end());
        for (int j = 0; j < len; j++) { ids[(size_t)i * jmax + j] = pr[j].first; pos[(size_t)i * jmax + j] = pr[j].second; }
      }
      // scratch (stream-ordered pool)
      std::vector<void*> tmp;
      auto dalloc = [&](size_t n) -> void* { void* q = nullptr; if (cudaMallocAsync(&q, std::max<size_t>(n, 16), t->stream) != cudaSuccess) return nullptr; tmp.push_back(q); return q; };
      const size_t nT = pp.tasks.size(), nK = pp.tokens.size(), nL = pp.tipList.size();
      Task* dT = (Task*)dalloc(nT * sizeof(Task)); Token* dK = (Token*)dalloc(nK * sizeof(Token)); int32_t* dL = (int32_t*)dalloc(nL * 4);
      int* dR = (int*)dalloc((size_t)(nb + 1) * 4); int* dI = (int*)dalloc((size_t)nb * 4);
      PathView pv{};
      pv.JMAX = jmax;
      int32_t* dIds = (int32_t*)dalloc(ids.size() * 4); int32_t* dPos = (int32_t*)dalloc(pos.size() * 4);
      pv.ids = dIds; pv.pos = dPos;
      pv.pe = (float*)dalloc((size_t)nb * jmax * R * Sp * 4);
      pv.pf = (uint8_t*)dalloc((size_t)nb * jmax * R * T);
      pv.rootdot = (double*)dalloc((size_t)nb * R * Sp * 8);
      RootView rb = t->rv;
      rb.comm.world = 1;
      rb.baseElist = t->rv.elist; rb.baseEid = t->rv.eid; rb.baseKmax = t->rv.kmax;
      rb.elist = (float*)dalloc((size_t)nb * Sp * Kc * RP * 4);
      rb.kmax = (int*)dalloc((size_t)nb * Sp * 4);
      rb.act = (int*)dalloc((size_t)nb * Sp * 4);
      rb.rowoff = (int*)dalloc((size_t)nb * (Sp + 1) * 4);
      rb.posIncl = (int*)dalloc((size_t)nb * Sp * 4);
      rb.chunkStart = (int*)dalloc((size_t)nb * (Sp + 2) * 4);
      rb.cout = (float*)dalloc((size_t)nb * Sp * RP * 4);
      rb.mxo = (float*)dalloc((size_t)nb * Sp * 4);
      rb.carryOut = (float*)dalloc((size_t)nb * RMAX * 4);
      rb.sitell = (double*)dalloc((size_t)nb * Sp * 8);
      rb.chunk = (double*)dalloc((size_t)nb * T * CPT * 8);
      rb.tileRec = (double*)dalloc((size_t)nb * T * CPT * xrec(t->rv.RP) * 8);
      rb.certDom = (int*)dalloc((size_t)nb * T * CPT * 4);
      rb.st = (DevStatus*)dalloc((size_t)nb * sizeof(DevStatus));
      bool okAlloc = true;
      for (void* q : tmp) okAlloc &= q != nullptr;
      if (!okAlloc || tmp.size() != 24) { for (void* q : tmp) if (q) cudaFreeAsync(q, t->stream); return fail(DMPC_ERR_CUDA, "out of device memory for the proposal batch"); }
      CUDA_TRY(cudaMemcpyAsync(dT, pp.tasks.data(), nT * sizeof(Task), cudaMemcpyHostToDevice, t->stream));
      if (nK) CUDA_TRY(cudaMemcpyAsync(dK, pp.tokens.data(), nK * sizeof(Token), cudaMemcpyHostToDevice, t->stream));
      if (nL) CUDA_TRY(cudaMemcpyAsync(dL, pp.tipList.data(), nL * 4, cudaMemcpyHostToDevice, t->stream));
      CUDA_TRY(cudaMemcpyAsync(dR, runStart.data(), (size_t)(nb + 1) * 4, cudaMemcpyHostToDevice, t->stream));
      CUDA_TRY(cudaMemcpyAsync(dIds, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, t->stream));
      CUDA_TRY(cudaMemcpyAsync(dPos, pos.data(), pos.size() * 4, cudaMemcpyHostToDevice, t->stream));
      CUDA_TRY(cudaMemsetAsync(dI, 0, (size_t)nb * 4, t->stream));
      CUDA_TRY(cudaMemsetAsync(rb.st, 0, (size_t)nb * sizeof(DevStatus), t->stream));
      const int mode = chainMode ? 1 : 0;
      {
        PruneArgs a{};
        a.tasks = dT; a.prog = dK; a.tipList = dL; a.runStart = dR; a.invalid = dI; a.pv = pv;
        k_prune<false, true, 2, 5><<<dim3((unsigned)t->R, (unsigned)t->T, (unsigned)nb), TILE / 2, 0, t->stream>>>(t->dv, a);
      }
      if (t->rv.RP == 4) k_root_gather<4, true><<<dim3((unsigned)t->T, (unsigned)nb), TILE, 0, t->stream>>>(t->dv, rb, mode, pv);
      else k_root_gather<8, true><<<dim3((unsigned)t->T, (unsigned)nb), TILE, 0, t->stream>>>(t->dv, rb, mode, pv);
      // root stage per proposal: quiet proposals were finished by the gather; the others take the certificate (mode 2) —
      // and, if for some proposal it does not hold, the literal chain + final over the batch (mode 1) afterwards
      stat.resize(nb); invalid.resize(nb);
      if (chainMode) {
        k_final<true><<<dim3((unsigned)t->T, (unsigned)nb), TILE, 0, t->stream>>>(t->dv, rb, 2, 1, 0, pv);
        t->st.kernel_launches += 3;
      } else {
        k_final<true><<<dim3((unsigned)t->T, (unsigned)nb), TILE, 0, t->stream>>>(t->dv, rb, 0, 1, 0, pv);
        t->st.kernel_launches += 3;
      }
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaMemcpyAsync(stat.data(), rb.st, (size_t)nb * sizeof(DevStatus), cudaMemcpyDeviceToHost, t->stream));
      CUDA_TRY(cudaMemcpyAsync(invalid.data(), dI, (size_t)nb * 4, cudaMemcpyDeviceToHost, t->stream));
      CUDA_TRY(cudaStreamSynchronize(t->stream));    // also: the pageable host vectors above were the copy sources
      bool needLiteral = false;
      for (int i = 0; i < nb; i++) needLiteral |= chainMode && stat[i].nact != 0 && stat[i].cert == CERT_LITERAL && stat[i].overflow <= t->rv.Kcap;
      if (needLiteral) {
        if ((rc = launchChain(t, &rb, nb))) return rc;
        k_final<true><<<dim3((unsigned)t->T, (unsigned)nb), TILE, 0, t->stream>>>(t->dv, rb, 1, 1, 0, pv);
        t->st.kernel_launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(stat.data(), rb.st, (size_t)nb * sizeof(DevStatus), cudaMemcpyDeviceToHost, t->stream));
        CUDA_TRY(cudaStreamSynchronize(t->stream));
      }
      for (void* q : tmp) cudaFreeAsync(q, t->stream);
      for (int i = 0; i < nb; i++)
        if (!invalid[i] && stat[i].overflow <= t->rv.Kcap) {
          logLikOut[member[i]] = stat[i].ll;
          needSeq[member[i]] = 0;
          nBatched++;
        }
      t->st.updates += (unsigned long long)nT * t->S * t->R;
    }
  }
  // whatever could not be batched (rescaling involved, exact mode): the ordinary proposal path, then an internal reject.
  // These evaluations reuse the rollback slots, which is why the call must not sit between a proposal and its
  // dmpc_restore_previous_config (see the header).
  deviceLock.unlock();                             // evaluate() takes it again
  for (int i = 0; i < nProps; i++) {
    if (!needSeq[i]) continue;
    const dmpc_proposal& q = props[i];
    const std::vector<uint8_t> savedWithin = ht.within;
    const int savedNact = t->lastNact;
    ht.negateAllUpdateFlags();
    if (q.kind == DMPC_PROPOSAL_SPLIT) {
      const int m = q.a - 1;
      ht.invalidateSolution(m);
      ht.within[ht.kids[2 * m]] = 0; ht.within[ht.kids[2 * m + 1]] = 0;
    } else {
      ht.invalidateSolution(ht.parent[q.a - 1]);
      ht.within[q.a - 1] = 1; ht.within[q.b - 1] = 1;
    }
    if ((rc = evaluate(t, withinTransMats, betweenTransMats, limProbs, &logLikOut[i]))) return rc;
    if ((rc = rollbackTouched(t))) return rc;
    ht.within = savedWithin;
    t->lastNact = savedNact;
  }
  if (nBatchedOut) *nBatchedOut = nBatched;
  return DMPC_OK;
}

int dmpc_get_stats(dmpc_tree* t, dmpc_stats* out) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!out) return fail(DMPC_ERR_ARG, "null pointer");
  t->timingOn = true;
  if (t->grp) {
    // counters add up over the shards; times are those of the slowest shard; shapes are the whole alignment's
    Group* g = t->grp;
    for (size_t i = 0; i < g->shards.size(); i++) {
      dmpc_stats s1;
      if ((rc = dmpc_get_stats(g->shards[i], &s1))) return rc;
      if (i == 0) { *out = s1; continue; }
      out->updates += s1.updates; out->kernel_launches += s1.kernel_launches; out->last_updates += s1.last_updates;
      out->device_bytes += s1.device_bytes; out->last_algorithmic_bytes += s1.last_algorithmic_bytes; out->last_fp64_ops += s1.last_fp64_ops;
      out->last_active_tiles += s1.last_active_tiles; out->last_chain_sites += s1.last_chain_sites; out->last_chain_rows += s1.last_chain_rows;
      out->last_eval_ms = std::max(out->last_eval_ms, s1.last_eval_ms); out->last_prune_ms = std::max(out->last_prune_ms, s1.last_prune_ms);
      out->prune_ms_total = std::max(out->prune_ms_total, s1.prune_ms_total);
      out->last_host_total_us = std::max(out->last_host_total_us, s1.last_host_total_us);
      out->create_upload_ms = std::max(out->create_upload_ms, s1.create_upload_ms);
      out->exponent_capacity = std::max(out->exponent_capacity, s1.exponent_capacity);
      out->exact_mode |= s1.exact_mode;
    }
    out->num_loci = g->S;
    return DMPC_OK;
  }
  if (t->evalTimePending) {                    // the evaluation's result was polled: its closing event may complete a moment later
    t->evalTimePending = false;
    float ms = 0;
    if (ensureDevice(t) == DMPC_OK && cudaEventSynchronize(t->ev[2]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, t->ev[0], t->ev[2]) == cudaSuccess)
      t->st.last_eval_ms = ms;
  }
  *out = t->st;
  out->exact_mode = t->exact ? 1 : 0;
  out->chain_group = t->chainGroup;
  out->plan_cache_hits = t->planCacheHits; out->graph_replays = t->graphReplays;
  out->num_tips = t->N; out->num_loci = t->S; out->num_rate_cats = t->R;
  return DMPC_OK;
}

int dmpc_get_root_timeline(dmpc_tree* t, double* usOut) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!usOut) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->grp) t = t->grp->shards[0];
  const unsigned long long* s = t->h_res->stamp;
  auto d = [&](int a, int b) { return (s[a] == ~0ull || s[b] == ~0ull || s[b] < s[a]) ? -1.0 : (double)(s[b] - s[a]) * 1e-3; };
  usOut[0] = d(0, 1);   // gather: first block -> last block found
  usOut[1] = d(1, 2);   // exchange (wait for the peers) / quiet reduction
  usOut[2] = d(2, 3);   // certificate (+ literal prefix)
  usOut[3] = t->seqPredicted || t->lastCert ? d(3, 4) : -1.0;   // gap until the final kernel's first block
  usOut[4] = t->lastCert ? d(4, 5) : -1.0;   // final: per-site terms
  usOut[5] = t->lastCert ? d(5, 6) : -1.0;   // final's exchange / reduction
  return DMPC_OK;
}

int dmpc_timer_begin(dmpc_tree* t) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (t->grp) { for (dmpc_tree* sh : t->grp->shards) if ((rc = dmpc_timer_begin(sh))) return rc; return DMPC_OK; }
  if ((rc = ensureDevice(t))) return rc;
  CUDA_TRY(cudaStreamSynchronize(t->stream));
  CUDA_TRY(cudaEventRecord(t->ev[3], t->stream));
  return DMPC_OK;
}

int dmpc_timer_end(dmpc_tree* t, double* msOut) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!msOut) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->grp) {                                    // the slowest shard's device time
    double worst = 0.0;
    for (dmpc_tree* sh : t->grp->shards) { double ms; if ((rc = dmpc_timer_end(sh, &ms))) return rc; worst = std::max(worst, ms); }
    *msOut = worst;
    return DMPC_OK;
  }
  if ((rc = ensureDevice(t))) return rc;
  CUDA_TRY(cudaEventRecord(t->ev[4], t->stream));
  CUDA_TRY(cudaEventSynchronize(t->ev[4]));
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, t->ev[3], t->ev[4]));
  *msOut = ms;
  return DMPC_OK;
}

int dmpc_plan_selfcheck(const int32_t* edge, int nEdgeRows, int numTips, const double* clusterMRCAs, int nClusterMRCAs,
                        int fuseTips, int dirtyVertex, int* out) {
  if (!edge || !out || (nClusterMRCAs > 0 && !clusterMRCAs)) return fail(DMPC_ERR_ARG, "null pointer");
  HostTree ht;
  if (!ht.build(edge, nEdgeRows, numTips)) return fail(DMPC_ERR_ARG, ht.err);
  if (dirtyVertex > 0 && (dirtyVertex <= numTips || dirtyVertex > ht.numVerts)) return fail(DMPC_ERR_ARG, "dirtyVertex must be internal");
  ht.associate(clusterMRCAs, nClusterMRCAs);
  ht.fuseTips = std::max(1, std::min(fuseTips, (int)FUSE_MAX_TIPS));
  {
    // tip counts: the one-pass version of link() (pre-ordered rows) and the incremental update of an NNI must agree with
    // the full traversal
    auto recount = [&](const char* what) -> bool {
      if (!ht.tipsValid) return true;
      const std::vector<int32_t> fast = ht.tips;
      ht.tipsValid = false;
      ht.computeTips();
      if (fast != ht.tips) { fail(DMPC_ERR_STATE, std::string("plan self-check: tip counts after ") + what); return false; }
      return true;
    };
    if (!recount("link")) return DMPC_ERR_STATE;
    HostTree sc = ht;                              // a scratch copy takes a few NNI moves (RearrangeTreeNNI, AugTree.cpp:211-226)
    std::vector<int> cand;
    sc.nniCandidates(sc.root(), nullptr, cand);
    for (int m = 0; m < 8 && !cand.empty(); m++) {
      int two[2];
      sc.twoVerticesForNNI(cand[sc.rng.uniform_int(cand.size())], nullptr, two);
      sc.rearrangeNNI(two[0], two[1]);
      const std::vector<int32_t> fast = sc.tips;
      const bool was = sc.tipsValid;
      sc.tipsValid = false;
      sc.computeTips();
      if (was && fast != sc.tips) return fail(DMPC_ERR_STATE, "plan self-check: tip counts after an NNI move");
      cand.clear();
      sc.nniCandidates(sc.root(), nullptr, cand);
    }
  }
  HostTree::Plan p;
  ht.plan(p);
  const bool full = dirtyVertex == 0;
  std::vector<uint8_t> dirty(ht.numVerts, 0);
  if (!full) {                                    // a proposal on top of the evaluated tree
    ht.negateAllUpdateFlags();
    if (dirtyVertex > 0) {                        // split / merge: the path from dirtyVertex to the root (InvalidateSolution)
      ht.invalidateSolution(dirtyVertex - 1);
    } else if (dirtyVertex == -1) {               // new between-cluster matrices (AugTree.cpp:273-287)
      ht.invalidateBetween();
    } else {                                      // one NNI move anywhere in the tree: two root paths that meet (AugTree.cpp:211-226)
      for (int i = 0; i < -dirtyVertex - 2; i++) ht.rng.next();      // -2, -3, ...: different picks from the Tausworthe stream
      std::vector<int> cand;
      ht.nniCandidates(ht.root(), nullptr, cand);
      if (!cand.empty()) {
        int two[2];
        ht.twoVerticesForNNI(cand[ht.rng.uniform_int(cand.size())], nullptr, two);
        ht.rearrangeNNI(two[0], two[1]);
      }
    }
    for (int v = numTips; v < ht.numVerts; v++) dirty[v] = !ht.solved[v];   // what TrySolve recomputes (AugTree.cpp:114-134)
    ht.plan(p);
  }
  const int N = numTips, nInt = N - 1;
  auto bad = [&](const std::string& m) { return fail(DMPC_ERR_STATE, "plan self-check: " + m); };
  std::vector<int> computedAt(nInt, -1);          // launch unit in which each internal vertex is computed
  std::vector<int> computedSeg(nInt, -1), computedPos(nInt, -1), segOfTask(p.tasks.size(), -1);
  // segments: contiguous, covering exactly the tasks below the chain, each inside one launch unit
  {
    const int nUnits = (int)p.levelStart.size() - 1 - p.chainLevels;
    if ((int)p.unitSeg.size() != nUnits + 1 || p.segStart.empty() || p.segStart.front() != 0) return bad("segment table shape");
    for (int u = 0; u < nUnits; u++) {
      if (p.unitSeg[u + 1] < p.unitSeg[u] || p.unitSeg[u + 1] >= (int)p.segStart.size()) return bad("unit -> segment range");
      if (p.segStart[p.unitSeg[u]] != p.levelStart[u] || p.segStart[p.unitSeg[u + 1]] != p.levelStart[u + 1]) return bad("segments do not tile their launch unit");
      for (int sg = p.unitSeg[u]; sg < p.unitSeg[u + 1]; sg++) {
        if (p.segStart[sg + 1] <= p.segStart[sg]) return bad("empty segment");
        for (int i = p.segStart[sg]; i < p.segStart[sg + 1]; i++) segOfTask[i] = sg;
      }
    }
  }
  std::vector<uint8_t> asTask(nInt, 0), readStored(nInt, 0);
  const int nLevels = (int)p.levelStart.size() - 1;
  int maxDepth = 0, maxTips = 0, fusedCount = 0;
  for (int l = 0; l < nLevels; l++)
    for (int i = p.levelStart[l]; i < p.levelStart[l + 1]; i++) {
      const Task& k = p.tasks[i];
      if (k.out < 0 || k.out >= nInt) return bad("task output out of range");
      if (computedAt[k.out] != -1) return bad("vertex computed twice");
      const int n0 = k.ntok & 0xffff, n1 = (k.ntok >> 16) & 0xffff;
      if (((k.flags & TF_C0FUSED) != 0) != (n0 > 0) || ((k.flags & TF_C1FUSED) != 0) != (n1 > 0)) return bad("fused flag without program");
      if (k.prog < 0 || k.prog + n0 + n1 > (int)p.tokens.size()) return bad("token range");
      if (k.tipOff < 0 || k.tipOff + k.nTips > (int)p.tipList.size()) return bad("tip list range");
      if (k.nTips > 2 * FUSE_MAX_TIPS || n0 + n1 > 3 * FUSE_MAX_TIPS + 4) return bad("task exceeds the kernel's shared-memory staging");
      maxTips = std::max(maxTips, k.nTips);
      for (int c = 0; c < 2; c++) {               // stored operands must have been computed at a lower level, or be clean
        const int idx = c ? k.c1 : k.c0;
        const bool tip = k.flags & (c ? TF_C1TIP : TF_C0TIP), fus = k.flags & (c ? TF_C1FUSED : TF_C0FUSED);
        if (tip) { if (idx < 0 || idx >= N) return bad("tip operand out of range"); continue; }
        if (idx < 0 || idx >= nInt) return bad("operand out of range");
        if (ht.kids[2 * (k.out + N) + c] != idx + N) return bad("operand is not the vertex's child");
        if (fus) continue;
        readStored[idx] = 1;
        // a stored operand comes from an earlier launch, or from an earlier task of the SAME segment (one block walks it)
        if (computedAt[idx] > l || (computedAt[idx] == l && !(computedSeg[idx] == segOfTask[i] && segOfTask[i] >= 0 && computedPos[idx] < i)))
          return bad("stored operand not computed before its reader");
        if (computedAt[idx] < 0 && (full || dirty[idx + N])) return bad("stored operand is dirty but not part of the plan");
        ht.computeTips();
        if (computedAt[idx] < 0 && ht.fused(idx + N)) return bad("a fused-size operand is read from HBM without having been materialised");
      }
      // run the programs symbolically: stack discipline, tips under each clade, every token's vertex new
      int off = k.prog;
      for (int c = 0; c < 2; c++) {
        const int n = c ? n1 : n0;
        if (!n) continue;
        int sp = c && n0 ? 1 : 0, tipsSeen = 0;
        bool haveReg = false;
        for (int j = off; j < off + n; j++) {
          const Token& t = p.tokens[j];
          const int op = t.op & 7;
          if (op == OP_SPILL) { if (!haveReg) return bad("spill without a value"); sp++; haveReg = false; maxDepth = std::max(maxDepth, sp); continue; }
          if (t.vertex < 0 || t.vertex >= nInt || computedAt[t.vertex] != -1) return bad("fused vertex out of range or computed twice");
          if (op == OP_CHERRY) { if (t.a < 0 || t.a >= k.nTips || t.b < 0 || t.b >= k.nTips) return bad("cherry tip index"); tipsSeen += 2; }
          else if (op == OP_REGTIP) { if (!haveReg || t.b < 0 || t.b >= k.nTips) return bad("reg-tip without a value"); tipsSeen += 1; }
          else if (op == OP_SLOTREG) { if (!haveReg || sp <= (c && n0 ? 1 : 0)) return bad("slot-reg without a spilled value"); sp--; }
          else return bad("unknown opcode");
          haveReg = true;
          computedAt[t.vertex] = l;
          fusedCount++;
        }
        if (sp != (c && n0 ? 1 : 0) || !haveReg) return bad("unbalanced program");
        if (tipsSeen > ht.fuseTips) return bad("fused clade larger than the threshold");
        off += n;
      }
      if (maxDepth > FUSE_DEPTH) return bad("stack deeper than FUSE_DEPTH");
      computedAt[k.out] = l; computedSeg[k.out] = segOfTask[i]; computedPos[k.out] = i;
      asTask[k.out] = 1;
    }
  if (full) { for (int v = 0; v < nInt; v++) if (computedAt[v] < 0) return bad("vertex never computed"); }
  else { for (int v = N; v < ht.numVerts; v++) if (dirty[v] && computedAt[v - N] < 0) return bad("dirty vertex never computed"); }
  if (nLevels > 0 && computedAt[0] != nLevels - 1) return bad("the root is not in the last level");
  // the walked chain (k_walk): one task per trailing level, each consuming the previous one's output; no fused operands
  // (materialised beforehand: level 0 of the plan holds exactly those clades)
  if (p.chainLevels) {
    if (p.chainLevels < CHAIN_MIN || p.chainLevels > nLevels) return bad("chain length");
    const int first = nLevels - p.chainLevels;
    for (int l = first; l < nLevels; l++) {
      if (p.levelStart[l + 1] - p.levelStart[l] != 1) return bad("a chain level holds more than one task");
      const Task& k = p.tasks[p.levelStart[l]];
      if (k.flags & (TF_C0FUSED | TF_C1FUSED)) return bad("fused operand inside the chain");
      const int fw = k.flags & (TF_C0FWD | TF_C1FWD);
      if (l == first) { if (!(k.flags & TF_HEAD) || fw) return bad("chain head flags"); continue; }
      if ((k.flags & TF_HEAD) || (fw != TF_C0FWD && fw != TF_C1FWD)) return bad("chain step flags");
      const Task& pr = p.tasks[p.levelStart[l - 1]];
      if ((fw == TF_C0FWD ? k.c0 : k.c1) != pr.out) return bad("forwarded operand is not the previous task's output");
      if (fw == TF_C0FWD ? (k.flags & TF_C0TIP) : (k.flags & TF_C1TIP)) return bad("forwarded operand flagged as a tip");
    }
    for (int l = 0; l < first; l++)
      for (int i = p.levelStart[l]; i < p.levelStart[l + 1]; i++)
        if (p.tasks[i].flags & (TF_HEAD | TF_C0FWD | TF_C1FWD)) return bad("chain flags outside the chain");
    ht.computeTips();
    int nMat = 0;
    for (int v = 0; v < nInt; v++)
      if (asTask[v] && ht.fused(v + N)) { nMat++; if (computedAt[v] != 0 || !readStored[v]) return bad("materialised clade not in level 0 or never read"); }
    if (nMat != p.materialised) return bad("materialised count");
  } else {
    for (const Task& k : p.tasks) if (k.flags & (TF_HEAD | TF_C0FWD | TF_C1FWD)) return bad("chain flags without a chain");
    if (p.materialised) return bad("materialised clades without a chain");
  }
  for (int v = 0; v < ht.numVerts; v++) if (ht.mat[v]) return bad("materialisation marks left behind");
  out[0] = (int)p.tasks.size(); out[1] = fusedCount; out[2] = nLevels; out[3] = (int)p.tokens.size(); out[4] = maxDepth; out[5] = maxTips;
  out[6] = p.chainLevels; out[7] = p.materialised;
  return DMPC_OK;
}

// ---------------------------------------------------------------- host-only scheduling for the CPU test-suite
struct dmpc_hostplan {
  HostTree ht;
  HostTree::Plan plan;
  bool nniPending = false;       // dmpc_hp_propose made NNI swaps whose edge matrix dmpc_hp_state has not exported yet
  int nniRoot = -1;
};

int dmpc_hp_create(const int32_t* edge, int nEdgeRows, int numTips, const double* clusterMRCAs, int nClusterMRCAs,
                   int fuseTips, dmpc_hostplan** out) {
  if (!edge || !out || (nClusterMRCAs > 0 && !clusterMRCAs)) return fail(DMPC_ERR_ARG, "null pointer");
  dmpc_hostplan* h = new dmpc_hostplan;
  if (!h->ht.build(edge, nEdgeRows, numTips)) { std::string e = h->ht.err; delete h; return fail(DMPC_ERR_ARG, e); }
  h->ht.associate(clusterMRCAs, nClusterMRCAs);
  h->ht.fuseTips = std::max(1, std::min(fuseTips, (int)FUSE_MAX_TIPS));
  // the same scheduler knobs a GPU handle reads (experiments; the CPU suite sweeps them)
  if (const char* cm = std::getenv("DMPC_CHAIN_MIN")) h->ht.chainMin = std::max((int)CHAIN_MIN, std::atoi(cm));
  if (const char* gc = std::getenv("DMPC_GROUP_COST")) h->ht.groupCost = std::max(0, std::atoi(gc));
  *out = h;
  return DMPC_OK;
}

int dmpc_hp_propose(dmpc_hostplan* h, int kind, const int32_t* args, int nArgs) {
  if (!h || (nArgs > 0 && !args)) return fail(DMPC_ERR_ARG, "null pointer");
  HostTree& ht = h->ht;
  auto vertexOK = [&](int v1) { return v1 >= 1 && v1 <= ht.numVerts; };
  ht.negateAllUpdateFlags();
  switch (kind) {
    case 0: ht.invalidateAll(); return DMPC_OK;                       // clusterFunArmadillo.cpp:142-164
    case 1: ht.invalidateBetween(); return DMPC_OK;                   // :117-138
    case 2: {                                                         // HandleSplit, AugTree.cpp:328-336
      if (nArgs != 1 || !vertexOK(args[0]) || args[0] <= ht.numTips) return fail(DMPC_ERR_ARG, "split: one internal mrca");
      const int m = args[0] - 1;
      ht.invalidateSolution(m);
      ht.within[ht.kids[2 * m]] = 0; ht.within[ht.kids[2 * m + 1]] = 0;
      return DMPC_OK;
    }
    case 3: {                                                         // HandleMerge, AugTree.cpp:338-346
      if (nArgs < 2) return fail(DMPC_ERR_ARG, "merge: at least two mrcas");
      for (int i = 0; i < nArgs; i++) if (!vertexOK(args[i])) return fail(DMPC_ERR_ARG, "merge: mrca out of range");
      if (ht.parent[args[0] - 1] < 0) return fail(DMPC_ERR_ARG, "cannot merge the root");
      ht.invalidateSolution(ht.parent[args[0] - 1]);
      for (int i = 0; i < nArgs; i++) ht.within[args[i] - 1] = 1;
      return DMPC_OK;
    }
    case 4: case 5: {                                                 // clusterFunArmadillo.cpp:168-242
      std::vector<long> mr;
      int origin, numMoves;
      if (kind == 4) {
        if (nArgs != 2 || !vertexOK(args[0])) return fail(DMPC_ERR_ARG, "within NNI: {mrca, numMoves}");
        origin = args[0] - 1; numMoves = args[1];
      } else {
        if (nArgs < 1) return fail(DMPC_ERR_ARG, "between NNI: {numMoves, mrcas...}");
        origin = ht.root(); numMoves = args[0];
        for (int i = 1; i < nArgs; i++) mr.push_back(args[i]);
      }
      const std::vector<long>* mp = kind == 5 ? &mr : nullptr;
      std::vector<int> cand;
      ht.nniCandidates(origin, mp, cand);
      if (cand.empty()) return fail(DMPC_ERR_NO_MOVE, "no vertex admits a nearest-neighbour interchange here");
      ht.beginNni();
      h->nniRoot = -1;
      for (int m = 0; m < numMoves; m++) {
        int two[2];
        const int root = cand[ht.rng.uniform_int(cand.size())];
        ht.twoVerticesForNNI(root, mp, two);
        ht.rearrangeNNI(two[0], two[1]);
        h->nniRoot = numMoves == 1 ? root : -1;
      }
      h->nniPending = true;
      return DMPC_OK;
    }
    default: return fail(DMPC_ERR_ARG, "unknown proposal kind");
  }
}

int dmpc_hp_plan(dmpc_hostplan* h, int32_t* tasks, int taskCap, int32_t* tokens, int tokenCap, int32_t* tips, int tipCap,
                 int32_t* levelStart, int levelCap, int32_t* counts) {
  if (!h || !tasks || !tokens || !tips || !levelStart || !counts) return fail(DMPC_ERR_ARG, "null pointer");
  static_assert(sizeof(Task) == 32 && sizeof(Token) == 16, "plan records are plain int32 tuples");
  HostTree::Plan& p = h->plan;
  h->ht.plan(p);
  const int nL = (int)p.levelStart.size() - 1;
  if ((int)p.tasks.size() > taskCap || (int)p.tokens.size() > tokenCap || (int)p.tipList.size() > tipCap || nL + 1 > levelCap)
    return fail(DMPC_ERR_ARG, "dmpc_hp_plan: output capacity too small");
  if (!p.tasks.empty()) std::memcpy(tasks, p.tasks.data(), p.tasks.size() * sizeof(Task));
  if (!p.tokens.empty()) std::memcpy(tokens, p.tokens.data(), p.tokens.size() * sizeof(Token));
  if (!p.tipList.empty()) std::memcpy(tips, p.tipList.data(), p.tipList.size() * sizeof(int32_t));
  for (int l = 0; l <= nL; l++) levelStart[l] = p.levelStart[l];
  counts[0] = (int)p.tasks.size(); counts[1] = (int)p.tokens.size(); counts[2] = (int)p.tipList.size();
  counts[3] = nL; counts[4] = p.chainLevels; counts[5] = p.materialised;
  return DMPC_OK;
}

int dmpc_hp_path(dmpc_hostplan* h, int kind, int a, int b, int32_t* tasks, int taskCap, int32_t* tokens, int tokenCap,
                 int32_t* tips, int tipCap, int32_t* counts) {
  if (!h || !tasks || !tokens || !tips || !counts) return fail(DMPC_ERR_ARG, "null pointer");
  HostTree& ht = h->ht;
  if (!ht.solved[ht.root()]) return fail(DMPC_ERR_STATE, "the committed state has pending invalidations");
  if (kind == DMPC_PROPOSAL_SPLIT) {
    if (a <= ht.numTips || a > ht.numVerts) return fail(DMPC_ERR_ARG, "split: the cluster MRCA must be an internal vertex");
  } else if (kind == DMPC_PROPOSAL_MERGE) {
    if (a < 1 || a > ht.numVerts || b < 1 || b > ht.numVerts || a == b || ht.parent[a - 1] < 0 || ht.parent[a - 1] != ht.parent[b - 1])
      return fail(DMPC_ERR_ARG, "merge: the two cluster MRCAs must be siblings");
  } else {
    return fail(DMPC_ERR_ARG, "unknown proposal kind");
  }
  HostTree::Plan pp;
  ht.proposalRun(kind == DMPC_PROPOSAL_MERGE, a, b, pp);
  if ((int)pp.tasks.size() > taskCap || (int)pp.tokens.size() > tokenCap || (int)pp.tipList.size() > tipCap)
    return fail(DMPC_ERR_ARG, "dmpc_hp_path: output capacity too small");
  if (!pp.tasks.empty()) std::memcpy(tasks, pp.tasks.data(), pp.tasks.size() * sizeof(Task));
  if (!pp.tokens.empty()) std::memcpy(tokens, pp.tokens.data(), pp.tokens.size() * sizeof(Token));
  if (!pp.tipList.empty()) std::memcpy(tips, pp.tipList.data(), pp.tipList.size() * sizeof(int32_t));
  counts[0] = (int)pp.tasks.size(); counts[1] = (int)pp.tokens.size(); counts[2] = (int)pp.tipList.size();
  return DMPC_OK;
}

int dmpc_hp_restore(dmpc_hostplan* h, const int32_t* edge, int nEdgeRows, int nniMove, const int32_t* clusterMRCAs,
                    int nClusterMRCAs, int splitMergeMove) {
  if (!h) return fail(DMPC_ERR_ARG, "null pointer");
  HostTree& ht = h->ht;
  if (nniMove) {                                  // BuildTreeNoAssign, AugTree.cpp:88-101
    if (!edge || nEdgeRows != ht.numVerts - 1) return fail(DMPC_ERR_ARG, "edge matrix required for an NNI rollback");
    if (!ht.restoreTopology(edge, nEdgeRows)) return fail(DMPC_ERR_ARG, ht.err);
  }
  h->nniPending = false;
  ht.rollbackTouched([](int, int) {});            // RestoreIterVecAndExp, IntermediateNode.h:26-33
  if (splitMergeMove) {
    if (nClusterMRCAs > 0 && !clusterMRCAs) return fail(DMPC_ERR_ARG, "null pointer");
    ht.associate(clusterMRCAs, nClusterMRCAs);
  }
  return DMPC_OK;
}

int dmpc_hp_state(dmpc_hostplan* h, uint8_t* cur, uint8_t* within, int32_t* edgeOut) {
  if (!h || !cur || !within || !edgeOut) return fail(DMPC_ERR_ARG, "null pointer");
  for (int v = 0; v < h->ht.numVerts; v++) { cur[v] = h->ht.cur[v]; within[v] = h->ht.within[v]; }
  // the same incremental export the NNI entry points use (a pending NNI proposal is exported once, like they do)
  if (h->nniPending) { h->ht.exportProposed(edgeOut, h->nniRoot); h->nniPending = false; }
  else if (h->ht.pendingNni && h->ht.pendingRoot >= 0 && h->ht.exportValid) { std::copy(h->ht.committedEdge.begin(), h->ht.committedEdge.end(), edgeOut); h->ht.exportSubtree(h->ht.pendingRoot, edgeOut, false); }
  else h->ht.exportEdges(edgeOut);
  return DMPC_OK;
}

int dmpc_hp_destroy(dmpc_hostplan* h) {
  delete h;
  return DMPC_OK;
}

int dmpc_get_partial(dmpc_tree* t, int vertex, double* sol, float* expo) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (t->grp) {                                    // the shards' site blocks side by side
    Group* g = t->grp;
    if (!sol || !expo) return fail(DMPC_ERR_ARG, "null pointer");
    for (size_t i = 0; i < g->shards.size(); i++)
      if ((rc = dmpc_get_partial(g->shards[i], vertex, sol + (size_t)g->begin[i] * t->R * 4, expo + (size_t)g->begin[i] * t->R))) return rc;
    return DMPC_OK;
  }
  if ((rc = ensureDevice(t))) return rc;
  if (vertex < t->N || vertex >= t->ht.numVerts || !sol || !expo) return fail(DMPC_ERR_ARG, "vertex must be internal");
  std::unique_lock<std::mutex> deviceLock(g_deviceMutex[t->device & 63], std::defer_lock);
  if (!t->isShard) deviceLock.lock();
  claimConstants(t);
  if (!t->constantsValid && (rc = uploadConstants(t, &t->wCol[0], &t->bCol[0], t->hPi))) return rc;
  HostTree& ht = t->ht;
  const int slot = ht.cur[vertex], idx = vertex - t->N;
  int partSlot = slot, partRow = idx;
  if ((rc = flushCur(t))) return rc;
  ht.computeTips();
  if (ht.fused(vertex)) {
    // a fused vertex has no stored partial: replay its program (read-only) into the scratch row
    HostTree::Plan rp;
    const Task k = ht.makeTask(vertex, t->nReal, 0, rp, false);
    rp.tasks.assign(1, k);
    if ((rc = uploadPlan(t, rp))) return rc;
    const dim3 grid((unsigned)t->R, (unsigned)t->T, 1);
    rc = launchPrune(t, grid, t->runTasks, 1);
    if (rc) return rc;
    CUDA_TRY(cudaGetLastError());
    partSlot = 0; partRow = t->nReal;
  }
  const size_t n = (size_t)t->R * t->Sp;
  std::vector<double> hs(n * 4);
  std::vector<float> he(n);
  std::vector<uint8_t> hf((size_t)t->R * t->T);
  CUDA_TRY(cudaStreamSynchronize(t->stream));
  CUDA_TRY(cudaMemcpy(hs.data(), t->dv.part[partSlot] + (size_t)partRow * n * 4, n * 4 * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(he.data(), t->dv.expo[slot] + (size_t)idx * n, n * sizeof(float), cudaMemcpyDeviceToHost));
  for (int b = 0; b < t->R * t->T; b++)
    CUDA_TRY(cudaMemcpy(&hf[b], t->dv.flag[slot] + (size_t)b * t->nFlag + idx, 1, cudaMemcpyDeviceToHost));
  for (int s = 0; s < t->S; s++)
    for (int r = 0; r < t->R; r++) {
      for (int k = 0; k < 4; k++) sol[((size_t)s * t->R + r) * 4 + k] = hs[((size_t)r * t->Sp + s) * 4 + k];
      expo[(size_t)s * t->R + r] = hf[r * t->T + s / TILE] ? he[(size_t)r * t->Sp + s] : 0.0f;
    }
  return DMPC_OK;
}

int dmpc_get_topology(dmpc_tree* t, int32_t* parent, int32_t* child0, int32_t* child1, int32_t* within) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!parent || !child0 || !child1 || !within) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->grp) return dmpc_get_topology(t->grp->shards[0], parent, child0, child1, within);
  for (int v = 0; v < t->ht.numVerts; v++) {
    parent[v] = t->ht.parent[v];
    child0[v] = t->ht.kids[2 * v]; child1[v] = t->ht.kids[2 * v + 1];
    within[v] = t->ht.within[v];
  }
  return DMPC_OK;
}

int dmpc_get_site_logliks(dmpc_tree* t, double* out) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!out) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->grp) {
    for (size_t i = 0; i < t->grp->shards.size(); i++)
      if ((rc = dmpc_get_site_logliks(t->grp->shards[i], out + t->grp->begin[i]))) return rc;
    return DMPC_OK;
  }
  if ((rc = ensureDevice(t))) return rc;
  CUDA_TRY(cudaStreamSynchronize(t->stream));
  CUDA_TRY(cudaMemcpy(out, t->rv.sitell, (size_t)t->S * sizeof(double), cudaMemcpyDeviceToHost));
  return DMPC_OK;
}

int dmpc_set_deferred(dmpc_tree* t, int on) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (t->grp || t->isShard) return fail(DMPC_ERR_STATE, "a multi-device handle runs its own exchange");
  t->deferred = on != 0;
  t->graphVersion++;
  return DMPC_OK;
}

int dmpc_shard_chain(dmpc_tree* t, const float* carryIn, float* carryOut) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if ((rc = ensureDevice(t))) return rc;
  if (!carryIn || !carryOut) return fail(DMPC_ERR_ARG, "null pointer");
  const bool chain = t->quirk && t->R > 1;
  if (!chain) { for (int r = 0; r < t->R; r++) carryOut[r] = 0.0f; return DMPC_OK; }
  float in[RMAX] = {0};
  bool zeroIn = true;
  for (int r = 0; r < t->R; r++) { in[r] = carryIn[r]; zeroIn &= carryIn[r] == 0.0f; }
  if (t->lastNact == 0) {
    // no rescaled site in this shard: the vector passes through unchanged
    for (int r = 0; r < t->R; r++) carryOut[r] = carryIn[r];
    if (zeroIn && t->carryIsZero) return DMPC_OK;                 // the gather's optimistic chunk sums stand
    CUDA_TRY(cudaMemcpyAsync(t->rv.carryIn, in, sizeof in, cudaMemcpyHostToDevice, t->stream));
    CUDA_TRY(cudaStreamSynchronize(t->stream));                   // `in` is a stack temporary
    t->carryIsZero = zeroIn; t->chunksOnHost = false;
    return DMPC_OK;
  }
  CUDA_TRY(cudaMemcpyAsync(t->rv.carryIn, in, sizeof in, cudaMemcpyHostToDevice, t->stream));
  if ((rc = launchChain(t))) return rc;
  float out[RMAX];
  CUDA_TRY(cudaMemcpyAsync(out, t->rv.carryOut, sizeof out, cudaMemcpyDeviceToHost, t->stream));
  CUDA_TRY(cudaStreamSynchronize(t->stream));
  for (int r = 0; r < t->R; r++) carryOut[r] = out[r];
  t->carryIsZero = zeroIn; t->chunksOnHost = false;
  return DMPC_OK;
}

int dmpc_shard_active(dmpc_tree* t) { return checkHandle(t) ? -1 : (t->lastNact != 0); }

int dmpc_shard_num_chunks(dmpc_tree* t) { return checkHandle(t) ? -1 : t->nChunks; }

int dmpc_shard_finish(dmpc_tree* t, double* chunkSums) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if ((rc = ensureDevice(t))) return rc;
  if (!chunkSums) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->chunksOnHost) {                                          // the gather already finished this shard
    std::memcpy(chunkSums, t->h_chunks, (size_t)t->nChunks * sizeof(double));
    return DMPC_OK;
  }
  const bool chain = t->quirk && t->R > 1;
  if ((rc = launchFinal(t, chain ? 1 : 0, 0, false))) return rc;
  CUDA_TRY(cudaMemcpyAsync(t->h_chunks, t->rv.chunk, (size_t)t->nChunks * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
  CUDA_TRY(cudaStreamSynchronize(t->stream));
  std::memcpy(chunkSums, t->h_chunks, (size_t)t->nChunks * sizeof(double));
  return DMPC_OK;
}

int dmpc_shard_fast_result(dmpc_tree* t, int* ok, double* logLik) {
  int rc;
  if ((rc = checkHandle(t))) return rc;
  if (!ok || !logLik) return fail(DMPC_ERR_ARG, "null pointer");
  if (t->rv.comm.world > 1 && t->h_res->xok == 2) return fail(DMPC_ERR_CUDA, "peer exchange timed out: a rank of the site-sharded run never answered");
  *ok = t->rv.comm.world > 1 && t->h_res->xok == 1;
  *logLik = *ok ? t->h_res->ll : NAN;
  return DMPC_OK;
}

int dmpc_comm_create(int device, int rank, int world, int maxChunksPerRank, dmpc_comm** out, void* ipcHandle64) {
  if (!out || !ipcHandle64 || world < 1 || world > CW || rank < 0 || rank >= world || maxChunksPerRank < 1)
    return fail(DMPC_ERR_ARG, "dmpc_comm_create: world must be 1..8, rank in range, handle buffer non-null");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CUDA_TRY(cudaSetDevice(device));
  dmpc_comm* c = new dmpc_comm;
  c->device = device; c->rank = rank; c->world = world; c->maxChunks = maxChunksPerRank;
  // a message: header, chunk sums, and (when a site rescaled) one certificate record per 64-site chunk
  c->width = XHDR + maxChunksPerRank + (maxChunksPerRank + CPT) * XREC_MAX;
  if (cudaMalloc(&c->local, c->bytes() + 64) != cudaSuccess) { delete c; return fail(DMPC_ERR_CUDA, "cudaMalloc of the mailbox failed"); }
  CUDA_TRY(cudaMemset(c->local, 0, c->bytes() + 64));
  c->seq = (unsigned*)((char*)c->local + c->bytes());          // counter lives behind the mailbox proper
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, c->local));
  std::memcpy(ipcHandle64, &h, 64);
  CUDA_TRY(cudaDeviceSynchronize());
  *out = c;
  return DMPC_OK;
}

int dmpc_comm_connect(dmpc_comm* c, const void* allHandles) {
  if (!c || !allHandles) return fail(DMPC_ERR_ARG, "null pointer");
  CUDA_TRY(cudaSetDevice(c->device));
  for (int p = 0; p < c->world; p++) {
    if (p == c->rank) { c->peer[p] = c->local; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)allHandles + 64 * p, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(&c->peer[p], h, cudaIpcMemLazyEnablePeerAccess));
  }
  c->connected = true;
  return DMPC_OK;
}

int dmpc_comm_destroy(dmpc_comm* c) {
  if (!c) return DMPC_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < c->world; p++)
    if (c->ipc && p != c->rank && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
  if (c->local) cudaFree(c->local);
  delete c;
  return DMPC_OK;
}

void dmpc_shard_range(int sitesTotal, int rank, int world, int* begin, int* end) {
  int b = 0, e = 0;
  if (sitesTotal > 0 && world > 0 && rank >= 0 && rank < world) shardRange(sitesTotal, rank, world, &b, &e);
  if (begin) *begin = b;
  if (end) *end = e;
}

double dmpc_reduce_chunks(const double* chunkSums, int n) {
  volatile double acc[RED];
  for (int tI = 0; tI < RED; tI++) {
    double a = 0.0;
    for (int c = tI; c < n; c += RED) { volatile double x = a + chunkSums[c]; a = x; }
    acc[tI] = a;
  }
  for (int d = RED / 2; d > 0; d >>= 1)
    for (int tI = 0; tI < d; tI++) acc[tI] = acc[tI] + acc[tI + d];
  return acc[0];
}

}  // extern "C"
