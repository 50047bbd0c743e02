// dmpc_oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the Felsenstein-pruning likelihood path of villandre/DMphyClus
// (reference sources under /root/reference/src, cited as file:line below).  It is the
// parity checker for the CUDA path in dmphyclus_b200/csrc and the "port" CPU baseline
// of bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product never does.
//
// What is restated, line for line (dense, i.e. WITHOUT the solution dictionary of
// TreeNode.h:73-78, which is a pure cache — SURVEY.md §0.4):
//   * node kernel                      IntermediateNode.cpp:42-71  (two-step rescale, fp32 exponent)
//   * tip partials                     InputNode.cpp:3-22
//   * tree build / within flagging     AugTree.cpp:11-57, 59-101
//   * post-order solve                 AugTree.cpp:114-134
//   * root reduction                   AugTree.cpp:289-326 (fp32 exponent vector that is NOT reset
//                                      between loci, fp32 expf, vertex-id summation order)
//   * invalidation / rollback          IntermediateNode.cpp:5-13, IntermediateNode.h:26-33,
//                                      AugTree.cpp:136-142, 211-226, 273-287, 328-374
//   * NNI selection / edge export      AugTree.cpp:144-209, 228-271
//   * entry-point call sequences       clusterFunArmadillo.cpp:36-62, 117-294
//   * alphabet                         clusterFunArmadillo.cpp:64-93
//
// Third-party arithmetic that the reference takes from libraries that are absent from
// /root/reference is restated from their published algorithms:
//   * Armadillo (RcppArmadillo, version unpinned by DESCRIPTION:15): 4x4 mat*vec is
//     gemv_emul_tinysq (left-to-right row sums), dot()/accu()/mean() use two interleaved
//     accumulators, max() is exact, vec/scalar is a true division, exp(fvec) is expf.
//   * GSL gsl_rng_taus (Tausworthe, L'Ecuyer 1996) with the default seed, and
//     gsl_rng_uniform_int's rejection loop.
// Parity status: pinned against the reference's own translation units compiled here
// against a header shim (oracle/_ref, see oracle/Makefile and tests/test_oracle_vs_ref.py);
// the library arithmetic above is pinned only by this restatement.
//
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).  -ffp-contract=off
// matters: R's default x86-64 flags have no FMA, so products and sums round separately.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------- gsl_rng_taus
// GSL rng/taus.c (generator "taus"): three-component combined Tausworthe.  Default seed
// 0 is mapped to 1; seeding is an LCG(69069) chain followed by six warm-up draws.
struct TausRng {
  uint32_t s1, s2, s3;
  static uint32_t lcg(uint32_t n) { return (uint32_t)(69069u * n); }
  uint32_t get() {
#define ORC_TAUSWORTHE(s, a, b, c, d) ((((s) & (c)) << (d)) ^ ((((s) << (a)) ^ (s)) >> (b)))
    s1 = ORC_TAUSWORTHE(s1, 13, 19, 4294967294u, 12);
    s2 = ORC_TAUSWORTHE(s2, 2, 25, 4294967288u, 4);
    s3 = ORC_TAUSWORTHE(s3, 3, 11, 4294967280u, 17);
#undef ORC_TAUSWORTHE
    return s1 ^ s2 ^ s3;
  }
  void seed(unsigned long s) {
    if (s == 0) s = 1;
    s1 = lcg((uint32_t)s);
    s2 = lcg(s1);
    s3 = lcg(s2);
    for (int i = 0; i < 6; i++) get();
  }
  // gsl_rng_uniform_int (rng/rng.c): range = max - min = 0xffffffff.
  // n == 0 is a GSL error (returns 0 after the handler) — callers guard.
  unsigned long uniform_int(unsigned long n) {
    const unsigned long range = 0xffffffffUL;
    unsigned long scale = range / n, k;
    do { k = get() / scale; } while (k >= n);
    return k;
  }
};

// ---------------------------------------------------------------- node
struct Node {
  int id = 0;
  Node* parent = nullptr;
  std::vector<Node*> children;   // order matters: AddChild appends (AugTree.cpp:81-85,213-217)
  bool isTip = false;
  bool within = false;           // _withinParentBranch (TreeNode.h:121)
  bool solved = false;           // _isSolved (IntermediateNode.h:50); tips are always solved
  bool updateFlag = false;       // _updateFlag (TreeNode.h:124)
  const uint8_t* input = nullptr;          // tips: 4-bit state masks, one per locus
  // dense solutions: two buffers so that "_previousIterVec" (IntermediateNode.h:52) is a
  // pointer swap, exactly like the iterator swap in the reference.
  std::vector<double> sol[2];    // [locus][rate][4]
  std::vector<float> expo[2];    // [locus][rate]
  int cur = 0, prev = 0;
};

struct Tree {
  int numTips = 0, numLoci = 0, numRates = 0;
  int withinIdx = 0, betweenIdx = 0;
  double logLik = 0.0;
  std::vector<Node> v;           // _vertexVector: tips 0..N-1, internal N..2N-2, root = N
  std::vector<uint8_t> alignment;  // [tip][locus]
  TausRng rng;
  int threads = 1;
  bool quirkCarry = true;        // AugTree.cpp:306 exponentVec declared outside the locus loop
  long updates = 0;              // node-site-rate updates performed (for the CPU baseline)
  std::string err;
};

inline void tipSolution(const Node& n, int locus, double out[4]) {
  // InputNode.cpp:16-18: conv_to<vec>::from(indicator), identical for every rate, exponent 0
  uint8_t m = n.input[locus];
  for (int j = 0; j < 4; j++) out[j] = ((m >> j) & 1) ? 1.0 : 0.0;
}

inline void getSolution(const Tree& t, const Node& n, int locus, int rate, double out[4]) {
  if (n.isTip) { tipSolution(n, locus, out); return; }
  const double* p = &n.sol[n.cur][((size_t)locus * t.numRates + rate) * 4];
  out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = p[3];
}

inline float getExponent(const Tree& t, const Node& n, int locus, int rate) {
  if (n.isTip) return 0.0f;
  return n.expo[n.cur][(size_t)locus * t.numRates + rate];
}

// IntermediateNode.cpp:42-71 for one locus.  P: numRates matrices, column-major 4x4 (R layout),
// used as P * v (IntermediateNode.cpp:57).
void computeSolution(Tree& t, Node& n, const double* P, int locus) {
  const int R = t.numRates;
  double* outS = &n.sol[n.cur][(size_t)locus * R * 4];
  float* outE = &n.expo[n.cur][(size_t)locus * R];
  for (int r = 0; r < R; r++) {
    outS[4 * r + 0] = outS[4 * r + 1] = outS[4 * r + 2] = outS[4 * r + 3] = 1.0;  // fill::ones, :52
    outE[r] = 0.0f;
  }
  for (Node* child : n.children) {              // :53
    for (int r = 0; r < R; r++) {               // :55
      const double* A = P + 16 * r;
      double x[4];
      getSolution(t, *child, locus, r, x);
      double* s = outS + 4 * r;
      // Armadillo gemv_emul_tinysq<4>: y[i] = A(i,0)*x0 + A(i,1)*x1 + A(i,2)*x2 + A(i,3)*x3
      for (int i = 0; i < 4; i++) {
        double y = A[i] * x[0] + A[i + 4] * x[1] + A[i + 8] * x[2] + A[i + 12] * x[3];
        s[i] = s[i] * y;                        // Schur product, :57
      }
      double myMax = s[0];
      for (int i = 1; i < 4; i++) if (s[i] > myMax) myMax = s[i];   // :58
      if (myMax < 1e-150) {                     // :59-60
        for (int i = 0; i < 4; i++) s[i] = s[i] / myMax;            // :62
        outE[r] = (float)std::log(myMax);       // :63 — assignment, double log rounded to float
      }
    }
  }
}

// IntermediateNode.cpp:26-35
void computeSolutions(Tree& t, Node& n, const double* P) {
  n.prev = n.cur;                // std::copy(_dictionaryIterVec -> _previousIterVec), :28
  n.cur = 1 - n.cur;
  size_t ns = (size_t)t.numLoci * t.numRates;
  if (n.sol[n.cur].size() != ns * 4) { n.sol[n.cur].assign(ns * 4, 0.0); n.expo[n.cur].assign(ns, 0.0f); }
  const int S = t.numLoci;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(t.threads) if (t.threads > 1)
#endif
  for (int i = 0; i < S; i++) computeSolution(t, n, P, i);   // :29-32 (sequential in the reference)
  t.updates += (long)ns;
  n.solved = true;               // :33
  n.updateFlag = true;           // :34
}

// AugTree.cpp:114-134.  CanSolve() always returns false for internal nodes
// (IntermediateNode.cpp:15-24 push_backs onto a pre-sized vector<bool>), so children are
// always visited; solved children return at once.
void trySolve(Tree& t, Node* vertex, const double* withinP, const double* betweenP) {
  if (vertex->isTip || vertex->solved) return;
  for (Node* c : vertex->children) trySolve(t, c, withinP, betweenP);
  if (vertex->children.at(0)->within) computeSolutions(t, *vertex, withinP);   // :125-128
  else computeSolutions(t, *vertex, betweenP);                                 // :129-132
}

// AugTree.cpp:289-326
void computeLoglik(Tree& t, const double* withinP, const double* betweenP, const double* limProbs) {
  const int R = t.numRates, S = t.numLoci, N = t.numTips;
  trySolve(t, &t.v[N], withinP, betweenP);                                     // :293
  std::vector<double> likProp((size_t)S * R);
  for (int l = 0; l < S; l++)
    for (int r = 0; r < R; r++) {
      double a[4];
      getSolution(t, t.v[N], l, r, a);
      // arma::dot -> op_dot::direct_dot_arma: two interleaved accumulators
      double v1 = 0.0, v2 = 0.0;
      v1 += a[0] * limProbs[0]; v2 += a[1] * limProbs[1];
      v1 += a[2] * limProbs[2]; v2 += a[3] * limProbs[3];
      likProp[(size_t)l * R + r] = v1 + v2;                                    // :300
    }
  std::vector<double> rateAveraged(S);
  std::vector<float> ev(R, 0.0f);                                              // :306 fvec, outside the loop
  const int nV = (int)t.v.size();
  for (int l = 0; l < S; l++) {
    if (!t.quirkCarry) std::fill(ev.begin(), ev.end(), 0.0f);                  // textbook switch (not the reference)
    for (int r = 0; r < R; r++)
      for (int i = N; i < nV; i++) ev[r] += getExponent(t, t.v[i], l, r);      // :309-315, fp32, vertex-id order
    float mx = ev[0];
    for (int r = 1; r < R; r++) if (ev[r] > mx) mx = ev[r];
    double maxExponent = mx;                                                   // :317
    for (int r = 0; r < R; r++) ev[r] -= (float)maxExponent;                   // :318 (fvec -= eT(double))
    // mean(likPropVec.rows(..) % exp(exponentVec)): mixed Schur promotes both sides to double,
    // exp() of an fvec is expf; op_mean::direct_mean = accumulate (two accumulators) / n.
    double a1 = 0.0, a2 = 0.0;
    int i = 0, j = 1;
    for (; j < R; i += 2, j += 2) {
      a1 += likProp[(size_t)l * R + i] * (double)std::exp(ev[i]);
      a2 += likProp[(size_t)l * R + j] * (double)std::exp(ev[j]);
    }
    if (i < R) a1 += likProp[(size_t)l * R + i] * (double)std::exp(ev[i]);
    double mean = (a1 + a2) / (double)R;
    rateAveraged[l] = std::log(mean) + maxExponent;                            // :322
  }
  // sum(vec) -> arrayops::accumulate, two interleaved accumulators                :325
  double a1 = 0.0, a2 = 0.0;
  int i = 0, j = 1;
  for (; j < S; i += 2, j += 2) { a1 += rateAveraged[i]; a2 += rateAveraged[j]; }
  if (i < S) a1 += rateAveraged[i];
  t.logLik = a1 + a2;
}

// AugTree.cpp:48-57
void bindMatrix(Node* vertex, bool withinCluster) {
  vertex->within = withinCluster;
  if (!vertex->isTip) for (Node* c : vertex->children) bindMatrix(c, withinCluster);
}

// AugTree.cpp:30-46.  clusterMRCAs are 1-based; ids <= numTips (singleton clusters) are skipped.
void associateTransProbMatrices(Tree& t, const std::vector<long>& clusterMRCAs) {
  for (Node& n : t.v) n.within = false;
  for (long i : clusterMRCAs)
    if (i > t.numTips)
      for (Node* c : t.v[i - 1].children) bindMatrix(c, true);
}

// AugTree.cpp:59-86 / 88-101.  edge: column-major E x 2, 1-based (R convention); parents = column 0.
bool linkEdges(Tree& t, const int* edge, int nEdgeRows) {
  const int nV = (int)t.v.size();
  for (int k = 0; k < nEdgeRows; k++) {
    int p = edge[k] - 1, c = edge[k + nEdgeRows] - 1;
    if (p < t.numTips || p >= nV || c < 0 || c >= nV) { t.err = "edge matrix entry out of range"; return false; }
    t.v[p].children.push_back(&t.v[c]);
    t.v[c].parent = &t.v[p];
  }
  return true;
}

// IntermediateNode.cpp:5-13
void invalidateSolution(Node* n) {
  n->solved = false;
  if (n->parent != nullptr) invalidateSolution(n->parent);
}

// AugTree.cpp:273-287
void checkAndInvalidateBetweenRecursive(Node* cur) {
  if (!cur->isTip) {
    if (!cur->children.at(0)->within) {
      cur->solved = false;
      for (Node* c : cur->children) checkAndInvalidateBetweenRecursive(c);
    }
  }
}

// AugTree.cpp:152-175
void nniVerticesWithin(Node* cur, std::vector<unsigned>& out) {
  bool movePossible = false;
  if (!cur->isTip) {
    for (Node* c : cur->children) { movePossible = movePossible || !c->isTip; if (movePossible) break; }
    if (movePossible) {
      out.push_back((unsigned)cur->id);
      for (Node* c : cur->children) nniVerticesWithin(c, out);
    }
  }
}

inline bool inMRCAs(const std::vector<long>& m, long id1based) {
  return std::find(m.begin(), m.end(), id1based) != m.end();
}

// AugTree.cpp:185-209
void nniVerticesBetween(Node* cur, std::vector<unsigned>& out, const std::vector<long>& mrcas) {
  bool movePossible = false;
  if (!cur->isTip && !inMRCAs(mrcas, cur->id + 1)) {
    for (Node* c : cur->children) {
      movePossible = movePossible || (!c->isTip && !inMRCAs(mrcas, c->id + 1));
      if (movePossible) break;
    }
    if (movePossible) {
      out.push_back((unsigned)cur->id);
      for (Node* c : cur->children) nniVerticesBetween(c, out, mrcas);
    }
  }
}

// AugTree.cpp:253-271
std::vector<unsigned> twoVerticesForNNI(Tree& t, Node* subtreeRoot, const std::vector<long>& mrcas) {
  std::vector<unsigned> g;
  for (Node* c : subtreeRoot->children) {
    if (c->isTip || inMRCAs(mrcas, c->id + 1)) g.push_back((unsigned)c->id);
    else {
      unsigned long sel = t.rng.uniform_int(c->children.size());
      g.push_back((unsigned)c->children.at(sel)->id);
    }
  }
  return g;
}

void removeChild(Node* p, Node* c) {   // IntermediateNode.cpp:73-84
  auto it = std::find(p->children.begin(), p->children.end(), c);
  if (it != p->children.end()) p->children.erase(it);
}

// AugTree.cpp:211-226
void rearrangeTreeNNI(Tree& t, unsigned id1, unsigned id2) {
  Node* a = &t.v.at(id1);
  Node* b = &t.v.at(id2);
  removeChild(a->parent, a); a->parent->children.push_back(b);
  removeChild(b->parent, b); b->parent->children.push_back(a);
  Node* p1 = a->parent; Node* p2 = b->parent;
  a->parent = p2; b->parent = p1;
  invalidateSolution(a->parent);
  invalidateSolution(b->parent);
}

// AugTree.cpp:228-251
void addEdgeRecursion(const Tree& t, int* out, int nRows, int& line, const Node* cur) {
  if (cur->parent != nullptr) {
    out[line] = cur->parent->id + 1;
    out[line + nRows] = cur->id + 1;
    line++;
  }
  if (!cur->isTip) for (const Node* c : cur->children) addEdgeRecursion(t, out, nRows, line, c);
}

void negateAllUpdateFlags(Tree& t) { for (Node& n : t.v) n.updateFlag = false; }   // AugTree.cpp:368-374

std::vector<long> toLongs(const double* x, int n) {
  std::vector<long> v(n);
  for (int i = 0; i < n; i++) v[i] = (long)x[i];
  return v;
}

}  // namespace

// ======================================================================== C API
extern "C" {

// gsl_rng_taus known-answer hook for tests: first n outputs after seeding with `seed`.
void orc_taus_stream(unsigned long seed, int n, unsigned long* out) {
  TausRng r; r.seed(seed);
  for (int i = 0; i < n; i++) out[i] = r.get();
}

const char* orc_error(void* h) { return ((Tree*)h)->err.c_str(); }
// Settings applied to handles created afterwards.  threads > 1 runs the site loop of
// IntermediateNode.cpp:29-32 under OpenMP (what `numLikThreads` was documented to do; the shipped
// reference is single-threaded).  quirk carry off = textbook per-site reduction.
static int g_threads = 1;
static bool g_quirkCarry = true;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
void orc_set_quirk_carry(int on) { g_quirkCarry = on != 0; }
long orc_updates(void* h) { return ((Tree*)h)->updates; }
void orc_reset_updates(void* h) { ((Tree*)h)->updates = 0; }

// clusterFunArmadillo.cpp:64-113 (defineMap + getConvertedAlignment) on a flat char matrix.
// chars: [numTips][numLoci] row-major, one byte per cell; states: the 4 names(limProbs) in order.
// Output: 4-bit mask per cell (bit j = state j allowed).  Unknown symbol -> mask 0 (the reference
// yields an empty uvec that later fails inside Armadillo); returns the number of such cells.
long orc_convert_alignment(const char* chars, long n, const char* states, uint8_t* out) {
  uint8_t map[256];
  std::memset(map, 0, sizeof map);
  for (int j = 0; j < 4; j++) map[(unsigned char)states[j]] = (uint8_t)(1u << j);
  // ambiguity codes are defined on the literal labels "a","t","c","g" (clusterFunArmadillo.cpp:78-90)
  uint8_t a = map[(unsigned char)'a'], tt = map[(unsigned char)'t'], c = map[(unsigned char)'c'], g = map[(unsigned char)'g'];
  map[(unsigned char)'r'] = a | g;  map[(unsigned char)'y'] = c | tt; map[(unsigned char)'s'] = c | g;
  map[(unsigned char)'w'] = a | tt; map[(unsigned char)'k'] = g | tt; map[(unsigned char)'m'] = a | c;
  map[(unsigned char)'b'] = c | g | tt; map[(unsigned char)'d'] = a | g | tt;
  map[(unsigned char)'h'] = a | c | tt; map[(unsigned char)'v'] = a | c | g;
  map[(unsigned char)'-'] = map[(unsigned char)'n'] = map[(unsigned char)'.'] = a | tt | c | g;
  long bad = 0;
  for (long i = 0; i < n; i++) { out[i] = map[(unsigned char)chars[i]]; bad += out[i] == 0; }
  return bad;
}

// logLikCpp, clusterFunArmadillo.cpp:36-62.  Returns a handle (the XPtr<AugTree>) or NULL.
void* orc_loglik_create(const int* edge, int nEdgeRows, const double* clusterMRCAs, int nMRCAs,
                        const double* limProbs, const double* within, const double* between,
                        int numRates, const uint8_t* tipmask, int numTips, int numLoci,
                        int withinIdx, int betweenIdx, double* logLik) {
  Tree* t = new Tree;
  t->threads = g_threads; t->quirkCarry = g_quirkCarry;
  t->numTips = numTips; t->numLoci = numLoci; t->numRates = numRates;
  t->withinIdx = withinIdx; t->betweenIdx = betweenIdx;
  t->alignment.assign(tipmask, tipmask + (size_t)numTips * numLoci);
  t->rng.seed(0);                               // gsl_rng_alloc(gsl_rng_taus): default seed, :51
  t->v.resize(nEdgeRows + 1);                   // AugTree.cpp:61
  for (int i = 0; i < (int)t->v.size(); i++) {
    t->v[i].id = i;
    t->v[i].isTip = i < numTips;
    t->v[i].solved = i < numTips;
    if (i < numTips) t->v[i].input = &t->alignment[(size_t)i * numLoci];
  }
  if (!linkEdges(*t, edge, nEdgeRows)) { delete t; return nullptr; }
  for (int i = numTips; i < (int)t->v.size(); i++)
    if (t->v[i].children.size() != 2) { delete t; return nullptr; }   // bifurcating only (IntermediateNode.cpp:59)
  associateTransProbMatrices(*t, toLongs(clusterMRCAs, nMRCAs));
  computeLoglik(*t, within, between, limProbs);
  *logLik = t->logLik;
  return t;
}

// clusterFunArmadillo.cpp:117-138
double orc_new_between(void* h, const double* within, const double* newBetween, int newIdx, const double* limProbs) {
  Tree& t = *(Tree*)h;
  negateAllUpdateFlags(t);
  t.betweenIdx = newIdx;
  checkAndInvalidateBetweenRecursive(&t.v[t.numTips]);
  computeLoglik(t, within, newBetween, limProbs);
  return t.logLik;
}

// clusterFunArmadillo.cpp:142-164
double orc_new_within(void* h, const double* newWithin, const double* between, int newIdx, const double* limProbs) {
  Tree& t = *(Tree*)h;
  negateAllUpdateFlags(t);
  t.withinIdx = newIdx;
  for (Node& n : t.v) if (!n.isTip) n.solved = false;   // InvalidateAll, AugTree.cpp:136-142
  computeLoglik(t, newWithin, between, limProbs);
  return t.logLik;
}

// clusterFunArmadillo.cpp:168-203.  edgeOut: (2N-2) x 2 column-major, 1-based, pre-order.
// Returns NaN (instead of the GSL abort) when there is no NNI candidate.
double orc_within_nni(void* h, const double* within, const double* between, int mrca, int numMoves,
                      const double* limProbs, int* edgeOut) {
  Tree& t = *(Tree*)h;
  negateAllUpdateFlags(t);
  std::vector<unsigned> cand;
  nniVerticesWithin(&t.v.at(mrca - 1), cand);
  if (cand.empty()) return NAN;
  std::vector<long> placeholder(1, -1);        // uvec(1) = -1 wraps to 2^64-1; never matches an id (:188-189)
  for (int m = 0; m < numMoves; m++) {
    unsigned long ri = t.rng.uniform_int(cand.size());
    std::vector<unsigned> g = twoVerticesForNNI(t, &t.v.at(cand.at(ri)), placeholder);
    rearrangeTreeNNI(t, g.at(0), g.at(1));
  }
  computeLoglik(t, within, between, limProbs);
  int line = 0, nRows = (int)t.v.size() - 1;
  addEdgeRecursion(t, edgeOut, nRows, line, &t.v[t.numTips]);
  return t.logLik;
}

// clusterFunArmadillo.cpp:207-242
double orc_between_nni(void* h, const double* within, const double* between, const double* clusterMRCAs,
                       int nMRCAs, int numMoves, const double* limProbs, int* edgeOut) {
  Tree& t = *(Tree*)h;
  std::vector<long> mrcas = toLongs(clusterMRCAs, nMRCAs);
  negateAllUpdateFlags(t);
  std::vector<unsigned> cand;
  nniVerticesBetween(&t.v[t.numTips], cand, mrcas);
  if (cand.empty()) return NAN;
  for (int m = 0; m < numMoves; m++) {
    unsigned long ri = t.rng.uniform_int(cand.size());
    std::vector<unsigned> g = twoVerticesForNNI(t, &t.v.at(cand.at(ri)), mrcas);
    rearrangeTreeNNI(t, g.at(0), g.at(1));
  }
  computeLoglik(t, within, between, limProbs);
  int line = 0, nRows = (int)t.v.size() - 1;
  addEdgeRecursion(t, edgeOut, nRows, line, &t.v[t.numTips]);
  return t.logLik;
}

// clusterFunArmadillo.cpp:246-276; HandleSplit/HandleMerge AugTree.cpp:328-346
double orc_split_merge(void* h, const double* within, const double* between, const int* mrcas, int n,
                       const double* limProbs) {
  Tree& t = *(Tree*)h;
  negateAllUpdateFlags(t);
  if (n == 1) {
    Node* m = &t.v.at(mrcas[0] - 1);
    invalidateSolution(m);
    for (Node* c : m->children) c->within = false;
  } else {
    invalidateSolution(t.v.at(mrcas[0] - 1).parent);
    for (int i = 0; i < n; i++) t.v.at(mrcas[i] - 1).within = true;
  }
  computeLoglik(t, within, between, limProbs);
  return t.logLik;
}

// clusterFunArmadillo.cpp:280-284; AugTree.cpp:348-366
void orc_restore(void* h, const int* edge, int nEdgeRows, int withinIdx, int betweenIdx, int nniMove,
                 const int* clusterMRCAs, int nMRCAs, int splitMergeMove) {
  Tree& t = *(Tree*)h;
  t.withinIdx = withinIdx; t.betweenIdx = betweenIdx;
  if (nniMove) {                                // BuildTreeNoAssign, AugTree.cpp:88-101
    for (Node& n : t.v) n.children.clear();
    linkEdges(t, edge, nEdgeRows);
  }
  for (Node& n : t.v) {                         // RestoreIterVecAndExp, IntermediateNode.h:26-33
    if (!n.isTip) { if (n.updateFlag) n.cur = n.prev; n.updateFlag = false; }
  }
  if (splitMergeMove) {
    std::vector<long> m(nMRCAs);
    for (int i = 0; i < nMRCAs; i++) m[i] = clusterMRCAs[i];
    associateTransProbMatrices(t, m);
  }
}

void orc_final_deallocate(void* h) { delete (Tree*)h; }   // clusterFunArmadillo.cpp:288-294

// clusterFunArmadillo.cpp:298-302 with the reference's sizeof values on LP64/libstdc++:
// sizeof(S) = 32, sizeof(mapContentType) = sizeof(std::vector<...>) = 24.
double orc_max_map_size_estimate(double allowedSizeInMegs, unsigned numRateCats) {
  return std::floor((allowedSizeInMegs * 1e6) / (32 + 24 * (double)numRateCats + 16));
}

// ---- introspection used by parity tests
double orc_loglik(void* h) { return ((Tree*)h)->logLik; }
void orc_get_partial(void* h, int node, double* sol /*[S][R][4]*/, float* expo /*[S][R]*/) {
  Tree& t = *(Tree*)h;
  const Node& n = t.v.at(node);
  size_t ns = (size_t)t.numLoci * t.numRates;
  std::memcpy(sol, n.sol[n.cur].data(), ns * 4 * sizeof(double));
  std::memcpy(expo, n.expo[n.cur].data(), ns * sizeof(float));
}
void orc_get_topology(void* h, int* parent, int* child0, int* child1, int* within) {
  Tree& t = *(Tree*)h;
  for (size_t i = 0; i < t.v.size(); i++) {
    const Node& n = t.v[i];
    parent[i] = n.parent ? n.parent->id : -1;
    child0[i] = n.isTip ? -1 : n.children[0]->id;
    child1[i] = n.isTip ? -1 : n.children[1]->id;
    within[i] = n.within ? 1 : 0;
  }
}

}  // extern "C"
