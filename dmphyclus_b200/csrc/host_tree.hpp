// host_tree.hpp — host-side topology and proposal state machine of the likelihood engine.
//
// The reference keeps a pointer graph of TreeNode objects (src/TreeNode.h:80-125,
// src/IntermediateNode.h:6-53) and walks it recursively.  Here the same state is a handful of flat
// arrays indexed by vertex id (tips 0..N-1, internal N..2N-2, root = N — AugTree.cpp:59-86), and what
// the reference expresses as "unsolved nodes reached from the root" (AugTree::TrySolve,
// AugTree.cpp:114-134) becomes a LEVEL SCHEDULE: the dirty nodes bucketed by height above the clean
// frontier, one batched kernel launch per bucket.  No arithmetic happens here.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "task.hpp"

namespace dmpc {

// GSL's gsl_rng_taus (rng/taus.c) with gsl_rng_alloc's default seed and gsl_rng_uniform_int's rejection
// loop (rng/rng.c): the reference draws its NNI picks from this stream and never seeds it
// (clusterFunArmadillo.cpp:51,187,227; AugTree.cpp:266), so reproducing it is part of trajectory parity.
struct TausRng {
  uint32_t s1 = 0, s2 = 0, s3 = 0;
  static uint32_t step(uint32_t s, int a, int b, uint32_t c, int d) { return ((s & c) << d) ^ (((s << a) ^ s) >> b); }
  uint32_t next() {
    s1 = step(s1, 13, 19, 4294967294u, 12);
    s2 = step(s2, 2, 25, 4294967288u, 4);
    s3 = step(s3, 3, 11, 4294967280u, 17);
    return s1 ^ s2 ^ s3;
  }
  void seed(unsigned long s) {
    if (s == 0) s = 1;
    s1 = 69069u * (uint32_t)s;
    s2 = 69069u * s1;
    s3 = 69069u * s2;
    for (int i = 0; i < 6; i++) next();
  }
  unsigned long uniform_int(unsigned long n) {   // caller guarantees n > 0
    const unsigned long scale = 0xffffffffUL / n;
    unsigned long k;
    do { k = next() / scale; } while (k >= n);
    return k;
  }
};

// A std::vector whose storage outlives it.  A logLikFromClusInd-style create / evaluate / destroy cycle allocates ~12 MB
// of host vectors at 100,000 tips; blocks of that size go back to the operating system when they are freed (malloc trims
// or unmaps them), so EVERY create paid the page faults of first touch again — about a third of its host time
// (tools/host_bench.cpp).  The storage of a dying vector is parked instead (emptied, capacity kept), and a vector about
// to grow beyond its capacity (assign / resize / reserve) first looks for a parked block that is large enough — the
// smallest such one.  Contents never survive: a parked vector is empty, and every user assigns or clears before it reads.
// Only the calls made on the PooledVec itself look into the pool; through a std::vector& it behaves like any vector.
template <class T>
struct PooledVec : std::vector<T> {
  typedef std::vector<T> Base;
  PooledVec() {}
  PooledVec(const PooledVec& o) : Base() { want(o.size()); Base::operator=(o); }
  PooledVec& operator=(const PooledVec& o) { want(o.size()); Base::operator=(o); return *this; }
  PooledVec& operator=(const Base& o) { want(o.size()); Base::operator=(o); return *this; }
  ~PooledVec() { park(); }
  void assign(size_t n, const T& v) { want(n); Base::assign(n, v); }
  template <class It, class = typename std::enable_if<!std::is_integral<It>::value>::type>
  void assign(It first, It last) { want((size_t)std::distance(first, last)); Base::assign(first, last); }
  void resize(size_t n) { want(n); Base::resize(n); }
  void reserve(size_t n) { want(n); Base::reserve(n); }

 private:
  struct Pool { std::mutex m; std::vector<Base> free; };
  static Pool& pool() { static Pool* p = new Pool; return *p; }       // never destroyed: handles may die after the statics
  // bytes below which parking is pointless / above which a block is not worth holding on to; blocks kept per element type
  static constexpr size_t KEEP_MIN = 64 * 1024, KEEP_MAX = (size_t)64 << 20, KEEP_MAX_VECS = 48;
  // about to hold n elements: if that needs new storage, take the smallest parked block that fits (what this vector
  // held so far moves along; its old block is parked in turn)
  void want(size_t n) {
    if (n <= this->capacity() || n * sizeof(T) < KEEP_MIN) return;
    Base got;
    {
      Pool& p = pool();
      std::lock_guard<std::mutex> lk(p.m);
      size_t best = p.free.size();
      for (size_t i = 0; i < p.free.size(); i++)
        if (p.free[i].capacity() >= n && (best == p.free.size() || p.free[i].capacity() < p.free[best].capacity())) best = i;
      if (best == p.free.size()) return;
      got.swap(p.free[best]);
      p.free[best].swap(p.free.back());
      p.free.pop_back();
    }
    got.assign(this->begin(), this->end());      // (callers grow before they fill: this copies nothing in practice)
    Base::swap(got);
    PooledVec old;
    old.Base::swap(got);                         // the block held so far: parked by old's destructor if it is worth it
  }
  void park() {
    if (this->capacity() * sizeof(T) < KEEP_MIN || this->capacity() * sizeof(T) > KEEP_MAX) return;
    this->clear();
    Pool& p = pool();
    std::lock_guard<std::mutex> lk(p.m);
    if (p.free.size() >= KEEP_MAX_VECS) return;
    p.free.emplace_back();
    p.free.back().swap(*this);
  }
};

struct HostTree {
  int numTips = 0, numVerts = 0;
  PooledVec<int32_t> parent;             // -1 for the root
  PooledVec<int32_t> kids;               // [2*v + i]: ordered children (-1 = absent); order = the reference's _children
  PooledVec<uint8_t> nkids;
  PooledVec<uint8_t> within;             // _withinParentBranch
  PooledVec<uint8_t> solved;             // _isSolved (tips: always 1)
  PooledVec<uint8_t> touched;            // _updateFlag
  PooledVec<uint8_t> cur;                // which of the two HBM slots holds the vertex's current partial
  int withinIdx = 0, betweenIdx = 0;
  TausRng rng;
  std::string err;
  // small-clade fusion: a vertex with at most fuseTips tips below it (and not the root) is never stored; it is
  // recomputed from the tip masks inside the task of its nearest stored ancestor.  1 = every vertex is stored.
  int fuseTips = 1;
  PooledVec<int32_t> tips;               // number of tips below each vertex (1 for a tip)
  bool tipsValid = false;
  PooledVec<uint8_t> linkScratch;
  PooledVec<uint8_t> mat;                // fused-size vertices the plan being built evaluates as stored tasks (see CHAIN_MIN)
  bool walkChains = true;                // false: never hand a chain to k_walk (its 32-bit element offsets would overflow)
  bool allInvalid = true;                // every internal vertex is unsolved (InvalidateAll, or nothing computed yet)
  // --- O(touched) NNI bookkeeping.  R hands the committed edge matrix back on every rejected NNI move
  // (RestorePreviousConfig -> BuildTreeNoAssign, AugTree.cpp:88-101, 348-366) and gets the proposed one from every NNI
  // call (BuildEdgeMatrix, :228-251): both are O(N) in the reference.  Here the swaps of the pending proposal are
  // logged; when the matrix R hands back IS the committed one (memcmp), the rollback undoes the log instead of
  // relinking 2N-2 edges, and the proposed matrix is the committed one with only the rows of the rearranged subtree
  // rewritten (a subtree's rows are contiguous in pre-order and a swap inside it keeps their number).
  struct NniUndo { int a, b, pa, pb; int32_t ka[2], kb[2]; bool tipsMoved; };
  std::vector<NniUndo> nniLog;           // swaps since the committed topology (the pending NNI proposal)
  bool pendingNni = false;               // an NNI proposal has been evaluated and neither restored nor followed by another proposal
  int pendingRoot = -1;                  // ... the vertex whose subtree it rearranged (-1: several moves, whole matrix re-exported)
  PooledVec<int32_t> committedEdge;      // R's edge matrix of the committed topology (column-major, 1-based); empty = unknown
  PooledVec<int32_t> rowOf;              // exportValid: row of each vertex's own edge in committedEdge
  bool exportChecked = false, exportValid = false;   // committedEdge is (known to be) exportEdges() of the committed topology
  bool linkOrdered = false;              // the rows link() last saw came in pre-order (a vertex's own row before its children's)
  bool edgePreorder = false;             // ... and so do committedEdge's

  bool isTip(int v) const { return v < numTips; }
  int internalIndex(int v) const { return v - numTips; }
  int root() const { return numTips; }

  void addChild(int p, int c) {          // IntermediateNode::AddChild: push_back
    if (nkids[p] < 2) kids[2 * p + nkids[p]] = c;
    nkids[p]++;
  }
  void removeChild(int p, int c) {       // IntermediateNode::RemoveChild (IntermediateNode.cpp:73-84): erase first match
    for (int i = 0; i < std::min<int>(nkids[p], 2); i++)
      if (kids[2 * p + i] == c) {
        if (i == 0) kids[2 * p] = kids[2 * p + 1];
        kids[2 * p + 1] = -1;
        nkids[p]--;
        return;
      }
  }

  // AugTree::BuildTree / BuildTreeNoAssign (AugTree.cpp:59-101): edge is column-major, 1-based.
  bool link(const int32_t* edge, int nEdgeRows) {
    tipsValid = false;
    std::fill(kids.begin(), kids.end(), -1);
    std::fill(nkids.begin(), nkids.end(), 0);
    std::fill(parent.begin(), parent.end(), -1);
    for (int k = 0; k < nEdgeRows; k++) {
      int p = edge[k] - 1, c = edge[k + nEdgeRows] - 1;
      if (p < numTips || p >= numVerts || c < 0 || c >= numVerts || c == p) { err = "edge matrix entry out of range"; return false; }
      addChild(p, c);
      parent[c] = p;
    }
    for (int v = numTips; v < numVerts; v++)
      if (nkids[v] != 2) { err = "the tree is not strictly bifurcating (IntermediateNode.cpp:59)"; return false; }
    // Tips below every vertex in one reverse pass over the rows, valid when they come in pre-order (a vertex's own row
    // before the rows of its children) — the order BuildEdgeMatrix emits (AugTree.cpp:228-251) and R hands back on every
    // NNI rollback.  Any other order falls back to the traversal in computeTips().
    tips.assign(numVerts, 0);
    for (int v = 0; v < numTips; v++) tips[v] = 1;
    PooledVec<uint8_t>& pend = linkScratch;          // child rows of each vertex not folded in yet
    pend.assign(numVerts, 0);
    for (int v = numTips; v < numVerts; v++) pend[v] = 2;
    bool ordered = true;
    for (int k = nEdgeRows - 1; k >= 0 && ordered; k--) {
      const int p = edge[k] - 1, c = edge[k + nEdgeRows] - 1;
      if (pend[c] != 0) ordered = false;
      tips[p] += tips[c];
      pend[p]--;
    }
    tipsValid = linkOrdered = ordered;
    return true;
  }

  bool build(const int32_t* edge, int nEdgeRows, int nTips) {
    numTips = nTips;
    numVerts = nEdgeRows + 1;
    if (numVerts != 2 * numTips - 1) { err = "edge matrix must have 2*numTips-2 rows"; return false; }
    parent.assign(numVerts, -1);
    kids.assign(2 * (size_t)numVerts, -1);
    nkids.assign(numVerts, 0);
    within.assign(numVerts, 0);
    solved.assign(numVerts, 0);
    touched.assign(numVerts, 0);
    cur.assign(numVerts, 1);             // first evaluation writes slot 0
    mat.assign(numVerts, 0);
    for (int v = 0; v < numTips; v++) solved[v] = 1;
    rng.seed(0);
    if (!link(edge, nEdgeRows)) return false;
    committedEdge.assign(edge, edge + 2 * (size_t)nEdgeRows);
    edgePreorder = linkOrdered;
    exportChecked = exportValid = false; pendingNni = false; nniLog.clear();
    return checkRooted();
  }

  // after link(): one tree, rooted at vertex numTips+1, holding every vertex
  bool checkRooted() {
    if (parent[root()] != -1) { err = "vertex numTips+1 must be the root"; return false; }
    // every vertex but the root is named as a child — with 2N-2 rows, exactly once (a child named twice would leave
    // another vertex hanging nowhere, and the traversal below would count the twice-named one twice)
    int orphans = 0;
    for (int v = 0; v < numVerts; v++) orphans += parent[v] == -1;
    if (orphans != 1) { err = "edge matrix is not a single rooted tree"; return false; }
    // ... and hangs below the root.  Rows in pre-order (tipsValid: every child's own rows come after the row that names
    // it, so the rows hold no cycle): following the parents from any vertex ends at the root — no traversal needed
    // (100,000 tips: 0.4 ms)
    if (tipsValid) return true;
    std::vector<int> stack(1, root());
    int seen = 0;
    while (!stack.empty()) {
      int v = stack.back(); stack.pop_back(); seen++;
      if (!isTip(v)) { stack.push_back(kids[2 * v]); stack.push_back(kids[2 * v + 1]); }
    }
    if (seen != numVerts) { err = "edge matrix is not a single rooted tree"; return false; }
    return true;
  }

  // AugTree::AssociateTransProbMatrices + BindMatrix (AugTree.cpp:30-57): every flag false, then for each cluster MRCA
  // that is an internal vertex, all its descendants' parent branches are "within".
  template <class T>
  void associate(const T* mrcas, int n) {
    std::fill(within.begin(), within.end(), 0);
    const int nRows = numVerts - 1;
    // vertices the per-cluster traversal would visit: when the clusters cover little of the tree (a caterpillar cut
    // into one clade and singletons) it is the cheaper of the two — it pays ~5x more per vertex than the pass per row
    long covered = tipsValid ? 0 : nRows;
    if (tipsValid)
      for (int i = 0; i < n; i++) { const long m = (long)mrcas[i]; if (m > numTips && m <= numVerts) covered += 2L * tips[m - 1] - 2; }
    if (edgePreorder && !pendingNni && nniLog.empty() && committedEdge.size() == 2 * (size_t)nRows && 5 * covered >= nRows) {
      // The committed edge matrix IS the topology (no swap since it was linked, exported or restored), and lists every vertex's own row before the rows of its children
      // (what BuildEdgeMatrix emits and R hands back): one pass down the rows — a child is within when its parent is a
      // cluster MRCA or is within itself — instead of a traversal per cluster (100,000 tips: 1.5 -> 0.3 ms).  mat is
      // all zero between plans and serves as the MRCA mark.
      for (int i = 0; i < n; i++) { const long m = (long)mrcas[i]; if (m > numTips && m <= numVerts) mat[m - 1] = 1; }
      const int32_t* pe = committedEdge.data();
      const int32_t* ce = pe + nRows;
      for (int k = 0; k < nRows; k++) { const int p = pe[k] - 1; within[ce[k] - 1] = within[p] | mat[p]; }
      for (int i = 0; i < n; i++) { const long m = (long)mrcas[i]; if (m > numTips && m <= numVerts) mat[m - 1] = 0; }
      return;
    }
    associateByTraversal(mrcas, n);
  }
  template <class T>
  void associateByTraversal(const T* mrcas, int n) {
    std::vector<int> stack;
    for (int i = 0; i < n; i++) {
      long m = (long)mrcas[i];
      if (m > numTips && m <= numVerts) {
        stack.push_back(kids[2 * (m - 1)]); stack.push_back(kids[2 * (m - 1) + 1]);
        while (!stack.empty()) {
          int v = stack.back(); stack.pop_back();
          within[v] = 1;
          if (!isTip(v)) { stack.push_back(kids[2 * v]); stack.push_back(kids[2 * v + 1]); }
        }
      }
    }
  }

  // AugTree.cpp:368-374.  Every proposal starts here, so an NNI proposal that was not restored has been ACCEPTED by now.
  void negateAllUpdateFlags() {
    std::fill(touched.begin(), touched.end(), 0);
    if (pendingNni) commitPendingNni();
  }
  void commitPendingNni() {
    const int nRows = numVerts - 1;
    if (exportValid && pendingRoot >= 0) {
      exportSubtree(pendingRoot, committedEdge.data(), true);      // the same rows the proposal handed to R
    } else {
      committedEdge.resize(2 * (size_t)nRows);
      exportEdges(committedEdge.data());
      indexRows();
      exportChecked = exportValid = true;
    }
    edgePreorder = true;                   // either way committedEdge is now this tree's own pre-order export
    pendingNni = false; pendingRoot = -1; nniLog.clear();
  }
  void indexRows() {
    const int nRows = numVerts - 1;
    rowOf.assign(numVerts, -1);
    for (int k = 0; k < nRows; k++) rowOf[committedEdge[k + nRows] - 1] = k;
  }
  // before the swaps of an NNI proposal: make sure we know whether the committed matrix is our own pre-order export
  void beginNni() {
    nniLog.clear();
    pendingRoot = -1;
    if (exportChecked) return;
    const int nRows = numVerts - 1;
    std::vector<int32_t> tmp(2 * (size_t)nRows);
    exportEdges(tmp.data());
    exportValid = committedEdge.size() == tmp.size() && std::equal(tmp.begin(), tmp.end(), committedEdge.begin());
    if (exportValid) indexRows();
    exportChecked = true;
  }
  // the rows of sr's subtree, in pre-order, into `out` at the place they have in committedEdge
  void exportSubtree(int sr, int32_t* out, bool reindex) {
    const int nRows = numVerts - 1;
    int line = parent[sr] == -1 ? 0 : rowOf[sr] + 1;
    std::vector<int> stack;
    stack.push_back(kids[2 * sr + 1]); stack.push_back(kids[2 * sr]);
    while (!stack.empty()) {
      int v = stack.back(); stack.pop_back();
      out[line] = parent[v] + 1; out[line + nRows] = v + 1;
      if (reindex) rowOf[v] = line;
      line++;
      if (!isTip(v)) { stack.push_back(kids[2 * v + 1]); stack.push_back(kids[2 * v]); }
    }
  }
  // the edge matrix of the PROPOSED topology (after beginNni + the swaps): numMoves == 1 rearranged only subtreeRoot's subtree
  void exportProposed(int32_t* out, int subtreeRoot) {
    const int nRows = numVerts - 1;
    pendingNni = true;
    if (exportValid && subtreeRoot >= 0) {
      std::copy(committedEdge.begin(), committedEdge.end(), out);
      exportSubtree(subtreeRoot, out, false);
      pendingRoot = subtreeRoot;
    } else {
      exportEdges(out);
      pendingRoot = -1;
    }
    (void)nRows;
  }
  // RestorePreviousConfig's BuildTreeNoAssign (AugTree.cpp:88-101, 348-366) for a rejected NNI proposal
  bool restoreTopology(const int32_t* edge, int nEdgeRows) {
    const size_t n2 = 2 * (size_t)nEdgeRows;
    if (pendingNni && !nniLog.empty() && committedEdge.size() == n2 && std::equal(edge, edge + n2, committedEdge.begin())) {
      for (size_t i = nniLog.size(); i-- > 0;) {                 // the committed matrix: undo the swaps, newest first
        const NniUndo& u = nniLog[i];
        kids[2 * u.pa] = u.ka[0]; kids[2 * u.pa + 1] = u.ka[1];
        kids[2 * u.pb] = u.kb[0]; kids[2 * u.pb + 1] = u.kb[1];
        parent[u.a] = u.pa; parent[u.b] = u.pb;
        if (u.tipsMoved) {
          const int delta = tips[u.b] - tips[u.a];
          for (int v = u.pa; v != -1; v = parent[v]) tips[v] -= delta;   // (parents are back in place: the chains rearrangeNNI walked)
          for (int v = u.pb; v != -1; v = parent[v]) tips[v] += delta;
        }
      }
      pendingNni = false; pendingRoot = -1; nniLog.clear();
      return true;
    }
    pendingNni = false; pendingRoot = -1; nniLog.clear();
    if (!link(edge, nEdgeRows) || !checkRooted()) return false;
    committedEdge.assign(edge, edge + n2);
    edgePreorder = linkOrdered;
    exportChecked = exportValid = false;
    return true;
  }

  // RestoreIterVecAndExp (IntermediateNode.h:26-33) for every vertex the pending proposal touched: the vertex goes back
  // to its other slot and its flag is cleared; f(v, slot) is told where it now lives.  A root-path proposal touches a
  // few dozen of the internal vertices, so the flags are skipped eight at a time where none is set.
  template <class F>
  void rollbackTouched(F f) {
    int v = numTips;
    const uint8_t* tp = touched.data();
    while (v < numVerts) {
      if (v + 8 <= numVerts) {
        uint64_t w;
        std::memcpy(&w, tp + v, 8);
        if (!w) { v += 8; continue; }
      }
      const int end = std::min(numVerts, v + 8);
      for (; v < end; v++)
        if (touched[v]) { cur[v] ^= 1; touched[v] = 0; f(v, (int)cur[v]); }
    }
  }

  void invalidateSolution(int v) { for (; v != -1; v = parent[v]) solved[v] = 0; }        // IntermediateNode.cpp:5-13
  void invalidateAll() { for (int v = numTips; v < numVerts; v++) solved[v] = 0; allInvalid = true; }   // AugTree.cpp:136-142
  void invalidateBetween() {                                                              // AugTree.cpp:273-287
    std::vector<int> stack(1, root());
    while (!stack.empty()) {
      int v = stack.back(); stack.pop_back();
      if (isTip(v) || within[kids[2 * v]]) continue;
      solved[v] = 0;
      stack.push_back(kids[2 * v]); stack.push_back(kids[2 * v + 1]);
    }
  }

  // membership in the cluster-MRCA list of a between-cluster NNI: a byte per vertex, filled once per call (the list has
  // one entry per cluster and is consulted for every vertex above the clusters)
  mutable std::vector<uint8_t> mrcaMark;
  mutable const std::vector<long>* mrcaMarked = nullptr;
  void markMrcas(const std::vector<long>& l) const {
    mrcaMark.assign((size_t)numVerts + 2, 0);
    for (long id1 : l) if (id1 >= 0 && id1 <= numVerts) mrcaMark[(size_t)id1] = 1;
    mrcaMarked = &l;
  }
  void unmarkMrcas() const { mrcaMarked = nullptr; }
  bool inList(const std::vector<long>& l, long id1) const {
    if (mrcaMarked == &l) return id1 >= 0 && id1 <= numVerts && mrcaMark[(size_t)id1];
    return std::find(l.begin(), l.end(), id1) != l.end();
  }

  // GetNNIverticesWithin / Between (AugTree.cpp:144-209): pre-order list of vertices that have at least one
  // internal child (which, for the between variant, is not a cluster MRCA; the walk stops at cluster MRCAs).
  void nniCandidates(int origin, const std::vector<long>* mrcas, std::vector<int>& out) const {
    std::vector<int> stack(1, origin);
    while (!stack.empty()) {
      int v = stack.back(); stack.pop_back();
      if (isTip(v) || (mrcas && inList(*mrcas, v + 1))) continue;
      bool possible = false;
      for (int i = 0; i < 2 && !possible; i++) {
        int c = kids[2 * v + i];
        possible = !isTip(c) && !(mrcas && inList(*mrcas, c + 1));
      }
      if (possible) {
        out.push_back(v);
        stack.push_back(kids[2 * v + 1]);   // pushed second-first so that child 0's subtree is listed first
        stack.push_back(kids[2 * v]);
      }
    }
  }

  // GetTwoVerticesForNNI (AugTree.cpp:253-271)
  void twoVerticesForNNI(int subtreeRoot, const std::vector<long>* mrcas, int out[2]) {
    for (int i = 0; i < 2; i++) {
      int c = kids[2 * subtreeRoot + i];
      if (isTip(c) || (mrcas && inList(*mrcas, c + 1))) out[i] = c;
      else out[i] = kids[2 * c + (int)rng.uniform_int(2)];
    }
  }

  // RearrangeTreeNNI (AugTree.cpp:211-226)
  void rearrangeNNI(int a, int b) {
    NniUndo u;
    u.a = a; u.b = b; u.pa = parent[a]; u.pb = parent[b];
    u.ka[0] = kids[2 * u.pa]; u.ka[1] = kids[2 * u.pa + 1]; u.kb[0] = kids[2 * u.pb]; u.kb[1] = kids[2 * u.pb + 1];
    u.tipsMoved = tipsValid;
    nniLog.push_back(u);
    if (tipsValid) {                       // the two subtrees trade places: only their old ancestors' counts move
      const int delta = tips[b] - tips[a];
      for (int v = parent[a]; v != -1; v = parent[v]) tips[v] += delta;
      for (int v = parent[b]; v != -1; v = parent[v]) tips[v] -= delta;
    }
    removeChild(parent[a], a); addChild(parent[a], b);
    removeChild(parent[b], b); addChild(parent[b], a);
    std::swap(parent[a], parent[b]);
    invalidateSolution(parent[a]);
    invalidateSolution(parent[b]);
  }

  // BuildEdgeMatrix (AugTree.cpp:228-251): pre-order, column-major, 1-based
  void exportEdges(int32_t* out) const {
    const int nRows = numVerts - 1;
    int line = 0;
    std::vector<int> stack(1, root());
    while (!stack.empty()) {
      int v = stack.back(); stack.pop_back();
      if (parent[v] != -1) { out[line] = parent[v] + 1; out[line + nRows] = v + 1; line++; }
      if (!isTip(v)) { stack.push_back(kids[2 * v + 1]); stack.push_back(kids[2 * v]); }
    }
  }

  void computeTips() {
    if (tipsValid) return;
    tips.assign(numVerts, 1);
    std::vector<int> pre;
    pre.reserve(numVerts);
    std::vector<int> stack(1, root());
    while (!stack.empty()) {
      int v = stack.back(); stack.pop_back();
      pre.push_back(v);
      if (!isTip(v)) { stack.push_back(kids[2 * v]); stack.push_back(kids[2 * v + 1]); }
    }
    for (size_t i = pre.size(); i-- > 0;) { int v = pre[i]; if (!isTip(v)) tips[v] = tips[kids[2 * v]] + tips[kids[2 * v + 1]]; }
    tipsValid = true;
  }
  bool fused(int v) const { return v >= numTips && v != numTips && tips[v] <= fuseTips; }

  // The pending evaluation as launchable work: the dirty STORED vertices level by level (children strictly
  // before parents), each with its operands; fused operands come with their token programs.  Mirrors TrySolve:
  // only unsolved vertices reachable from the root through unsolved vertices are recomputed — plus the fused
  // clades hanging off them, which are recomputed whenever their stored ancestor is.  Updates cur/touched/solved
  // the way ComputeSolutions does (IntermediateNode.cpp:28-34).
  struct Plan {
    PooledVec<Task> tasks;
    PooledVec<Token> tokens;
    PooledVec<int32_t> tipList;                  // tips read by the fused programs, task by task
    std::vector<int> levelStart;                 // task range of every LAUNCH UNIT: [materialised clades], the rounds, then one
                                                 // unit per chain level (k_walk takes those together)
    std::vector<int> segStart;                   // task range of every segment (group) of the non-chain units, in task order
    std::vector<int> unitSeg;                    // first segment of each non-chain unit (+ end): unit u = segments
                                                 // [unitSeg[u], unitSeg[u+1]), one block column (gridDim.z) each
    long long storedReads = 0, tipReads = 0, fusedOps = 0;
    long long fp64Ops = 0;                       // DMUL + DADD per (site, rate) of the whole plan: a 4x4 mat-vec is 16 + 12, the
                                                 // product of the two child terms 4; a tip's term is a table look-up
    int chainLevels = 0;                         // trailing single-task levels that k_walk evaluates (0 or >= CHAIN_MIN)
    int materialised = 0;                        // fused-size clades evaluated as stored tasks for that chain (level 0 of the plan)
    bool uses[2] = {false, false};               // slots this plan writes
    PooledVec<int> pre, level, cnt, fill;        // scratch
    PooledVec<int> grp, gRound, gCost, gHead, gNext, gTail;   // scratch of the path grouping
  };

  // mutate = false: introspection replay — tokens address the CURRENT slots and no state changes
  void emitClade(int u, Plan& p, int tipBase, bool mutate = true) {
    const int k0 = kids[2 * u], k1 = kids[2 * u + 1];
    Token t;
    t.a = t.b = 0;
    if (isTip(k0) && isTip(k1)) {
      t.op = OP_CHERRY;
      t.a = (int)p.tipList.size() - tipBase; p.tipList.push_back(k0);
      t.b = (int)p.tipList.size() - tipBase; p.tipList.push_back(k1);
      p.tipReads += 2;
    } else if (isTip(k0) || isTip(k1)) {
      emitClade(isTip(k0) ? k1 : k0, p, tipBase, mutate);
      t.op = OP_REGTIP | (isTip(k0) ? TOK_FIRST : 0);
      t.b = (int)p.tipList.size() - tipBase; p.tipList.push_back(isTip(k0) ? k0 : k1);
      p.tipReads += 1;
    } else {
      const int h = tips[k0] >= tips[k1] ? k0 : k1, o = h == k0 ? k1 : k0;   // heavier side first: stack depth <= log2
      emitClade(h, p, tipBase, mutate);
      Token sp; sp.op = OP_SPILL; sp.a = sp.b = 0; sp.vertex = 0;
      p.tokens.push_back(sp);
      emitClade(o, p, tipBase, mutate);
      t.op = OP_SLOTREG | (h == k0 ? TOK_FIRST : 0);
    }
    const int slot = mutate ? cur[u] ^ 1 : cur[u];
    if (within[k0]) t.op |= TOK_SET;                  // AugTree.cpp:125: the first child's flag picks the matrix set
    if (slot) t.op |= TOK_SLOT;
    t.vertex = u - numTips;
    p.tokens.push_back(t);
    p.fusedOps++;
    p.fp64Ops += (t.op & 7) == OP_CHERRY ? 4 : (t.op & 7) == OP_REGTIP ? 32 : 60;
    p.uses[slot] = true;
    if (mutate) { cur[u] = (uint8_t)slot; touched[u] = 1; solved[u] = 1; }
  }

  // the task of vertex v written to row `outRow`, slot `outSlot`; appends the programs of fused operands to p.tokens
  Task makeTask(int v, int outRow, int outSlot, Plan& p, bool mutate) {
    Task k;
    k.out = outRow;
    k.flags = outSlot ? TF_OUTSLOT : 0;
    k.prog = (int)p.tokens.size(); k.ntok = 0;
    k.tipOff = (int)p.tipList.size(); k.nTips = 0;
    for (int j = 0; j < 2; j++) {
      const int c = kids[2 * v + j];
      int32_t& ref = j ? k.c1 : k.c0;
      if (isTip(c)) {
        ref = c; k.flags |= j ? TF_C1TIP : TF_C0TIP; p.tipReads++;
      } else if (fused(c) && !mat[c]) {
        ref = c - numTips; k.flags |= j ? TF_C1FUSED : TF_C0FUSED;
        const int off = (int)p.tokens.size();
        emitClade(c, p, k.tipOff, mutate);
        k.ntok |= ((int)p.tokens.size() - off) << (16 * j);
      } else {
        ref = c - numTips; if (cur[c]) k.flags |= j ? TF_C1SLOT : TF_C0SLOT; p.storedReads++;
      }
    }
    k.nTips = (int)p.tipList.size() - k.tipOff;
    p.fp64Ops += 4 + 28 * ((k.flags & TF_C0TIP ? 0 : 1) + (k.flags & TF_C1TIP ? 0 : 1));
    if (within[kids[2 * v]]) k.flags |= TF_WITHIN;      // AugTree.cpp:125: the FIRST child's flag picks the matrix set
    return k;
  }

  // Proposal paths (dmpc_eval_proposals): the stored vertices from `s` (or its first stored ancestor) up to the root,
  // bottom-up, appended to p as one run.  Nothing is mutated; the on-path operand of every task but the first is the
  // previous task's output, forwarded in registers.  The parent of a stored vertex is stored, so the run is contiguous.
  void pathRun(int s, Plan& p) {
    computeTips();
    int v = s;
    while (fused(v)) v = parent[v];
    int prev = -1;
    for (; v != -1; prev = v, v = parent[v]) {
      Task k = makeTask(v, v - numTips, cur[v], p, false);   // nothing is written: `out` only names the vertex
      if (prev >= 0) {
        if (kids[2 * v] == prev) k.flags = (k.flags & ~TF_C0SLOT) | TF_C0FWD;
        else k.flags = (k.flags & ~TF_C1SLOT) | TF_C1FWD;
      }
      p.tasks.push_back(k);
    }
  }

  // One proposal of a batch (dmpc_eval_proposals) as a run: the cluster flags a split (HandleSplit, AugTree.cpp:328-336)
  // or a merge (HandleMerge, :338-346) would set are applied to a scratch copy of the two flags involved, the path is
  // planned, the flags are put back.  Ids are 1-based as in dmpc_proposal; the caller has validated them.
  void proposalRun(bool merge, int a1, int b1, Plan& p) {
    if (!merge) {
      const int m = a1 - 1, k0 = kids[2 * m], k1 = kids[2 * m + 1];
      const uint8_t w0 = within[k0], w1 = within[k1];
      within[k0] = 0; within[k1] = 0;
      pathRun(m, p);
      within[k0] = w0; within[k1] = w1;
    } else {
      const int a = a1 - 1, b = b1 - 1;
      const uint8_t wa = within[a], wb = within[b];
      within[a] = 1; within[b] = 1;
      pathRun(parent[a], p);
      within[a] = wa; within[b] = wb;
    }
  }

  // Groups.  Evaluated level by level, every level of the dirty tree costs a launch plus a full load -> compute ->
  // store round trip, and the upper levels of a tree hold a handful of tasks each: with few site tiles per GPU (a
  // site-sharded run) the device idles through two dozen dependent launches.  Instead the dirty stored vertices are
  // split into vertical PATHS: a vertex joins the path of the child whose path started later (higher round), or opens a
  // new round when both children's rounds are equal — the Strahler numbering of the dirty tree, so the number of
  // rounds is ~log4 of the number of stored vertices instead of the tree's height.  One launch per round; one block
  // column (gridDim.z) per path, walking its tasks bottom-up in program order: a thread only ever reads partials of its
  // own (site, rate), so inside a path the child -> parent dependency is the thread's own store followed by its own load.
  // groupCost caps the serial work of a path (stored vertex = 2 units, fused vertex = 1); 0 = one task per segment and
  // one launch per level (the round-1 schedule, kept for comparison).
  int groupCost = 128;
  int chainMin = CHAIN_MIN;              // trailing single-task levels are walked by k_walk from this length on; shorter ones are
                                         // an ordinary path of the last round(s)

  void plan(Plan& p) {
    allInvalid = false;
    p.tasks.clear(); p.tokens.clear(); p.tipList.clear(); p.levelStart.clear(); p.segStart.clear(); p.unitSeg.clear();
    p.storedReads = p.tipReads = p.fusedOps = p.fp64Ops = 0;
    p.chainLevels = p.materialised = 0;
    p.uses[0] = p.uses[1] = false;
    if (solved[root()]) { p.levelStart.push_back(0); p.segStart.push_back(0); p.unitSeg.push_back(0); return; }
    computeTips();
    PooledVec<int>& pre = p.pre;
    pre.clear();
    if (p.tokens.capacity() == 0) {          // first (full) plan of a handle: no growth by doubling through 100,000s of entries
      p.tokens.reserve((size_t)numTips + numTips / 4); p.tipList.reserve(numTips); pre.reserve(numTips);
    }
    std::vector<int> stack(1, root());
    while (!stack.empty()) {
      int v = stack.back(); stack.pop_back();
      pre.push_back(v);
      for (int i = 0; i < 2; i++) { int c = kids[2 * v + i]; if (!isTip(c) && !fused(c) && !solved[c]) stack.push_back(c); }
    }
    PooledVec<int>& level = p.level;       // height above the clean frontier; clean, fused vertices and tips = 0
    if ((int)level.size() != numVerts) level.assign(numVerts, 0);   // all zero between calls (reset below for the dirty set)
    int maxLevel = 0;
    for (size_t i = pre.size(); i-- > 0;) {   // reverse pre-order visits children before parents
      int v = pre[i];
      int l = 1 + std::max(level[kids[2 * v]], level[kids[2 * v + 1]]);
      level[v] = l;
      maxLevel = std::max(maxLevel, l);
    }
    p.cnt.assign(maxLevel + 1, 0);
    for (int v : pre) p.cnt[level[v]]++;
    // the trailing chain: levels chainL0..maxLevel hold one task each
    int chainL0 = maxLevel + 1;
    for (int l = maxLevel; l >= 1 && p.cnt[l] == 1; l--) chainL0 = l;
    p.chainLevels = maxLevel - chainL0 + 1;
    if (p.chainLevels < chainMin || !walkChains) { p.chainLevels = 0; chainL0 = maxLevel + 1; }
    std::vector<int> matList;                // fused clades hanging off the chain: stored tasks of a leading unit
    if (p.chainLevels)
      for (int v : pre)
        if (level[v] >= chainL0)
          for (int i = 0; i < 2; i++) { const int c = kids[2 * v + i]; if (!isTip(c) && fused(c)) { mat[c] = 1; matList.push_back(c); } }
    p.materialised = (int)matList.size();

    // ---- the vertices below the chain, grouped into paths and rounds
    PooledVec<int>& grp = p.grp;             // group of each dirty stored vertex below the chain
    if ((int)grp.size() != numVerts) grp.assign(numVerts, -1);
    PooledVec<int>&gRound = p.gRound, &gCost = p.gCost, &gHead = p.gHead, &gNext = p.gNext, &gTail = p.gTail;
    gRound.clear(); gCost.clear(); gHead.clear(); gTail.clear();
    if ((int)gNext.size() != numVerts) gNext.assign(numVerts, -1);   // members of a group: linked list, bottom-up
    int nRounds = 0;
    for (size_t i = pre.size(); i-- > 0;) {
      const int v = pre[i];
      if (level[v] >= chainL0) continue;
      int cost = 2, best = -1, other = -1;
      for (int j = 0; j < 2; j++) {
        const int c = kids[2 * v + j];
        if (isTip(c)) continue;
        if (fused(c)) { cost += tips[c] - 1; continue; }
        if (level[c] == 0) continue;         // clean stored child: already in HBM
        const int g = grp[c];
        if (best < 0) best = g;
        else if (gRound[g] > gRound[best] || (gRound[g] == gRound[best] && gCost[g] > gCost[best])) { other = best; best = g; }
        else other = g;
      }
      int round = 0;
      if (other >= 0) round = gRound[other] + 1;                                        // the path not taken closes below v
      // v extends the path of its later child — unless both children's paths are of the same round: pulling one of them
      // (all of it) into the next round would make two independent subtrees run one after the other
      const bool join = groupCost > 0 && best >= 0 && gCost[best] + cost <= groupCost && (other < 0 || gRound[best] >= round);
      if (join) {
        gCost[best] += cost;
        gNext[gTail[best]] = v; gTail[best] = v;
        grp[v] = best;
        nRounds = std::max(nRounds, gRound[best] + 1);
      } else {
        if (best >= 0) round = std::max(round, gRound[best] + 1);
        grp[v] = (int)gRound.size();
        gRound.push_back(round); gCost.push_back(cost); gHead.push_back(v); gTail.push_back(v);
        nRounds = std::max(nRounds, round + 1);
      }
    }
    const int nGroups = (int)gRound.size();
    // task order: materialised clades (one segment each), then round by round, group by group, bottom-up; then the chain
    std::vector<int> order;
    order.reserve(pre.size() + matList.size());
    p.unitSeg.push_back(0);
    for (int c : matList) { p.segStart.push_back((int)order.size()); order.push_back(c); }
    if (!matList.empty()) { p.levelStart.push_back(0); p.unitSeg.push_back((int)p.segStart.size()); }
    {
      PooledVec<int>& cnt = p.cnt;            // groups per round, then their order
      cnt.assign(nRounds + 1, 0);
      for (int g = 0; g < nGroups; g++) cnt[gRound[g] + 1]++;
      for (int r = 0; r < nRounds; r++) cnt[r + 1] += cnt[r];
      PooledVec<int>& byRound = p.fill;
      byRound.assign(nGroups, 0);
      std::vector<int> pos(cnt.begin(), cnt.end() - 1);
      for (int g = 0; g < nGroups; g++) byRound[pos[gRound[g]]++] = g;
      for (int r = 0; r < nRounds; r++) {
        // longest paths first (blocks are dispatched in segment order): the round's tail is then made of short ones
        std::stable_sort(byRound.begin() + cnt[r], byRound.begin() + cnt[r + 1], [&](int a, int b) { return gCost[a] > gCost[b]; });
        p.levelStart.push_back((int)order.size());
        for (int q = cnt[r]; q < cnt[r + 1]; q++) {
          p.segStart.push_back((int)order.size());
          for (int v = gHead[byRound[q]]; v != -1; v = gNext[v]) order.push_back(v);
        }
        p.unitSeg.push_back((int)p.segStart.size());
      }
    }
    p.segStart.push_back((int)order.size());
    if (p.chainLevels) {
      const size_t base = order.size();
      order.resize(base + p.chainLevels);
      for (int v : pre) if (level[v] >= chainL0) order[base + (level[v] - chainL0)] = v;
      for (int l = 0; l < p.chainLevels; l++) p.levelStart.push_back((int)base + l);
    }
    p.levelStart.push_back((int)order.size());
    p.tasks.resize(order.size());
    for (size_t i = 0; i < order.size(); i++) {
      const int v = order[i];
      const int outSlot = cur[v] ^ 1;
      Task k = makeTask(v, v - numTips, outSlot, p, true);   // materialised clades come first: the chain sees their new slot
      if (level[v] >= chainL0) {                             // a chain task (materialised clades have level 0)
        if (level[v] == chainL0) k.flags |= TF_HEAD;
        else k.flags |= level[kids[2 * v]] == level[v] - 1 ? TF_C0FWD : TF_C1FWD;   // the one dirty vertex of the level below
      }
      p.tasks[i] = k;
      p.uses[outSlot] = true;
      cur[v] = (uint8_t)outSlot;                      // ComputeSolutions: _previousIterVec <- current (:28) ...
      touched[v] = 1;                                 // ... _updateFlag = true (:34)
      solved[v] = 1;                                  // ... _isSolved = true (:33)
    }
    for (int c : matList) mat[c] = 0;
    for (int v : pre) { level[v] = 0; grp[v] = -1; gNext[v] = -1; }
  }
};

}  // namespace dmpc
