// host_bench.cpp — the engine's HOST side alone (csrc/host_tree.hpp: no CUDA), timed and fuzzed on a CPU box.
//
//   g++ -O2 -std=c++17 -o /tmp/host_bench tools/host_bench.cpp && /tmp/host_bench            (timings)
//   g++ -O1 -g -std=c++17 -pthread -fsanitize=address,undefined -o /tmp/host_fuzz tools/host_bench.cpp && /tmp/host_fuzz --fuzz 2000 1 4
//
// Timings: what logLikCpp spends on the host before the first pruning launch can be queued — link + validate the tree
// (HostTree::build), cluster flags (associate), the plan of a full evaluation (plan) — and what a proposal costs
// (NNI swaps + plan + edge export + rollback), for the BASELINE.json shapes.
// Fuzz: random proposal sequences (all six kinds, accepted and rejected, NNI rollbacks through the undo log and through
// a full relink) on random / balanced / caterpillar trees and every fusion threshold, plus malformed edge matrices
// through build(); run under AddressSanitizer / UBSan it checks the flat-array state machine for out-of-range accesses,
// and it asserts the invariants the kernels rely on after every plan (every dirty vertex computed once, children
// before parents, tips below a vertex consistent with a fresh traversal).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>

#include "../dmphyclus_b200/csrc/host_tree.hpp"

using namespace dmpc;
using Clock = std::chrono::steady_clock;
static double us(Clock::time_point a) { return std::chrono::duration<double, std::micro>(Clock::now() - a).count(); }

// rooted binary tree with n tips as R's phylo$edge (column-major, 1-based, root = n + 1, rows in pre-order)
static std::vector<int32_t> makeTree(int n, std::mt19937_64& rng, int shape) {
  std::vector<int> left(n - 1), right(n - 1);
  int root;
  if (shape == 0) {                       // random pairwise joining
    std::vector<int> act(n);
    for (int i = 0; i < n; i++) act[i] = i;
    for (int k = 0; k < n - 1; k++) {
      const int m = (int)act.size();
      int i = (int)(rng() % m), j = (int)(rng() % (m - 1));
      if (j >= i) j++;
      left[k] = act[i]; right[k] = act[j];
      act[i] = n + k; act[j] = act.back(); act.pop_back();
    }
    root = act[0];
  } else if (shape == 1) {                // balanced
    std::vector<int> level(n);
    for (int i = 0; i < n; i++) level[i] = i;
    std::shuffle(level.begin(), level.end(), rng);
    int k = 0;
    while (level.size() > 1) {
      std::vector<int> nxt;
      for (size_t a = 0; a + 1 < level.size(); a += 2) { left[k] = level[a]; right[k] = level[a + 1]; nxt.push_back(n + k); k++; }
      if (level.size() % 2) nxt.push_back(level.back());
      level.swap(nxt);
    }
    root = level[0];
  } else {                                // caterpillar
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::shuffle(order.begin(), order.end(), rng);
    int cur = order[0];
    for (int k = 0; k < n - 1; k++) { left[k] = cur; right[k] = order[k + 1]; cur = n + k; }
    root = cur;
  }
  const int rows = 2 * n - 2;
  std::vector<int32_t> edge(2 * (size_t)rows);
  std::vector<int> id(2 * n - 1);
  for (int i = 0; i < n; i++) id[i] = i + 1;
  int nxt = n + 1, row = 0;
  std::vector<std::pair<int, int>> st(1, {root, 0});
  while (!st.empty()) {
    auto [v, par] = st.back(); st.pop_back();
    if (v >= n) id[v] = nxt++;
    if (par) { edge[row] = par; edge[row + rows] = id[v]; row++; }
    if (v >= n) { st.push_back({right[v - n], id[v]}); st.push_back({left[v - n], id[v]}); }
  }
  return edge;
}

// disjoint clades as cluster MRCAs (1-based): walk down from the root, cutting where a subtree has <= n / k tips
static std::vector<double> makeClusters(const HostTree& ht, int k) {
  std::vector<double> m;
  const int cap = std::max(1, ht.numTips / std::max(1, k));
  std::vector<int> st(1, ht.root());
  while (!st.empty()) {
    int v = st.back(); st.pop_back();
    if (ht.isTip(v) || ht.tips[v] <= cap) { m.push_back(v + 1); continue; }
    st.push_back(ht.kids[2 * v]); st.push_back(ht.kids[2 * v + 1]);
  }
  return m;
}

#define CHECK(c, msg) do { if (!(c)) { std::fprintf(stderr, "host_bench: %s (%s:%d)\n", msg, __FILE__, __LINE__); std::abort(); } } while (0)

// the invariants of a plan that do not need the kernels: every planned vertex was dirty and reachable, exactly once;
// operands are tips, clean stored vertices, fused clades, or tasks that come earlier in the plan
static void checkPlan(const HostTree& ht, const HostTree::Plan& p, const std::vector<uint8_t>& solvedBefore) {
  const int N = ht.numTips;
  std::vector<int> at(ht.numVerts, -1);
  for (size_t i = 0; i < p.tasks.size(); i++) {
    const Task& k = p.tasks[i];
    CHECK(k.out >= 0 && k.out < ht.numVerts - N, "task output out of range");
    CHECK(at[k.out + N] < 0, "vertex planned twice");
    at[k.out + N] = (int)i;
  }
  std::vector<uint8_t> tok(ht.numVerts, 0);
  for (const Token& t : p.tokens)
    if ((t.op & 7) != OP_SPILL) { CHECK(t.vertex >= 0 && t.vertex < ht.numVerts - N, "token vertex out of range"); CHECK(!tok[t.vertex + N], "fused vertex computed twice"); tok[t.vertex + N] = 1; }
  for (size_t i = 0; i < p.tasks.size(); i++) {
    const Task& k = p.tasks[i];
    const int v = k.out + N;
    CHECK(!solvedBefore[v] || ht.fused(v), "a clean stored vertex was planned");
    for (int j = 0; j < 2; j++) {
      const int c = ht.kids[2 * v + j];
      const bool tip = k.flags & (j ? TF_C1TIP : TF_C0TIP), fus = k.flags & (j ? TF_C1FUSED : TF_C0FUSED);
      CHECK(tip == ht.isTip(c), "tip flag");
      if (tip) { CHECK((j ? k.c1 : k.c0) == c, "tip operand"); continue; }
      CHECK((j ? k.c1 : k.c0) == c - N, "internal operand");
      if (fus) { CHECK(ht.fused(c) && tok[c], "fused operand without a program"); continue; }
      if (at[c] >= 0) CHECK(at[c] < (int)i, "operand planned after its consumer");
      else CHECK(solvedBefore[c], "dirty operand not planned");
    }
    CHECK(k.tipOff >= 0 && k.tipOff + k.nTips <= (int)p.tipList.size(), "tip slice");
    CHECK(k.prog >= 0 && k.prog + (k.ntok & 0xffff) + (k.ntok >> 16) <= (int)p.tokens.size(), "token slice");
  }
  CHECK(ht.solved[ht.root()], "root not solved after the plan");
  for (int v = N; v < ht.numVerts; v++) CHECK(ht.mat[v] == 0, "materialisation mark left behind");
}

// the one-pass flagging over the committed edge matrix against the per-cluster traversal, on an arbitrary MRCA list
// (nested clades, tips and out-of-range ids included: both must treat them alike)
static void checkAssociate(HostTree& ht, std::mt19937_64& rng) {
  const std::vector<uint8_t> keep(ht.within.begin(), ht.within.end());
  std::vector<double> ml;
  const int k = 1 + (int)(rng() % 9);
  for (int i = 0; i < k; i++) ml.push_back((double)(rng() % (ht.numVerts + 3)));
  ht.associate(ml.data(), (int)ml.size());
  const std::vector<uint8_t> fast(ht.within.begin(), ht.within.end());
  std::fill(ht.within.begin(), ht.within.end(), 0);
  ht.associateByTraversal(ml.data(), (int)ml.size());
  CHECK(std::equal(fast.begin(), fast.end(), ht.within.begin()), "one-pass cluster flags differ from the traversal's");
  for (int v = ht.numTips; v < ht.numVerts; v++) CHECK(ht.mat[v] == 0, "MRCA marks left behind");
  std::copy(keep.begin(), keep.end(), ht.within.begin());
}

static void checkTips(HostTree& ht) {
  if (!ht.tipsValid) return;
  std::vector<int32_t> have = ht.tips;
  ht.tipsValid = false;
  ht.computeTips();
  CHECK(have == ht.tips, "incremental tip counts differ from a fresh traversal");
}

// PooledVec: storage is recycled, contents never are; copies are deep; small vectors stay out of the pool
static void checkPool() {
  const size_t big = 300000;
  const int* block = nullptr;
  {
    PooledVec<int> a;
    a.assign(big, 7);
    block = a.data();
    PooledVec<int> b(a), c;
    c = a;
    CHECK(b.data() != a.data() && c.data() != a.data() && b == a && c == a, "copies are deep");
    b[5] = 9;
    CHECK(a[5] == 7, "copies are deep");
  }
  {
    PooledVec<int> d;
    CHECK(d.empty() && d.capacity() == 0, "a fresh vector holds nothing until it grows");
    d.assign(big, 3);                            // (single-threaded here: one of the three parked blocks comes back)
    CHECK(d.capacity() >= big && d.size() == big, "assign");
    for (size_t i = 0; i < big; i += 997) CHECK(d[i] == 3, "recycled storage must not leak old contents");
    PooledVec<int> e;
    e.resize(10);
    e[3] = 1;
    e.resize(big);                               // growth keeps what was there, like any vector
    CHECK(e[3] == 1 && e[big - 1] == 0, "resize keeps the contents and zero-fills the rest");
    std::vector<int> plain(5, 2);
    e = plain;
    CHECK(e.size() == 5 && e[4] == 2, "assignment from a plain vector");
  }
  (void)block;
}

static int fuzz(int rounds, uint64_t seed) {
  std::mt19937_64 rng(seed);
  checkPool();
  long plans = 0, moves = 0;
  for (int r = 0; r < rounds; r++) {
    // (every 50th tree is large enough for its vectors to go through the recycled-storage pool)
    const int n = r % 50 == 7 ? 20000 + (int)(rng() % 20000) : 2 + (int)(rng() % (r % 16 == 0 ? 3000 : 120));
    const int shape = (int)(rng() % 3);
    std::vector<int32_t> edge = makeTree(n, rng, shape);
    const int rows = 2 * n - 2;
    if (n <= 3000 && rng() % 4 == 0) {          // rows in another order: the traversal fallbacks
      std::vector<int> perm(rows);
      for (int i = 0; i < rows; i++) perm[i] = i;
      std::shuffle(perm.begin(), perm.end(), rng);
      std::vector<int32_t> e2(edge.size());
      for (int i = 0; i < rows; i++) { e2[i] = edge[perm[i]]; e2[i + rows] = edge[perm[i] + rows]; }
      // children keep their order only if both rows of a parent stay in relative order: sort pairs back
      for (int i = 0; i < rows; i++)
        for (int j = i + 1; j < rows; j++)
          if (e2[i] == e2[j] && perm[i] > perm[j]) { std::swap(e2[i + rows], e2[j + rows]); std::swap(perm[i], perm[j]); }
      edge.swap(e2);
    }
    HostTree ht;
    CHECK(ht.build(edge.data(), rows, n), "valid tree refused");
    ht.fuseTips = 1 + (int)(rng() % FUSE_MAX_TIPS);
    ht.groupCost = rng() % 5 == 0 ? 0 : (rng() % 3 == 0 ? 8 + (int)(rng() % 64) : 128);
    ht.computeTips();
    std::vector<double> mr = makeClusters(ht, 1 + (int)(rng() % 12));
    ht.associate(mr.data(), (int)mr.size());
    checkAssociate(ht, rng);
    HostTree::Plan p;
    std::vector<uint8_t> before(ht.solved.begin(), ht.solved.end());
    ht.plan(p); plans++;
    checkPlan(ht, p, before);
    ht.negateAllUpdateFlags();
    std::vector<int32_t> committed(edge.size()), proposed(edge.size());
    ht.exportEdges(committed.data());
    for (int step = 0; step < 12; step++) {
      const int kind = (int)(rng() % 6);
      std::vector<uint8_t> curBefore = ht.cur, withinBefore = ht.within;
      ht.negateAllUpdateFlags();
      bool nni = false, sm = false;
      if (kind == 0) ht.invalidateAll();
      else if (kind == 1) ht.invalidateBetween();
      else if (kind == 2) {                       // split
        int m = -1;
        for (double x : mr) if ((int)x > n && rng() % 2) { m = (int)x - 1; break; }
        if (m < 0) continue;
        ht.invalidateSolution(m);
        ht.within[ht.kids[2 * m]] = 0; ht.within[ht.kids[2 * m + 1]] = 0;
        sm = true;
      } else if (kind == 3) {                     // merge of two sibling clusters, if any
        int a = -1, b = -1;
        for (size_t i = 0; i < mr.size() && a < 0; i++)
          for (size_t j = i + 1; j < mr.size(); j++)
            if (ht.parent[(int)mr[i] - 1] >= 0 && ht.parent[(int)mr[i] - 1] == ht.parent[(int)mr[j] - 1]) { a = (int)mr[i] - 1; b = (int)mr[j] - 1; break; }
        if (a < 0) continue;
        ht.invalidateSolution(ht.parent[a]);
        ht.within[a] = 1; ht.within[b] = 1;
        sm = true;
      } else {                                    // NNI inside a cluster (4) or between the clusters (5)
        std::vector<long> ml;
        for (double x : mr) ml.push_back((long)x);
        int origin = ht.root();
        if (kind == 4) { origin = (int)mr[rng() % mr.size()] - 1; }
        else ht.markMrcas(ml);
        std::vector<int> cand;
        ht.nniCandidates(origin, kind == 5 ? &ml : nullptr, cand);
        if (cand.empty()) { ht.unmarkMrcas(); continue; }
        ht.beginNni();
        const int nm = 1 + (int)(rng() % 3);
        int last = -1;
        for (int m = 0; m < nm; m++) {
          const int root = cand[ht.rng.uniform_int(cand.size())];
          int two[2];
          ht.twoVerticesForNNI(root, kind == 5 ? &ml : nullptr, two);
          ht.rearrangeNNI(two[0], two[1]);
          last = root; moves++;
        }
        ht.unmarkMrcas();
        nni = true;
        before = ht.solved;
        ht.plan(p); plans++;
        checkPlan(ht, p, before);
        ht.exportProposed(proposed.data(), nm == 1 ? last : -1);
        std::vector<int32_t> fresh(edge.size());
        ht.exportEdges(fresh.data());
        CHECK(fresh == proposed, "incremental edge export differs from a full one");
        checkTips(ht);
      }
      if (!nni) {
        before = ht.solved;
        ht.plan(p); plans++;
        checkPlan(ht, p, before);
      }
      const bool accept = rng() % 3 == 0;
      if (accept) {
        if (nni) committed = proposed;
        if (rng() % 2) checkAssociate(ht, rng);      // (an accepted NNI is still pending here: the traversal path)
        continue;
      }
      // reject: RestorePreviousConfig
      if (nni) {
        std::vector<int32_t> back = committed;
        if (rng() % 4 == 0) {                     // hand back a permuted copy now and then: the full relink path
          // (swap two sibling rows' positions is not order-preserving; instead just force the relink by clearing the log)
          ht.nniLog.clear();
        }
        CHECK(ht.restoreTopology(back.data(), rows), "restore refused the committed matrix");
        std::vector<int32_t> now(edge.size());
        ht.exportEdges(now.data());
        CHECK(now == committed, "topology after the rollback is not the committed one");
        checkTips(ht);
      }
      {
        // the eight-at-a-time rollback against the plain loop over every vertex
        std::vector<int> want, got;
        for (int v = n; v < ht.numVerts; v++) if (ht.touched[v]) want.push_back(2 * v + (ht.cur[v] ^ 1));
        ht.rollbackTouched([&](int v, int slot) { got.push_back(2 * v + slot); });
        CHECK(got == want, "rollbackTouched visits other vertices (or other slots) than the plain loop");
        for (int v = 0; v < ht.numVerts; v++) CHECK(!ht.touched[v], "a touched flag survived the rollback");
      }
      CHECK(ht.cur == curBefore, "slot bits after the rollback");
      if (sm) ht.within = withinBefore;
      checkAssociate(ht, rng);
      // (the reference leaves _isSolved as the proposal set it: everything planned is solved)
    }
  }
  // malformed edge matrices: build() must refuse or accept them without touching memory out of range
  long refused = 0;
  for (int r = 0; r < rounds; r++) {
    const int n = 2 + (int)(rng() % 40);
    std::vector<int32_t> edge = makeTree(n, rng, (int)(rng() % 3));
    const int rows = 2 * n - 2;
    const int hits = 1 + (int)(rng() % 3);
    for (int h = 0; h < hits; h++) {
      const size_t at = rng() % edge.size();
      switch (rng() % 4) {
        case 0: edge[at] = (int32_t)(rng() % (2 * n + 3)) - 1; break;
        case 1: edge[at] = edge[rng() % edge.size()]; break;
        case 2: edge[at] = INT32_MAX - (int32_t)(rng() % 3); break;
        default: edge[at] = -(int32_t)(rng() % 5); break;
      }
    }
    HostTree ht;
    if (!ht.build(edge.data(), rows, n)) { refused++; continue; }
    ht.fuseTips = 1 + (int)(rng() % FUSE_MAX_TIPS);
    HostTree::Plan p;
    std::vector<uint8_t> before = ht.solved;
    ht.plan(p);                                   // whatever was accepted is a proper tree: it must plan
    checkPlan(ht, p, before);
  }
  std::printf("fuzz ok: %d trees, %ld plans, %ld NNI moves; %ld of %d damaged edge matrices refused\n", rounds, plans, moves, refused, rounds);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && !std::strcmp(argv[1], "--fuzz")) {
    // --fuzz ROUNDS [SEED [THREADS]]: several threads at once share the recycled-storage pool (PooledVec), as the shards
    // of a multi-device handle do when they are created and destroyed on their worker threads
    const int rounds = argc > 2 ? std::atoi(argv[2]) : 500, nThreads = argc > 4 ? std::atoi(argv[4]) : 1;
    const uint64_t seed = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 1;
    std::vector<std::thread> th;
    for (int i = 1; i < nThreads; i++) th.emplace_back([=] { fuzz(rounds, seed + 1000u * i); });
    const int rc = fuzz(rounds, seed);
    for (auto& t : th) t.join();
    return rc;
  }
  struct Shape { const char* name; int n, k, shape; } shapes[] = {
      {"config2  1,000 tips, 20 clusters", 1000, 20, 0},     {"config4  5,000 tips, 50 clusters", 5000, 50, 0},
      {"config3 10,000 tips, 100 clusters", 10000, 100, 0},  {"config5 100,000 tips, random", 100000, 200, 0},
      {"config5 100,000 tips, balanced", 100000, 200, 1},    {"config5 100,000 tips, caterpillar", 100000, 200, 2}};
  std::printf("%-36s %10s %10s %10s %10s | %10s %10s %10s\n", "shape (fuse 15)", "build us", "assoc us", "plan us", "replan us", "NNI us", "restore us", "tasks");
  for (const Shape& s : shapes) {
    std::mt19937_64 rng(2001);
    std::vector<int32_t> edge = makeTree(s.n, rng, s.shape);
    const int rows = 2 * s.n - 2, reps = s.n >= 100000 ? 20 : 200;
    double tb = 1e30, ta = 1e30, tp = 1e30, tr = 1e30;
    size_t tasks = 0;
    std::vector<double> mr;
    for (int rep = 0; rep < reps; rep++) {
      HostTree ht;
      ht.fuseTips = 15;
      auto t0 = Clock::now();
      CHECK(ht.build(edge.data(), rows, s.n), "build");
      tb = std::min(tb, us(t0));
      if (mr.empty()) { ht.computeTips(); mr = makeClusters(ht, s.k); }
      t0 = Clock::now();
      ht.associate(mr.data(), (int)mr.size());
      ta = std::min(ta, us(t0));
      HostTree::Plan p;
      t0 = Clock::now();
      ht.plan(p);
      tp = std::min(tp, us(t0));
      tasks = p.tasks.size();
      ht.negateAllUpdateFlags(); ht.invalidateAll();
      t0 = Clock::now();
      ht.plan(p);                                  // vectors already sized: what a re-plan costs
      tr = std::min(tr, us(t0));
    }
    // a within-cluster NNI proposal and its rollback, steady state
    HostTree ht;
    ht.fuseTips = 15;
    ht.build(edge.data(), rows, s.n);
    ht.associate(mr.data(), (int)mr.size());
    HostTree::Plan p;
    ht.plan(p);
    std::vector<int32_t> committed(edge.size()), proposed(edge.size());
    ht.exportEdges(committed.data());
    double tn = 0, tq = 0;
    int done = 0;
    for (int it = 0; it < 400; it++) {
      const int origin = (int)mr[it % mr.size()] - 1;
      auto t0 = Clock::now();
      ht.negateAllUpdateFlags();
      std::vector<int> cand;
      ht.nniCandidates(origin, nullptr, cand);
      if (cand.empty()) continue;
      ht.beginNni();
      const int root = cand[ht.rng.uniform_int(cand.size())];
      int two[2];
      ht.twoVerticesForNNI(root, nullptr, two);
      ht.rearrangeNNI(two[0], two[1]);
      ht.plan(p);
      ht.exportProposed(proposed.data(), root);
      tn += us(t0);
      t0 = Clock::now();
      CHECK(ht.restoreTopology(committed.data(), rows), "restore");
      ht.rollbackTouched([](int, int) {});
      tq += us(t0);
      done++;
    }
    std::printf("%-36s %10.1f %10.1f %10.1f %10.1f | %10.2f %10.2f %10zu\n", s.name, tb, ta, tp, tr, done ? tn / done : 0.0, done ? tq / done : 0.0, tasks);
  }
  return 0;
}
