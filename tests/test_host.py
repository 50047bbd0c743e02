"""Host-side logic that needs no GPU: tree utilities, synthetic generators, the restated MCMC driver (over the
oracle), priors."""
import math

import numpy as np
import pytest

from dmphyclus_b200 import chain, synth
from dmphyclus_b200.phylo import Phylo, check_clades, cluster_mrcas
from dmphyclus_b200.workloads import Problem


@pytest.mark.parametrize("shape", ["random", "balanced", "caterpillar"])
@pytest.mark.parametrize("n", [2, 3, 17, 256])
def test_random_tree_is_ape_style(n, shape):
    e = synth.random_tree(n, np.random.default_rng(n), shape)
    assert e.shape == (2 * n - 2, 2) and e[0, 0] == n + 1
    assert sorted(e[:, 1]) == [v for v in range(1, 2 * n) if v != n + 1]          # every non-root vertex has one parent
    assert np.all(np.bincount(e[:, 0])[n + 1:] == 2) and np.all(e[:, 0] > n)     # strictly bifurcating
    phy = Phylo(e, n)
    assert sorted(phy.descendants(phy.root)) == list(range(1, n + 1))
    if shape == "caterpillar":
        assert phy.depth().max() == n - 1


def test_clusters_are_disjoint_clades():
    n, k = 200, 9
    e = synth.random_tree(n, np.random.default_rng(5))
    mrcas, clus = synth.pick_clusters(e, n, k)
    phy = Phylo(e, n)
    assert len(mrcas) == k and set(clus) == set(range(1, k + 1)) and check_clades(phy, clus)
    assert cluster_mrcas(phy, clus) == mrcas
    assert cluster_mrcas(phy, np.ones(n, int)) == [n]        # the reference's single-cluster quirk (DMphyClusSource.R:78-80)


def test_alignments(packaged):
    rng = np.random.default_rng(3)
    a = synth.iid_alignment(50, 4000, rng)
    assert set(np.unique(a)) <= set(b"atcgnryswkm")
    assert 0.005 < (a == ord("n")).mean() < 0.02
    e = synth.random_tree(30, rng)
    mrcas, _ = synth.pick_clusters(e, 30, 3)
    b = synth.evolve_alignment(e, 30, 500, packaged["within"][4], packaged["between"][4], mrcas, synth.LIM_PROBS, rng)
    assert b.shape == (30, 500) and set(np.unique(b)) <= set(b"atcg")
    assert (b == b[0]).all(0).mean() > 0.5        # short branches: most columns are constant


def test_clus_ind_log_prior():
    """clusIndLogPrior (DMphyClusInterface.R:147-162) against the closed form for two clusters."""
    c = np.array([1, 1, 1, 2, 2])
    a = 2.5
    want = (math.lgamma(6) - math.lgamma(4) - math.lgamma(3)
            + math.lgamma(a + 3) + math.lgamma(a + 2) - math.lgamma(2 * a + 5) - (2 * math.lgamma(a) - math.lgamma(2 * a)))
    assert abs(chain.clusIndLogPrior(c, a) - want) < 1e-12
    assert chain.clusIndLogPrior(np.array([7, 7, 7, 9, 9]), a) == chain.clusIndLogPrior(c, a)


def test_chain_runs_and_is_reproducible(oracle, packaged):
    p = Problem(packaged, 30, 80, 4, 3, "evolved", seed=8)
    st = chain.ChainSettings(limProbs=p.pi, withinTransMatAll=p.W, betweenTransMatAll=p.B, shapeForAlpha=1000,
                             scaleForAlpha=0.1, alphaMin=10, poisRateNumClus=2, numSplitMergeMoves=3)
    al = oracle.getConvertedAlignment(list("atcg"), p.chars)
    a, calls = chain.run_chain(oracle, al, p.edge, p.clus, 100.0, st, 12, seed=4)
    b, _ = chain.run_chain(oracle, al, p.edge, p.clus, 100.0, st, 12, seed=4)
    assert calls >= 12 * 6
    for x, y in zip(a, b):
        assert np.array_equal(x["clusInd"], y["clusInd"]) and x["logLik"] == y["logLik"]
    for x in a:       # every sample is a valid clade partition of the sampled tree, and its logLik is the true one
        assert check_clades(Phylo(x["edge"], p.n), x["clusInd"])
    last = a[-1]
    phy = Phylo(last["edge"], p.n)
    r = oracle.logLikCpp(last["edge"], last["clusterNodeIndices"], p.pi, p.W[last["withinMatListIndex"] - 1],
                         p.B[last["betweenMatListIndex"] - 1], 1, al, p.n, p.S, 1, 1)
    oracle.finalDeallocate(r["solutionPointer"])
    assert abs(r["logLik"] - last["logLik"]) <= 1e-12 * abs(last["logLik"])


def test_chain_bookkeeping_invariants(oracle, packaged):
    """What the driver's shortcuts rely on, checked after every iteration: clusterCounts is table(clusInd) with labels
    1..K, cluster k's tips are exactly the tips below clusterNodeIndices[k] (so its size is the length of
    phangorn::Descendants of its MRCA, DMphyClusSource.R:384-388), and the cached tree view follows accepted NNIs."""
    p = Problem(packaged, 60, 40, 6, 3, "evolved", seed=15)
    st = chain.ChainSettings(limProbs=p.pi, withinTransMatAll=p.W, betweenTransMatAll=p.B, shapeForAlpha=1000,
                             scaleForAlpha=0.1, alphaMin=10, poisRateNumClus=2, numSplitMergeMoves=3)
    al = oracle.getConvertedAlignment(list("atcg"), p.chars)
    seen = {"n": 0, "edges": set()}

    def check(it, cv):
        labels, cnt = np.unique(cv.clusInd, return_counts=True)
        assert list(labels) == list(range(1, len(labels) + 1)) == sorted(cv.clusterCounts)
        assert [cv.clusterCounts[int(l)] for l in labels] == list(cnt)
        phy = Phylo(cv.edge, cv.n_tips)
        assert np.array_equal(cv.phylo().parent, phy.parent)
        for k, m in enumerate(cv.clusterNodeIndices):
            tips = np.sort(phy.descendants(m))
            assert np.array_equal(tips, np.nonzero(cv.clusInd == k + 1)[0] + 1)
            assert len(tips) == cv.clusterCounts[k + 1]
        seen["n"] += 1
        seen["edges"].add(cv.edge.tobytes())

    chain.run_chain(oracle, al, p.edge, p.clus, 100.0, st, 25, seed=6, on_iter=check)
    assert seen["n"] == 25 and len(seen["edges"]) > 1          # some NNI was accepted along the way


def test_tree_view_fast_paths_agree():
    """Phylo.kids / descendants (vectorised child table) against the list-of-lists view, on bifurcating trees of every
    shape and on a tree with a polytomy (generic path)."""
    for shape in ("random", "balanced", "caterpillar"):
        rng = np.random.default_rng(5)
        e = synth.random_tree(40, rng, shape)
        a, b = Phylo(e, 40), Phylo(e, 40)
        full = b.children                      # forces the list view on b
        for v in range(1, 2 * 40):
            assert a.kids(v) == full[v]
            assert np.array_equal(a.descendants(v), b.descendants(v))
        assert a._children is None             # the fast paths never built the full list
    poly = np.array([[4, 1], [4, 2], [4, 3]])
    p = Phylo(poly, 3)
    assert p.kids(4) == [1, 2, 3] and list(p.descendants(4)) == [1, 2, 3]


def test_wrapper_converts_read_only_inputs_once(packaged):
    """engine.Engine marshals a transition-matrix list / edge matrix once when it is handed the same READ-ONLY array
    again (the MCMC driver proposes from fixed lists and rolls back to one committed edge matrix); writable arrays are
    converted on every call, so editing them in place is always seen."""
    from dmphyclus_b200 import build
    from dmphyclus_b200.engine import Engine
    build.build()
    eng = Engine()
    w = np.array(packaged["within"][4])
    p1 = eng._mp(w)
    w[0, 0, 1] += 0.25
    p2 = eng._mp(w)
    assert p1._arr is not p2._arr and p2._arr[4] == w[0, 0, 1]          # column-major: P(0,1) sits at [0 + 4*1]
    w.setflags(write=False)
    assert eng._mp(w) is eng._mp(w)
    edge = np.array([[4, 1], [4, 5], [5, 2], [5, 3]], dtype=np.int32)
    (a, n), (b, _) = eng._ep(edge), eng._ep(edge)
    assert n == 4 and a is not b
    edge.setflags(write=False)
    assert eng._ep(edge)[0] is eng._ep(edge)[0] and list(eng._ep(edge)[0]._arr) == [4, 4, 5, 5, 1, 5, 2, 3]


def test_workload_shapes_match_baseline():
    from dmphyclus_b200.workloads import CONFIGS
    assert CONFIGS["config2"] == (1000, 10000, 20, 3) and CONFIGS["config3"] == (10000, 30000, 100, 3)
    assert CONFIGS["config4"][:2] == (5000, 5000) and CONFIGS["config5"][:2] == (100000, 1500)


@pytest.mark.parametrize("shape", ["random", "balanced", "caterpillar"])
@pytest.mark.parametrize("fuse", [1, 2, 3, 4, 8, 15])
def test_plan_builder_invariants(shape, fuse):
    """The host-side plan (stored tasks per level + fused-clade token programs + the walked chain) over many trees:
    every vertex computed exactly once, children before parents, stack depth and staging sizes within what k_prune
    assumes; the trailing chain is one task per level forwarding into the next, with the fused clades hanging off it
    materialised in level 0.  Checked for full evaluations and for root-path proposals from many vertices."""
    import ctypes as C
    from dmphyclus_b200 import build
    from dmphyclus_b200.engine import load_library
    build.build()
    lib = load_library()
    chains = mats = 0
    for n in (2, 3, 5, 16, 17, 64, 257, 1000):
        rng = np.random.default_rng(n * 31 + fuse)
        e = synth.random_tree(n, rng, shape)
        mrcas, _ = synth.pick_clusters(e, n, min(5, n))
        edge = np.ascontiguousarray(e.T, dtype=np.int32).ravel()
        m = np.asarray(mrcas, np.float64)
        out = (C.c_int * 8)()

        def check(dirty):
            rc = lib.dmpc_plan_selfcheck(edge.ctypes.data_as(C.POINTER(C.c_int32)), len(edge) // 2, n,
                                         m.ctypes.data_as(C.POINTER(C.c_double)), len(m), fuse, dirty, out)
            assert rc == 0, (n, shape, fuse, dirty, lib.dmpc_last_error().decode())
            return list(out)

        stored, fused, levels, tokens, depth, tips, chain_levels, materialised = check(0)
        assert stored + fused == n - 1 and depth <= 3 and tips <= 30
        assert chain_levels == 0 or chain_levels >= 4
        if fuse == 1:
            assert fused == 0 and tokens == 0 and materialised == 0
        if shape == "caterpillar" and n > 40:
            assert stored >= n - 1 - fuse          # only the bottom of a caterpillar can be fused
            assert chain_levels >= stored - 2      # ... and its spine is one long chain
        # proposals: the root path from a sample of internal vertices (InvalidateSolution, IntermediateNode.cpp:5-13)
        internal = np.arange(n + 1, 2 * n)
        for v in (internal if n <= 64 else rng.choice(internal, 48, replace=False)):
            o = check(int(v))
            assert o[6] == 0 or (o[6] >= 4 and o[6] <= o[2])
            if o[6]:
                assert o[6] + (1 if o[7] else 0) == o[2]      # a single path: everything but the materialised level is chain
            chains += o[6] > 0
            mats += o[7]
        if n >= 5:
            o = check(-1)                                     # new between-cluster matrices: the top of the tree
            assert o[6] == 0 or o[6] >= 4
            for pick in range(12):                            # NNI moves: two dirty paths that meet below the root
                o = check(-2 - pick)
                assert o[6] == 0 or (o[6] >= 4 and o[6] <= o[2])
    if fuse > 1 and shape == "random":
        assert chains > 0 and mats > 0


def test_driver_helpers_against_their_plain_forms():
    """The vectorised helpers of the driver mirror against the term-by-term forms they replace: sibling cluster pairs as
    split(clusterNodeIndices, f = parent) gives them (/root/reference/R/DMphyClusSource.R:157-165), and the
    Dirichlet-multinomial prior from the cluster counts (R/DMphyClusInterface.R:147-162)."""
    from scipy.special import gammaln
    rng = np.random.default_rng(3)
    for trial in range(40):
        n = int(rng.integers(8, 400))
        edge = synth.random_tree(n, rng, ["random", "balanced", "caterpillar"][trial % 3])
        phy = Phylo(edge, n)
        k = int(rng.integers(1, 30))
        mrcas = [int(x) for x in rng.choice(np.arange(1, 2 * n), size=min(k, 2 * n - 2), replace=False) if x != n + 1]
        if not mrcas:
            continue
        groups = {}
        for m in mrcas:
            groups.setdefault(int(phy.parent[m]), []).append(m)
        want = [groups[p] for p in sorted(groups) if len(groups[p]) > 1]
        assert chain._pairs_and_splits(phy, mrcas, None)[0] == want
        assert chain._num_pairs(phy, mrcas) == len(want)
        counts = rng.integers(1, 50, size=len(mrcas)).astype(float)
        alpha = float(rng.uniform(10.0, 120.0))
        nn = int(counts.sum())
        plain = (gammaln(nn + 1) - gammaln(counts + 1).sum() + (gammaln(alpha + counts).sum() - gammaln((alpha + counts).sum()))
                 - (len(counts) * gammaln(alpha) - gammaln(len(counts) * alpha)))
        got = chain._log_prior_of_counts(list(counts), nn, alpha)
        assert abs(got - plain) <= 1e-10 * max(1.0, abs(plain))
    # three clusters under one parent cannot happen in a binary tree, but the grouping must not lose one if it did
    class Star:
        parent = np.array([0, 9, 9, 9, 8, 8, 7, 0, 7, 7])
    assert chain._pairs_and_splits(Star, [4, 1, 2, 5, 3, 6], None)[0] == [[4, 5], [1, 2, 3]]
    assert chain._num_pairs(Star, [4, 1, 2, 5, 3, 6]) == 2


def test_split_below_matches_children_and_descendants():
    """Phylo.split_below (the tips of a clade sorted by the child of its root they hang under) against
    Children + Descendants, on both of its paths: the traversal (small tree / table present) and the climb."""
    rng = np.random.default_rng(8)
    for n, shape in ((30, "random"), (700, "caterpillar"), (5000, "random"), (6000, "balanced")):
        edge = synth.random_tree(n, rng, shape)
        for trial in range(12):
            ref = Phylo(edge, n)
            node = int(rng.integers(n + 1, 2 * n))
            tips = np.sort(ref.descendants(node))
            want_kids = ref.kids(node)
            want = [set(ref.descendants(k).tolist()) for k in want_kids]
            phy = Phylo(edge, n)                                        # a fresh view: no table yet
            kids, parts = phy.split_below(node, tips)
            assert kids == want_kids and [set(p.tolist()) for p in parts] == want
            assert sum(len(p) for p in parts) == len(tips)
            if 2 * n > 8192:
                assert phy._table is None                               # the climb built no table
                with pytest.raises(ValueError):
                    outside = np.setdiff1d(np.arange(1, n + 1), tips)
                    if len(outside) == 0:
                        raise ValueError("whole tree")
                    phy.split_below(node, np.append(tips, outside[0]))  # a tip that is not below the vertex
