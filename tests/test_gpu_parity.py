"""GPU parity tests proper: the CUDA engine, called through its C-ABI, against the CPU oracle on the same
seeded inputs.  Tolerances (BASELINE.json north_star): log-likelihood within 1e-9 relative in fp64;
partial-likelihood vectors and fp32 rescale exponents are compared bit-for-bit except for the exponents'
last fp32 bit (CUDA's double log() may differ from glibc's by 1 ulp before the rounding to float)."""
import numpy as np
import pytest

from dmphyclus_b200 import chain, synth
from helpers import Problem, rel

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _vignette(packaged, backend, mrcas, wi, bi):
    edge = np.array([[7, 8], [8, 9], [9, 1], [9, 2], [8, 3], [7, 10], [10, 11], [11, 4], [11, 5], [10, 6]])
    al = backend.getConvertedAlignment(list("atcg"), packaged["seqs"])
    r = backend.logLikCpp(edge, mrcas, synth.LIM_PROBS, packaged["within"][wi - 1], packaged["between"][bi - 1], 1, al,
                          6, 1617, wi, bi)
    backend.finalDeallocate(r["solutionPointer"])
    return r["logLik"]


@pytest.mark.parametrize("mrcas,wi,bi,expected", [
    ([8, 10], 5, 5, -3072.967109235698), ([8, 10], 1, 1, -3148.266004654241), ([8, 10], 10, 10, -3038.497338364486),
    ([8, 10], 5, 1, -3130.306408470979), ([7], 5, 5, -3510.074064616059), ([1, 2, 3, 4, 5, 6], 5, 5, -3259.751644946885)])
def test_vignette_known_answers(packaged, engine, oracle, mrcas, wi, bi, expected):
    """Config 1 data (seqsToCluster, packaged matrices): BASELINE.md §5 derived answers and the oracle."""
    got = _vignette(packaged, engine, mrcas, wi, bi)
    assert rel(got, expected) < TOL
    assert rel(got, _vignette(packaged, oracle, mrcas, wi, bi)) < 1e-13


@pytest.mark.parametrize("n,S,k,R,kind,shape", [
    (2, 5, 1, 3, "iid", "random"), (3, 1, 2, 1, "iid", "random"), (8, 130, 3, 3, "evolved", "random"),
    (64, 257, 6, 3, "iid", "balanced"), (300, 100, 5, 2, "iid", "caterpillar"), (700, 140, 8, 3, "iid", "random"),
    (700, 64, 8, 1, "iid", "random"), (1024, 48, 10, 3, "evolved", "random"), (1500, 40, 12, 5, "iid", "random"),
    (900, 33, 4, 8, "iid", "random"), (1024, 70, 6, 3, "diverged", "random"), (2000, 40, 8, 3, "diverged", "random"),
    (1200, 300, 8, 2, "diverged", "balanced"), (1500, 1100, 5, 3, "diverged", "random")])
def test_full_evaluation_matches_oracle(packaged, engine, oracle, n, S, k, R, kind, shape):
    """logLikCpp (full pruning + root reduction), including ragged site counts, R in 1..8, ambiguity codes,
    rescale-heavy iid data (fp32 exponents + cross-locus carry) and every tree shape."""
    if R > 3:   # more rate categories than the packaged lists hold: reuse matrices of other indices
        packaged = dict(packaged)
        packaged["within"] = np.concatenate([packaged["within"]] * 3, axis=1)[:, :R]
        packaged["between"] = np.concatenate([np.roll(packaged["between"], i, axis=0) for i in range(3)], axis=1)[:, :R]
    p = Problem(packaged, n, S, k, R, kind, shape, seed=n + S)
    hg, lg = p.create(engine)
    ho, lo = p.create(oracle)
    try:
        assert rel(lg, lo) < TOL, (lg, lo)
        for v in {n, n + (n - 1) // 2, 2 * n - 2}:          # root, a middle vertex, the last vertex
            sg, eg = engine.partial(hg, v)
            so, eo = oracle.partial(ho, v)
            assert np.array_equal(sg, so), f"partials of vertex {v} differ"
            assert np.allclose(eg, eo, rtol=2e-7, atol=0), f"exponents of vertex {v} differ"
    finally:
        engine.finalDeallocate(hg)
        oracle.finalDeallocate(ho)


def test_textbook_mode_matches_oracle(packaged, oracle):
    """quirk_carry = 0 (per-site exponent vector): the switchable 'textbook' reduction."""
    from dmphyclus_b200.engine import Engine
    from oracle.oracle import Oracle
    eng, orc = Engine(quirk_carry=False), Oracle(quirk_carry=False)
    p = Problem(packaged, 1024, 90, 6, 3, "diverged", seed=11)
    hg, lg = p.create(eng)
    ho, lo = p.create(orc)
    hq, lq = p.create(oracle)
    assert rel(lg, lo) < TOL
    assert rel(lo, lq) > 1e-6       # the carry quirk is visible on this input, so the two modes are distinguishable
    eng.finalDeallocate(hg); orc.finalDeallocate(ho); oracle.finalDeallocate(hq)


@pytest.mark.parametrize("n,S,k,R,kind,iters,seed", [
    (6, 300, 2, 3, "evolved", 25, 1), (40, 150, 5, 3, "evolved", 20, 2), (700, 40, 8, 3, "iid", 8, 3),
    (700, 40, 8, 1, "iid", 8, 4), (300, 50, 6, 3, "iid", 10, 5), (1024, 48, 8, 3, "diverged", 8, 6)])
def test_mcmc_trajectories_identical(packaged, engine, oracle, n, S, k, R, kind, iters, seed):
    """All six likelihood entry points + RestorePreviousConfig in the reference's move order
    (.performStepPhylo), same numpy stream for both backends: identical cluster-indicator and topology
    trajectories, every per-iteration log-likelihood within 1e-9."""
    p = Problem(packaged, n, S, k, R, kind, seed=seed)
    st = chain.ChainSettings(limProbs=p.pi, withinTransMatAll=p.W, betweenTransMatAll=p.B, shapeForAlpha=1000,
                             scaleForAlpha=0.1, alphaMin=10, poisRateNumClus=2, numSplitMergeMoves=3)
    outs = []
    for be in (engine, oracle):
        al = be.getConvertedAlignment(list("atcg"), p.chars)
        out, calls = chain.run_chain(be, al, p.edge, p.clus, 100.0, st, iters, seed=seed + 100)
        outs.append(out)
    for a, b in zip(*outs):
        assert np.array_equal(a["clusInd"], b["clusInd"])
        assert np.array_equal(a["edge"], b["edge"])
        assert a["withinMatListIndex"] == b["withinMatListIndex"] and a["betweenMatListIndex"] == b["betweenMatListIndex"]
        assert rel(a["logLik"], b["logLik"]) < TOL


def test_reject_restores_state(packaged, engine, oracle):
    """A rejected proposal leaves no trace: after RestorePreviousConfig a full recompute gives the original
    value, and the committed partials are the pre-proposal ones."""
    p = Problem(packaged, 200, 70, 5, 3, "iid", seed=21)
    h, l0 = p.create(engine)
    w, b = p.W[4], p.B[4]
    root_before = engine.partial(h, p.n)[0].copy()
    l1 = engine.newWithinTransProbsLogLik(h, p.W[0], b, 1, 1, p.pi)
    assert l1 != l0
    engine.RestorePreviousConfig(h, np.zeros((1, 2), np.int32), 5, 5, False, p.mrcas, False)
    assert np.array_equal(engine.partial(h, p.n)[0], root_before)
    r = engine.withinClusNNIlogLik(h, w, b, max(p.mrcas), 1, 1, p.pi)
    engine.RestorePreviousConfig(h, p.edge, 5, 5, True, p.mrcas, False)
    assert np.array_equal(engine.partial(h, p.n)[0], root_before)
    assert engine.recomputeAll(h, w, b, p.pi) == l0
    engine.finalDeallocate(h)


def test_error_paths(packaged, engine):
    from dmphyclus_b200.engine import DmpcError
    p = Problem(packaged, 10, 20, 2, 3, "iid", seed=5)
    al = engine.getConvertedAlignment(list("atcg"), p.chars)
    bad_edge = p.edge.copy(); bad_edge[0, 1] = 99
    with pytest.raises(DmpcError):
        engine.logLikCpp(bad_edge, p.mrcas, p.pi, p.W[4], p.B[4], 1, al, p.n, p.S, 5, 5)
    with pytest.raises(ValueError):
        engine.getConvertedAlignment(list("atcg"), np.frombuffer(b"acgtX", np.uint8).reshape(1, 5))
    al0 = al.copy(); al0[0, 0] = 0
    with pytest.raises(DmpcError):
        engine.logLikCpp(p.edge, p.mrcas, p.pi, p.W[4], p.B[4], 1, al0, p.n, p.S, 5, 5)
    h, _ = p.create(engine)
    with pytest.raises(DmpcError):   # a tip has no NNI candidate: error, not the reference's GSL abort
        engine.withinClusNNIlogLik(h, p.W[4], p.B[4], 1, 1, 1, p.pi)
    engine.finalDeallocate(h)


# ---------------------------------------------------------------------------------------------- golden fixtures
from helpers import GoldenCase, compare_to_golden, load_golden  # noqa: E402
from script import GOLDEN_CASES, VIGNETTE_CASES, run_script, vignette_loglik  # noqa: E402


@pytest.mark.parametrize("name", [c[0] for c in GOLDEN_CASES])
def test_golden_script_on_gpu(engine, name):
    """The scripted sequence over all entry points against what the reference's own TUs produced
    (tests/golden/ref_vectors.npz): integers and partial vectors bit-exact, log-likelihoods within 1e-9."""
    case = GoldenCase(load_golden(), name)
    compare_to_golden(run_script(engine, case), case.expected, ll_tol=TOL, exact_partials=True)


def test_golden_vignette_on_gpu(engine, packaged):
    for (m, wi, bi), want in zip(VIGNETTE_CASES, load_golden()["vignette/logliks"]):
        assert rel(vignette_loglik(engine, packaged, m, wi, bi), want) < 1e-13


# ---------------------------------------------------------------------------------------------- site sharding
@pytest.mark.parametrize("n,S,k,R,kind,parts", [(700, 1300, 6, 3, "iid", 2), (1024, 2100, 6, 3, "diverged", 4),
                                                 (300, 900, 4, 1, "iid", 3), (64, 700, 4, 3, "evolved", 2)])
def test_site_shards_reproduce_single_handle_bitwise(packaged, engine, n, S, k, R, kind, parts):
    """The multi-GPU protocol with every shard on this one GPU: contiguous CHUNK-aligned site blocks, the fp32
    chain handed from block to block, chunk sums added in the fixed order -> the SAME BITS as one handle."""
    from dmphyclus_b200.engine import Engine
    from dmphyclus_b200.sharded import CHUNK
    p = Problem(packaged, n, S, k, R, kind, seed=S)
    al = engine.getConvertedAlignment(list("atcg"), p.chars)
    h, want = p.create(engine)
    per = -(-(-(-S // parts)) // CHUNK) * CHUNK
    shards = []
    for i in range(parts):
        b0, b1 = min(S, i * per), min(S, (i + 1) * per)
        if b1 > b0:
            e = Engine(site_begin=b0, sites_total=S)
            r = e.logLikCpp(p.edge, p.mrcas, p.pi, p.W[4], p.B[4], 1, al[:, b0:b1], n, b1 - b0, 5, 5, deferred=True)
            assert np.isnan(r["logLik"])
            shards.append((e, r["solutionPointer"]))

    def reduce():
        carry = np.zeros(R, np.float32)
        chunks = []
        for e, hh in shards:
            carry = e.shard_chain(hh, carry)
        for e, hh in shards:
            chunks.append(e.shard_finish(hh))
        return engine.reduce_chunks(np.concatenate(chunks))

    assert reduce() == want
    # a proposal on every shard, then the same exchange
    want2 = engine.newBetweenTransProbsLogLik(h, p.W[4], p.B[7], 1, 8, p.pi)
    for e, hh in shards:
        e.newBetweenTransProbsLogLik(hh, p.W[4], p.B[7], 1, 8, p.pi)
    assert reduce() == want2
    assert np.array_equal(np.concatenate([e.site_logliks(hh) for e, hh in shards]), engine.site_logliks(h))
    for e, hh in shards:
        e.finalDeallocate(hh)
    engine.finalDeallocate(h)


# ---------------------------------------------------------------------------------------------- full size
@pytest.mark.parametrize("kind", ["evolved", "iid"])
def test_config2_full_size_against_oracle(packaged, engine, kind):
    """BASELINE.json config 2 at its full size (1,000 seqs x 10,000 sites, 20 clusters, R = 3): the dense CPU
    oracle still finishes in seconds there, so this is a direct comparison, plus size-independent properties."""
    from oracle.oracle import Oracle
    p = Problem(packaged, 1000, 10000, 20, 3, kind, seed=2001)
    orc = Oracle(threads=4)
    hg, lg = p.create(engine)
    ho, lo = p.create(orc)
    assert rel(lg, lo) < TOL, (lg, lo)
    sg, eg = engine.partial(hg, p.n)
    so, eo = orc.partial(ho, p.n)
    assert np.array_equal(sg, so)
    assert np.allclose(eg, eo, rtol=2e-7, atol=0)
    site = engine.site_logliks(hg)
    assert abs(site.sum() - lg) <= 1e-12 * abs(lg)              # the total is the sum of the per-site terms
    assert engine.recomputeAll(hg, p.W[4], p.B[4], p.pi) == lg   # idempotent and deterministic
    # a proposal that is rolled back leaves the value unchanged; one that is kept matches the oracle
    a = engine.newWithinTransProbsLogLik(hg, p.W[1], p.B[4], 1, 2, p.pi)
    b = orc.newWithinTransProbsLogLik(ho, p.W[1], p.B[4], 1, 2, p.pi)
    assert rel(a, b) < TOL
    engine.RestorePreviousConfig(hg, np.zeros((1, 2), np.int32), 5, 5, False, p.mrcas, False)
    assert engine.recomputeAll(hg, p.W[4], p.B[4], p.pi) == lg
    engine.finalDeallocate(hg)
    orc.finalDeallocate(ho)


def test_site_order_and_duplication_properties(packaged):
    """Textbook reduction (no cross-locus carry): per-site terms are a function of the site column alone, so
    permuting the columns permutes them and duplicating the alignment doubles the total."""
    from dmphyclus_b200.engine import Engine
    eng = Engine(quirk_carry=False)
    p = Problem(packaged, 2000, 3000, 10, 3, "iid", seed=77)
    al = eng.getConvertedAlignment(list("atcg"), p.chars)
    perm = np.random.default_rng(1).permutation(p.S)

    def run(a):
        r = eng.logLikCpp(p.edge, p.mrcas, p.pi, p.W[4], p.B[4], 1, a, p.n, a.shape[1], 5, 5)
        s = eng.site_logliks(r["solutionPointer"])
        eng.finalDeallocate(r["solutionPointer"])
        return r["logLik"], s

    l0, s0 = run(al)
    l1, s1 = run(np.ascontiguousarray(al[:, perm]))
    assert np.array_equal(s1, s0[perm])
    l2, s2 = run(np.concatenate([al, al], axis=1))
    assert np.array_equal(s2[:p.S], s0) and np.array_equal(s2[p.S:], s0)
    assert rel(l2, 2 * l0) < 1e-13 and rel(l1, l0) < 1e-13


# ---------------------------------------------------------------------------------------------- small-clade fusion
@pytest.mark.parametrize("fuse", [1, 2, 3, 8, 15])
@pytest.mark.parametrize("n,S,k,R,kind,shape", [(300, 140, 5, 3, "evolved", "random"), (700, 130, 8, 3, "iid", "random"),
                                                 (257, 70, 4, 1, "iid", "balanced"), (120, 60, 4, 4, "iid", "caterpillar"),
                                                 (9, 40, 2, 2, "evolved", "random")])
def test_fusion_does_not_change_results(packaged, oracle, fuse, n, S, k, R, kind, shape):
    """Vertices with <= fuse_tips tips below them are recomputed in registers instead of stored: every setting gives
    the same partials (bit for bit), exponents, topology and log-likelihoods over the whole scripted call sequence."""
    from dmphyclus_b200.engine import Engine
    if R > 3:
        packaged = dict(packaged)
        packaged["within"] = np.concatenate([packaged["within"]] * 2, axis=1)[:, :R]
        packaged["between"] = np.concatenate([packaged["between"], np.roll(packaged["between"], 1, axis=0)], axis=1)[:, :R]
    p = Problem(packaged, n, S, k, R, kind, shape, seed=n + fuse)
    got, want = run_script(Engine(fuse_tips=fuse), p), run_script(oracle, p)
    compare_to_golden(got, want, ll_tol=TOL, exact_partials=True)


@pytest.mark.parametrize("start_exact", [False, True])
def test_fused_clades_that_rescale(packaged, oracle, start_exact):
    """Matrices with 1e-80 off-diagonals make three-tip clades rescale.  The fast fused path must notice (violation
    flag), re-run the evaluation in exact mode and from then on keep per-vertex exponents for fused vertices."""
    from dmphyclus_b200.engine import Engine
    p = Problem(packaged, 400, 150, 6, 3, "tiny", seed=5)
    eng = Engine(fuse_tips=8, fuse_exact=start_exact)
    got, want = run_script(eng, p), run_script(oracle, p)
    compare_to_golden(got, want, ll_tol=TOL, exact_partials=True)
    hg, lg = p.create(eng)
    ho, lo = p.create(oracle)
    assert eng.stats(hg)["exact_mode"] == 1 and eng.stats(hg)["last_fused_vertices"] > 0
    assert rel(lg, lo) < TOL
    for v in range(p.n, 2 * p.n - 1, 37):                 # stored and fused vertices alike
        (sg, eg), (so, eo) = eng.partial(hg, v), oracle.partial(ho, v)
        assert np.array_equal(sg, so) and np.allclose(eg, eo, rtol=2e-7, atol=0), v
    eng.finalDeallocate(hg); oracle.finalDeallocate(ho)


def test_fused_vertices_are_readable(packaged, engine, oracle):
    """dmpc_get_partial of a vertex that is never stored replays its clade program."""
    p = Problem(packaged, 200, 90, 5, 3, "evolved", seed=3)
    hg, _ = p.create(engine)
    ho, _ = p.create(oracle)
    st = engine.stats(hg)
    assert st["last_fused_vertices"] > st["last_stored_vertices"] > 0
    assert st["last_fused_vertices"] + st["last_stored_vertices"] == p.n - 1
    for v in range(p.n, 2 * p.n - 1):
        assert np.array_equal(engine.partial(hg, v)[0], oracle.partial(ho, v)[0]), v
    engine.finalDeallocate(hg); oracle.finalDeallocate(ho)


# ---------------------------------------------------------------------------------------------- walked chains
@pytest.mark.parametrize("n,S,R,kind,fuse", [(3000, 300, 3, "iid", 15), (700, 520, 3, "evolved", 1), (1200, 257, 2, "diverged", 4)])
def test_caterpillar_spine_is_walked(packaged, oracle, n, S, R, kind, fuse):
    """A caterpillar's spine is one chain of single-task levels: k_walk evaluates it with the running partial in
    registers (many descriptor chunks and flag windows, several site tiles, rescaling at most steps on iid data).
    Partials bit-exact and exponents to the last bit at vertices all along the spine."""
    from dmphyclus_b200.engine import Engine
    eng = Engine(fuse_tips=fuse)
    p = Problem(packaged, n, S, 4, R, kind, "caterpillar", seed=n + fuse)
    hg, lg = p.create(eng)
    ho, lo = p.create(oracle)
    st = eng.stats(hg)
    assert st["last_walk_steps"] >= n - 1 - 16 and st["last_walk_steps"] + st["last_materialised"] + 4 >= st["last_stored_vertices"]
    assert rel(lg, lo) < TOL, (lg, lo)
    for v in sorted({n, n + 1, n + 31, n + 32, n + 255, n + 256, n + 257, n + (n - 1) // 2, 2 * n - 20, 2 * n - 2}):
        (sg, eg), (so, eo) = eng.partial(hg, v), oracle.partial(ho, v)
        assert np.array_equal(sg, so), f"partials of vertex {v} differ"
        assert np.allclose(eg, eo, rtol=2e-7, atol=0), f"exponents of vertex {v} differ"
    assert eng.recomputeAll(hg, p.W[4], p.B[4], p.pi) == lg          # cached plan + graph replay of the same chain
    a = eng.newBetweenTransProbsLogLik(hg, p.W[4], p.B[1], 1, 2, p.pi)
    b = oracle.newBetweenTransProbsLogLik(ho, p.W[4], p.B[1], 1, 2, p.pi)
    assert rel(a, b) < TOL
    eng.finalDeallocate(hg); oracle.finalDeallocate(ho)


@pytest.mark.parametrize("n,S,k,R,kind,fuse", [(1000, 300, 24, 3, "evolved", 15), (1000, 300, 24, 3, "iid", 8), (600, 260, 12, 1, "diverged", 15)])
def test_root_path_proposals_are_walked(packaged, oracle, n, S, k, R, kind, fuse):
    """Split / merge proposals dirty one root path: a chain.  The fused clades hanging off it are materialised (stored
    for the occasion), the path is walked; accepted or rejected, every partial along the path, the rollback and the
    next full evaluation agree with the oracle."""
    from dmphyclus_b200.engine import Engine
    from dmphyclus_b200.phylo import Phylo
    eng = Engine(fuse_tips=fuse)
    p = Problem(packaged, n, S, k, R, kind, seed=n + k + fuse)
    hg, l0 = p.create(eng)
    ho, _ = p.create(oracle)
    W, B = p.W[4], p.B[4]
    phy = Phylo(p.edge, p.n)
    walked = mats = 0
    props = _split_merge_candidates(p)
    assert len(props) >= 6
    for i, q in enumerate(props[:10]):
        ids = [q[1]] if q[0] == "split" else [q[1], q[2]]
        got = eng.clusSplitMergeLogLik(hg, W, B, ids, 1, p.pi)
        want = oracle.clusSplitMergeLogLik(ho, W, B, ids, 1, p.pi)
        assert rel(got, want) < TOL, (q, got, want)
        st = eng.stats(hg)
        walked += st["last_walk_steps"] > 0
        mats += st["last_materialised"]
        v = ids[0] if q[0] == "split" else int(phy.parent[ids[0]])
        while v:                                            # every vertex of the dirtied path, stored or fused
            (sg, eg), (so, eo) = eng.partial(hg, v), oracle.partial(ho, v)
            assert np.array_equal(sg, so) and np.allclose(eg, eo, rtol=2e-7, atol=0), (q, v)
            v = int(phy.parent[v])
        for be, h in ((eng, hg), (oracle, ho)):
            be.RestorePreviousConfig(h, np.zeros((1, 2), np.int32), 5, 5, False, p.mrcas, True)
        if i % 3 == 2:
            assert np.array_equal(eng.partial(hg, p.n)[0], oracle.partial(ho, p.n)[0])
    assert walked > 0 and (fuse == 1 or mats > 0)
    assert eng.recomputeAll(hg, W, B, p.pi) == l0
    eng.finalDeallocate(hg); oracle.finalDeallocate(ho)


# ---------------------------------------------------------------------------------------------- proposal batches
def _split_merge_candidates(p):
    """Every legal split (clusters whose MRCA is internal) and merge (sibling cluster MRCAs) of the current clustering."""
    from dmphyclus_b200.phylo import Phylo
    phy = Phylo(p.edge, p.n)
    props = [("split", m) for m in p.mrcas if m > p.n]
    by_parent = {}
    for m in p.mrcas:
        by_parent.setdefault(int(phy.parent[m]), []).append(m)
    props += [("merge", g[0], g[1]) for g in by_parent.values() if len(g) == 2]
    return props


@pytest.mark.parametrize("n,S,k,R,kind,fuse", [(300, 700, 24, 3, "evolved", 8), (300, 700, 24, 3, "evolved", 1), (64, 300, 9, 1, "evolved", 8),
                                                (900, 300, 30, 3, "iid", 8), (500, 520, 40, 2, "evolved", 15), (1024, 600, 25, 3, "diverged", 15),
                                                (700, 300, 20, 1, "iid", 8)])
def test_proposal_batch_equals_one_by_one(packaged, oracle, n, S, k, R, kind, fuse):
    """dmpc_eval_proposals: B proposals per launch == the same proposals evaluated one at a time through
    clusSplitMergeLogLik + RestorePreviousConfig (bitwise on the GPU, 1e-9 against the oracle); the committed state is
    untouched.  The iid case rescales everywhere: the batch then runs the full root reduction (exponent lists with the
    path vertices' new exponents spliced in, the sequential fp32 chain) per proposal."""
    from dmphyclus_b200.engine import Engine
    eng = Engine(fuse_tips=fuse)
    p = Problem(packaged, n, S, k, R, kind, seed=n + k)
    props = _split_merge_candidates(p)
    assert len(props) >= 5
    hg, l0 = p.create(eng)
    ho, _ = p.create(oracle)
    W, B = p.W[4], p.B[4]
    got, nb = eng.evalProposals(hg, W, B, p.pi, props)
    assert nb == len(props)
    one_by_one, want = [], []
    for q in props:
        ids = [q[1]] if q[0] == "split" else [q[1], q[2]]
        for be, h, out in ((eng, hg, one_by_one), (oracle, ho, want)):
            out.append(be.clusSplitMergeLogLik(h, W, B, ids, 1, p.pi))
            be.RestorePreviousConfig(h, np.zeros((1, 2), np.int32), 5, 5, False, p.mrcas, True)
    assert list(got) == one_by_one
    assert np.all(np.abs(got - np.array(want)) <= TOL * np.abs(np.array(want)))
    assert eng.recomputeAll(hg, W, B, p.pi) == l0                 # nothing moved
    got2, _ = eng.evalProposals(hg, W, B, p.pi, props[:3])
    assert list(got2) == one_by_one[:3]
    eng.finalDeallocate(hg); oracle.finalDeallocate(ho)


# ---------------------------------------------------------------------------------------------- the R-facing shim
@pytest.mark.parametrize("fast_masks", [False, True])
def test_rcpp_shim_end_to_end(packaged, oracle, fast_masks, monkeypatch):
    """rpkg/src/clusterFunArmadillo.cpp — the file a maintainer drops into the R package — compiled against the Rcpp
    stand-in (there is no R here) and driven through fake R objects: list-of-lists alignment, column-major matrices,
    1-based ids, external pointer, edge matrix returned as a numeric matrix.  Same scripted sequence as the oracle."""
    from oracle.oracle import Reference, build_shimtest
    path = build_shimtest()
    if path is None:
        pytest.skip("libshimtest.so not built")
    if fast_masks:
        monkeypatch.setenv("SHIM_FAST_MASKS", "1")     # the packed-mask attribute getConvertedAlignment attaches
    else:
        monkeypatch.delenv("SHIM_FAST_MASKS", raising=False)
    shim = Reference(path)
    for n, S, k, R, kind in ((40, 70, 5, 3, "evolved"), (500, 60, 6, 3, "iid")):
        p = Problem(packaged, n, S, k, R, kind, seed=n)
        compare_to_golden(run_script(shim, p), run_script(oracle, p), ll_tol=TOL, exact_partials=True)


def test_chain_walker_group_sizes_agree(packaged, oracle):
    """The sequential exponent chain pads a site's rows to groups of 4 (long lists) or walks them one by one (sparse
    lists); the engine switches after seeing the first evaluation.  Both must give the oracle's value.  (The chain
    certificate is switched off here: this is about the literal walker, k_chain.)"""
    from dmphyclus_b200.engine import Engine
    engine = Engine(chain_cert=False)
    p = Problem(packaged, 1024, 900, 6, 3, "diverged", seed=3)       # few rescales per site -> the sparse walker
    hg, l0 = p.create(engine)
    ho, lo = p.create(oracle)
    assert engine.stats(hg)["last_chain_sites"] > 0 and rel(l0, lo) < TOL
    g_after_first = engine.stats(hg)["chain_group"]
    l1 = engine.recomputeAll(hg, p.W[4], p.B[4], p.pi)
    assert l1 == l0
    a = engine.newBetweenTransProbsLogLik(hg, p.W[4], p.B[2], 1, 3, p.pi)
    b = oracle.newBetweenTransProbsLogLik(ho, p.W[4], p.B[2], 1, 3, p.pi)
    assert rel(a, b) < TOL
    q = Problem(packaged, 900, 300, 5, 3, "iid", seed=4)             # ~12 rescales per site and rate -> groups of 4
    hq, lq = q.create(engine)
    hq2, lq2 = q.create(oracle)
    assert rel(lq, lq2) < TOL and engine.recomputeAll(hq, q.W[4], q.B[4], q.pi) == lq
    assert {g_after_first, engine.stats(hq)["chain_group"]} == {1, 4}
    for be, h in ((engine, hg), (oracle, ho), (engine, hq), (oracle, hq2)):
        be.finalDeallocate(h)


def test_plan_cache_and_graph_replay(packaged, engine, oracle):
    """Repeated full evaluations of an unchanged tree replay the cached plan (as one CUDA graph launch); proposals,
    rollbacks and topology changes in between must invalidate or bypass it correctly."""
    p = Problem(packaged, 300, 600, 6, 3, "evolved", seed=12)
    hg, l0 = p.create(engine)
    ho, _ = p.create(oracle)
    W, B = p.W, p.B
    for _ in range(4):
        assert engine.recomputeAll(hg, W[4], B[4], p.pi) == l0
    st = engine.stats(hg)
    assert st["plan_cache_hits"] >= 3 and st["graph_replays"] >= 3
    # new within-cluster matrices = full recompute through the cached plan, accepted, then rejected
    for idx, keep in ((7, True), (2, False), (9, True)):
        a = engine.newWithinTransProbsLogLik(hg, W[idx - 1], B[4], 1, idx, p.pi)
        b = oracle.newWithinTransProbsLogLik(ho, W[idx - 1], B[4], 1, idx, p.pi)
        assert rel(a, b) < TOL
        if not keep:
            for be, h in ((engine, hg), (oracle, ho)):
                be.RestorePreviousConfig(h, np.zeros((1, 2), np.int32), 7, 5, False, p.mrcas, False)
    # a topology change (accepted NNI) must not be served from the stale plan
    m = [x for x in p.mrcas if x > p.n][0]
    ra = engine.withinClusNNIlogLik(hg, W[8], B[4], m, 2, 1, p.pi)
    rb = oracle.withinClusNNIlogLik(ho, W[8], B[4], m, 2, 1, p.pi)
    assert np.array_equal(ra["edge"], rb["edge"]) and rel(ra["logLik"], rb["logLik"]) < TOL
    a = engine.newWithinTransProbsLogLik(hg, W[0], B[4], 1, 1, p.pi)
    b = oracle.newWithinTransProbsLogLik(ho, W[0], B[4], 1, 1, p.pi)
    assert rel(a, b) < TOL
    assert np.array_equal(engine.partial(hg, p.n)[0], oracle.partial(ho, p.n)[0])
    engine.finalDeallocate(hg); oracle.finalDeallocate(ho)
