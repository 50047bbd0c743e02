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
