"""Pins the CPU oracle against the reference's own translation units run HERE (oracle/_ref/libref.so, built from
/root/reference/src against oracle/shim).  Skipped where neither the prebuilt library nor /root/reference exists."""
import numpy as np
import pytest

from dmphyclus_b200 import chain
from dmphyclus_b200.workloads import Problem
from script import run_script


@pytest.fixture(scope="module")
def reference():
    from oracle.oracle import Reference
    try:
        return Reference()
    except (FileNotFoundError, OSError) as e:
        pytest.skip(f"reference TUs not available: {e}")


@pytest.mark.parametrize("n,S,k,R,kind,shape,seed", [
    (2, 7, 1, 3, "iid", "random", 1), (3, 5, 2, 1, "iid", "random", 2), (17, 33, 4, 3, "evolved", "random", 3),
    (128, 20, 7, 3, "iid", "balanced", 4), (400, 12, 5, 3, "iid", "caterpillar", 5), (900, 16, 9, 3, "iid", "random", 6),
    (1100, 12, 6, 2, "diverged", "random", 7)])
def test_script_bit_identical(oracle, reference, packaged, n, S, k, R, kind, shape, seed):
    p = Problem(packaged, n, S, k, R, kind, shape, seed=seed)
    a, b = run_script(oracle, p), run_script(reference, p)
    assert a.keys() == b.keys()
    for key in a:
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("n,S,k,R,kind,iters,seed", [(6, 200, 2, 3, "evolved", 30, 1), (60, 40, 5, 3, "evolved", 15, 2),
                                                     (700, 12, 8, 3, "iid", 6, 3), (40, 30, 3, 1, "iid", 12, 4),
                                                     (90, 25, 6, 2, "tiny", 10, 5), (250, 10, 4, 3, "diverged", 8, 6),
                                                     (33, 64, 2, 2, "evolved", 20, 7)])
def test_mcmc_trajectory_bit_identical(oracle, reference, packaged, n, S, k, R, kind, iters, seed):
    """The restated driver over both backends with one numpy stream: every sample identical, including the NNI picks
    drawn from the C++-side gsl_rng_taus stream."""
    p = Problem(packaged, n, S, k, R, kind, seed=seed)
    st = chain.ChainSettings(limProbs=p.pi, withinTransMatAll=p.W, betweenTransMatAll=p.B, shapeForAlpha=1000,
                             scaleForAlpha=0.1, alphaMin=10, poisRateNumClus=2, numSplitMergeMoves=3)
    outs = []
    for be in (oracle, reference):
        al = be.getConvertedAlignment(list("atcg"), p.chars)
        outs.append(chain.run_chain(be, al, p.edge, p.clus, 100.0, st, iters, seed=seed + 50)[0])
    for a, b in zip(*outs):
        assert np.array_equal(a["clusInd"], b["clusInd"]) and np.array_equal(a["edge"], b["edge"])
        assert a["logLik"] == b["logLik"] and a["logPostProb"] == b["logPostProb"]


def test_alphabet_matches_reference(oracle, reference):
    """getConvertedAlignment over every byte the reference's table knows, plus unknown symbols (empty uvec)."""
    chars = np.frombuffer(b"atcgryswkmbdhv-n.xz?AN", np.uint8).reshape(1, -1)
    ref_masks = reference.getConvertedAlignment(list("atcg"), chars, allow_unknown=True)
    import ctypes as C
    out = np.empty(chars.shape, np.uint8)
    bad = oracle.lib.orc_convert_alignment(chars.ctypes.data_as(C.c_char_p), chars.size, b"atcg",
                                           out.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert np.array_equal(out, ref_masks) and bad == int((ref_masks == 0).sum()) == 5
    # permuted state order (names(limProbs) decides the bit positions)
    assert np.array_equal(reference.getConvertedAlignment(list("gcta"), chars[:, :16]),
                          oracle.getConvertedAlignment(list("gcta"), chars[:, :16]))


def test_dictionary_is_a_pure_cache(reference, oracle, packaged):
    """checkAndCullMap (clusterFunArmadillo.cpp:305-332) never changes results: the dense layouts may ignore it."""
    p = Problem(packaged, 50, 40, 4, 3, "evolved", seed=9)
    h, l0 = p.create(reference)
    assert reference.dictionary_size(h) > 0
    reference.checkAndCullMap(h, 10, 0.5, p.W[4], p.B[4], p.pi)
    l1 = reference.newWithinTransProbsLogLik(h, p.W[4], p.B[4], 1, 5, p.pi)
    assert l1 == l0
    reference.finalDeallocate(h)
    assert oracle.getMaxMapSizeEstimate(100.0, 3) == reference.getMaxMapSizeEstimate(100.0, 3)
