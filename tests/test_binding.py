"""The ctypes binding (dmphyclus_b200/engine.py) without a device: every entry point marshals its arguments the way
`include/dmpc.h` declares them — checked by calling through to libdmpc with a NULL handle, which each routine refuses
with its own status ("pointer is null.", /root/reference/src/clusterFunArmadillo.cpp:136) instead of ctypes refusing the
argument list — and the conversion caches hand back what a fresh conversion would."""
import ctypes as C

import numpy as np
import pytest

from dmphyclus_b200 import synth
from dmphyclus_b200.engine import DmpcError, Engine


@pytest.fixture(scope="module")
def eng():
    return Engine()


@pytest.fixture(scope="module")
def mats(packaged):
    w = np.array(packaged["within"][4, :3])
    b = np.array(packaged["between"][4, :3])
    pi = np.array(synth.LIM_PROBS)
    for a in (w, b, pi):
        a.setflags(write=False)
    return w, b, pi


def _refused(call):
    with pytest.raises(DmpcError) as e:
        call()
    assert "null" in str(e.value)


@pytest.mark.parametrize("n_tips", [6, 1000, 20000])      # both kinds of edge output buffer (ctypes array / numpy memory)
def test_every_entry_point_marshals_its_arguments(eng, mats, n_tips):
    w, b, pi = mats
    eng._shape[None] = (n_tips, 10, 3)
    mrcas = [n_tips + 2, n_tips + 3, 4]
    edge = np.zeros((2 * n_tips - 2, 2), np.int32)
    try:
        _refused(lambda: eng.newBetweenTransProbsLogLik(None, w, b, 1, 3, pi))
        _refused(lambda: eng.newWithinTransProbsLogLik(None, w, b, 1, 3, pi))
        _refused(lambda: eng.withinClusNNIlogLik(None, w, b, n_tips + 2, 1, 1, pi))
        _refused(lambda: eng.betweenClusNNIlogLik(None, w, b, mrcas, 1, 1, pi))
        _refused(lambda: eng.betweenClusNNIlogLik(None, w, b, np.asarray(mrcas, np.float64), 1, 1, pi))
        _refused(lambda: eng.clusSplitMergeLogLik(None, w, b, [n_tips + 2], 1, pi))
        _refused(lambda: eng.clusSplitMergeLogLik(None, w, b, np.array([3, 4]), 1, pi))
        _refused(lambda: eng.RestorePreviousConfig(None, edge, 5, 5, True, mrcas, False))
        _refused(lambda: eng.RestorePreviousConfig(None, np.zeros((1, 2), np.int32), 5, 5, False, np.asarray(mrcas), True))
        _refused(lambda: eng.RestorePreviousConfig(None, edge, 5, 5, False, [], True))
        _refused(lambda: eng.checkAndCullMap(None, 1000, 0.5, w, b, pi))
        _refused(lambda: eng.recomputeAll(None, w, b, pi))
        _refused(lambda: eng.evalProposals(None, w, b, pi, [("split", n_tips + 2), ("merge", 3, 4)]))
        _refused(lambda: eng.stats(None))
        _refused(lambda: eng.root_timeline(None))
        _refused(lambda: eng.timer_begin(None))
        _refused(lambda: eng.timer_end(None))
        _refused(lambda: eng.partial(None, n_tips + 1))
        _refused(lambda: eng.topology(None))
        _refused(lambda: eng.site_logliks(None))
        _refused(lambda: eng.finalDeallocate(None))
    finally:
        eng._shape.pop(None, None)
    assert eng.getMaxMapSizeEstimate(100.0, 3) == np.floor(100e6 / (32.0 + 24.0 * 3 + 16.0))


def test_id_list_cache_follows_the_list(eng):
    m = [11, 12, 13, 14]
    p1, n1 = eng._lp(m, np.int32, C.POINTER(C.c_int32))
    p2, n2 = eng._lp(m, np.int32, C.POINTER(C.c_int32))
    assert n1 == n2 == 4 and [p1[i] for i in range(4)] == m
    assert C.addressof(p1.contents) == C.addressof(p2.contents)             # converted once
    m[2] = 99                                                                # edited in place: must not be served stale
    p3, n3 = eng._lp(m, np.int32, C.POINTER(C.c_int32))
    assert n3 == 4 and [p3[i] for i in range(4)] == [11, 12, 99, 14]
    m.append(7)
    p4, n4 = eng._lp(m, np.int32, C.POINTER(C.c_int32))
    assert n4 == 5 and p4[4] == 7
    # the same list as doubles (betweenClusNNIlogLik takes a NumericVector) is a separate entry
    pd, nd = eng._lp(m, np.float64, C.POINTER(C.c_double))
    assert nd == 5 and [pd[i] for i in range(5)] == [11.0, 12.0, 99.0, 14.0, 7.0]
    # arrays and tuples take the uncached path
    pa, na = eng._lp(np.array([5.0, 6.0]), np.int32, C.POINTER(C.c_int32))
    assert na == 2 and (pa[0], pa[1]) == (5, 6)
    pe, ne = eng._lp([], np.int32, C.POINTER(C.c_int32))
    assert ne == 0
    for k in range(200):                                                     # the cache stays bounded
        eng._lp([k, k + 1], np.int32, C.POINTER(C.c_int32))
    assert len(eng._lcache) <= 64


@pytest.mark.parametrize("n_tips", [5, 4097, 4098, 30000])
def test_edge_output_buffer_is_a_column_major_view(eng, n_tips):
    """What the library writes through the pointer (column-major: parents, then children) is what the [E][2] view shows,
    and the view keeps its memory alive on its own."""
    buf, edge = Engine._edge_out(n_tips)
    rows = 2 * n_tips - 2
    assert edge.shape == (rows, 2) and edge.dtype == np.int32
    raw = C.cast(buf, C.POINTER(C.c_int32))
    want = np.arange(2 * rows, dtype=np.int32) * 3 + 1
    C.memmove(raw, want.ctypes.data, want.nbytes)
    del buf, raw
    assert np.array_equal(edge[:, 0], want[:rows]) and np.array_equal(edge[:, 1], want[rows:])
    e2 = np.array(edge)                                                      # what chain._frozen_edge does with it
    assert e2.flags.owndata and np.array_equal(e2, edge)


def test_edge_output_buffer_through_the_library(eng):
    """End to end without a device: dmpc_hp_state exports the edge matrix of a host-side plan handle (BuildEdgeMatrix
    order) through the same kind of buffer the NNI entry points use."""
    rng = np.random.default_rng(5)
    n = 300
    edge = np.asarray(synth.random_tree(n, rng, "random"), np.int32)
    e = np.ascontiguousarray(edge.T).ravel()
    h = C.c_void_p()
    lib = eng.lib
    assert lib.dmpc_hp_create(e.ctypes.data_as(C.POINTER(C.c_int32)), len(e) // 2, n, None, 0, 15, C.byref(h)) == 0
    try:
        buf, view = Engine._edge_out(n)
        cur = np.zeros(2 * n - 1, np.uint8)
        within = np.zeros(2 * n - 1, np.uint8)
        assert lib.dmpc_hp_state(h, cur.ctypes.data_as(C.POINTER(C.c_uint8)), within.ctypes.data_as(C.POINTER(C.c_uint8)), buf) == 0
        assert np.array_equal(view, edge)                                    # synth emits pre-order rows: the export is the input
    finally:
        lib.dmpc_hp_destroy(h)


def test_result_objects_of_the_nni_entry_points(mats, monkeypatch):
    """The lines after the library call (building the R-style result list) — reached here by ignoring the status of a
    NULL-handle call: `edge` is the [2N-2][2] view of the buffer that was handed to the library."""
    w, b, pi = mats
    eng = Engine()
    monkeypatch.setattr(eng, "_check", lambda rc: None)
    for n_tips in (50, 9000):
        eng._shape[None] = (n_tips, 10, 3)
        for res in (eng.withinClusNNIlogLik(None, w, b, n_tips + 2, 1, 1, pi),
                    eng.betweenClusNNIlogLik(None, w, b, [n_tips + 2, n_tips + 3], 1, 1, pi)):
            assert set(res) == {"logLik", "edge"} and isinstance(res["logLik"], float)
            assert res["edge"].shape == (2 * n_tips - 2, 2) and res["edge"].dtype == np.int32
            np.array(res["edge"], dtype=np.int32)            # what the driver does with it (chain._frozen_edge)
    assert eng.RestorePreviousConfig(None, np.zeros((1, 2), np.int32), 5, 5, False, [3, 4], True) is None
    assert isinstance(eng.clusSplitMergeLogLik(None, w, b, [7], 1, pi), float)


def test_loglikcpp_marshalling_with_cached_and_fresh_inputs(mats, packaged):
    """logLikCpp converts read-only inputs once (a logLikFromClusInd-style loop hands the same objects in again) and
    writeable ones every time; either way the call reaches the library, which — without a GPU — refuses with
    DMPC_ERR_CUDA, and malformed inputs are refused before that."""
    import torch
    from dmphyclus_b200.workloads import Problem
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the refusal path is exercised on CPU-only boxes")
    w, b, pi = mats
    eng = Engine()
    p = Problem(packaged, 12, 9, 3, 3, "iid", seed=4)
    al = eng.getConvertedAlignment(list(synth.STATES), p.chars)
    edge_ro = np.array(p.edge, np.int32)
    edge_ro.setflags(write=False)
    for edge in (edge_ro, edge_ro, np.array(p.edge), p.edge.tolist() if hasattr(p.edge, "tolist") else p.edge):
        for mr in (p.mrcas, list(p.mrcas), np.asarray(p.mrcas, np.float64)):
            with pytest.raises(DmpcError) as e:
                eng.logLikCpp(edge, mr, pi, w, b, 1, al, p.n, p.S, 5, 5)
            assert e.value.code == -2
    assert len(eng._ecache) == 1                                   # the read-only edge matrix was converted once
    with pytest.raises(ValueError):
        eng.logLikCpp(edge_ro, p.mrcas, pi, w, b[:2], 1, al, p.n, p.S, 5, 5)          # lists of different length
    with pytest.raises(ValueError):
        eng.logLikCpp(edge_ro, p.mrcas, pi[:3], w, b, 1, al, p.n, p.S, 5, 5)          # limProbs
    with pytest.raises(ValueError):
        eng.logLikCpp(edge_ro, p.mrcas, pi, w, b, 1, al[:, :5], p.n, p.S, 5, 5)       # alignment shape
    with pytest.raises(ValueError):
        eng.logLikCpp(edge_ro, p.mrcas, pi, w[:, :3], b, 1, al, p.n, p.S, 5, 5)       # not 4x4 matrices
