"""Edge matrices that are not a rooted, strictly bifurcating tree on vertices 1..2N-1 with root N+1 are refused by the
host-side tree builder (HostTree::build — the checks the reference leaves to R's ape and to a failed assertion,
/root/reference/src/AugTree.cpp:59-86, IntermediateNode.cpp:59).  Device-free: through dmpc_hp_create."""
import ctypes as C

import numpy as np
import pytest

from dmphyclus_b200 import synth
from dmphyclus_b200.engine import load_library


def _create(lib, edge, n):
    e = np.ascontiguousarray(np.asarray(edge, np.int32).T).ravel()
    h = C.c_void_p()
    rc = lib.dmpc_hp_create(e.ctypes.data_as(C.POINTER(C.c_int32)), len(e) // 2, n, None, 0, 15, C.byref(h))
    if rc == 0:
        lib.dmpc_hp_destroy(h)
    return rc, lib.dmpc_last_error().decode()


@pytest.fixture(scope="module")
def lib():
    lib = load_library()
    lib.dmpc_hp_create.restype = C.c_int
    lib.dmpc_hp_create.argtypes = [C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int,
                                   C.POINTER(C.c_void_p)]
    lib.dmpc_hp_destroy.argtypes = [C.c_void_p]
    lib.dmpc_last_error.restype = C.c_char_p
    return lib


@pytest.mark.parametrize("shape", ["random", "balanced", "caterpillar"])
def test_well_formed_trees_are_accepted_in_any_row_order(lib, shape):
    rng = np.random.default_rng(11)
    n = 40
    edge = np.asarray(synth.random_tree(n, rng, shape))
    assert _create(lib, edge, n)[0] == 0
    assert _create(lib, edge[rng.permutation(len(edge))], n)[0] == 0          # rows shuffled: the traversal path of the check
    assert _create(lib, edge[::-1], n)[0] == 0


def test_malformed_trees_are_refused(lib):
    rng = np.random.default_rng(12)
    n = 12
    good = np.asarray(synth.random_tree(n, rng, "random"))
    assert _create(lib, good, n)[0] == 0

    def refused(edge, what, rows_of=n):
        rc, msg = _create(lib, edge, rows_of)
        assert rc != 0, what
        return msg

    assert "rows" in refused(good[:-1], "a row short")
    e = good.copy(); e[3, 1] = 2 * n + 5
    assert "out of range" in refused(e, "child id past the last vertex")
    e = good.copy(); e[3, 0] = 1
    assert "out of range" in refused(e, "a tip as a parent")
    # one internal vertex with three children, another with one
    e = good.copy()
    parents = np.unique(e[:, 0])
    a, b = parents[0], parents[1]
    e[np.flatnonzero(e[:, 0] == b)[0], 0] = a
    assert "bifurcating" in refused(e, "three children / one child")
    # the root named as somebody's child
    e = good.copy()
    row = np.flatnonzero((e[:, 1] > n) & (e[:, 1] != n + 1))[0]
    e[row, 1] = n + 1
    assert refused(e, "the root is a child")
    # a child named twice (so another vertex hangs nowhere): in pre-order and shuffled
    e = good.copy()
    tips_rows = np.flatnonzero(e[:, 1] <= n)
    e[tips_rows[1], 1] = e[tips_rows[0], 1]
    for variant in (e, e[rng.permutation(len(e))]):
        assert "single rooted tree" in refused(variant, "duplicate child")
    # two internal vertices that are each other's children (a cycle cut off from the root)
    e = good.copy()
    inner = [v for v in np.unique(e[:, 0]) if v != n + 1]
    u, w = inner[0], inner[1]
    ru, rw = np.flatnonzero(e[:, 1] == u)[0], np.flatnonzero(e[:, 1] == w)[0]
    e[ru, 0], e[rw, 0] = w, u                      # u's parent row now says w, w's says u
    keep_two = lambda v: np.flatnonzero(e[:, 0] == v)
    if len(keep_two(u)) == 2 and len(keep_two(w)) == 2:
        pass                                       # degrees intact: the connectivity check has to catch it
    assert refused(e, "cycle")


def test_restore_refuses_a_malformed_matrix(lib):
    """dmpc_restore_previous_config re-links the tree from the matrix R hands back (BuildTreeNoAssign,
    /root/reference/src/AugTree.cpp:88-101): the same checks apply there."""
    lib.dmpc_hp_restore.restype = C.c_int
    lib.dmpc_hp_restore.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int]
    rng = np.random.default_rng(13)
    n = 16
    good = np.asarray(synth.random_tree(n, rng, "random"))
    e = np.ascontiguousarray(good.astype(np.int32).T).ravel()
    h = C.c_void_p()
    assert lib.dmpc_hp_create(e.ctypes.data_as(C.POINTER(C.c_int32)), len(e) // 2, n, None, 0, 15, C.byref(h)) == 0
    other = np.asarray(synth.random_tree(n, rng, "random"))                      # a different, well-formed tree: accepted
    o = np.ascontiguousarray(other.astype(np.int32).T).ravel()
    assert lib.dmpc_hp_restore(h, o.ctypes.data_as(C.POINTER(C.c_int32)), len(o) // 2, 1, None, 0, 0) == 0
    bad = good.copy()
    tips_rows = np.flatnonzero(bad[:, 1] <= n)
    bad[tips_rows[1], 1] = bad[tips_rows[0], 1]                                  # a tip named twice
    b = np.ascontiguousarray(bad.astype(np.int32).T).ravel()
    assert lib.dmpc_hp_restore(h, b.ctypes.data_as(C.POINTER(C.c_int32)), len(b) // 2, 1, None, 0, 0) != 0
    assert b"single rooted tree" in lib.dmpc_last_error()
    assert lib.dmpc_hp_restore(h, e.ctypes.data_as(C.POINTER(C.c_int32)), len(e) // 2, 1, None, 0, 0) == 0   # and recovers
    lib.dmpc_hp_destroy(h)
