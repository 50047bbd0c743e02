"""ctypes binding of the parity oracle (`oracle/liborc.so`, built from `oracle/dmpc_oracle.cpp`).

TEST INFRASTRUCTURE ONLY: imported by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs — never by the product package `dmphyclus_b200`.

The class exposes the reference's eleven Rcpp entry points under their own names
(`/root/reference/src/clusterFunArmadillo.cpp:36-332`, `R/RcppExports.R:4-46`) on numpy arrays:
  * a transition-matrix "list" is an array [R][4][4] with m[r][i, j] = R's mat(i, j)
  * an edge matrix is an int array [E][2], 1-based, parents in column 0 (ape's `phylo$edge`)
  * `alignmentBin` is the packed form of getConvertedAlignment's result: uint8 [N][S], bit j set
    when state j of names(limProbs) is compatible with the cell.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dbl_p = C.POINTER(C.c_double)
_int_p = C.POINTER(C.c_int)
_u8_p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    """Compile liborc.so (and, where /root/reference exists, oracle/_ref/libref.so)."""
    lib = os.path.join(HERE, "liborc.so")
    src = os.path.join(HERE, "dmpc_oracle.cpp")
    if force or not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "liborc.so"], stdout=subprocess.DEVNULL)
    return lib


def _mats(m) -> np.ndarray:
    """[R][4][4] (row, col) -> flat column-major per matrix, as R hands them to C++."""
    m = np.asarray(m, dtype=np.float64)
    assert m.ndim == 3 and m.shape[1:] == (4, 4), m.shape
    return np.ascontiguousarray(m.transpose(0, 2, 1)).ravel()


def _edge(e) -> np.ndarray:
    e = np.asarray(e)
    assert e.ndim == 2 and e.shape[1] == 2
    return np.ascontiguousarray(e.T, dtype=np.int32).ravel()


def _p(a, t):
    return a.ctypes.data_as(t)


class Oracle:
    """One instance per loaded library; handles are opaque ints (the reference's XPtr<AugTree>)."""

    name = "oracle"

    def __init__(self, threads: int = 1, quirk_carry: bool = True):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_loglik_create.restype = C.c_void_p
        L.orc_loglik_create.argtypes = [_int_p, C.c_int, _dbl_p, C.c_int, _dbl_p, _dbl_p, _dbl_p, C.c_int,
                                        _u8_p, C.c_int, C.c_int, C.c_int, C.c_int, _dbl_p]
        for f in (L.orc_new_between, L.orc_new_within):
            f.restype = C.c_double
            f.argtypes = [C.c_void_p, _dbl_p, _dbl_p, C.c_int, _dbl_p]
        L.orc_within_nni.restype = C.c_double
        L.orc_within_nni.argtypes = [C.c_void_p, _dbl_p, _dbl_p, C.c_int, C.c_int, _dbl_p, _int_p]
        L.orc_between_nni.restype = C.c_double
        L.orc_between_nni.argtypes = [C.c_void_p, _dbl_p, _dbl_p, _dbl_p, C.c_int, C.c_int, _dbl_p, _int_p]
        L.orc_split_merge.restype = C.c_double
        L.orc_split_merge.argtypes = [C.c_void_p, _dbl_p, _dbl_p, _int_p, C.c_int, _dbl_p]
        L.orc_restore.restype = None
        L.orc_restore.argtypes = [C.c_void_p, _int_p, C.c_int, C.c_int, C.c_int, C.c_int, _int_p, C.c_int, C.c_int]
        L.orc_final_deallocate.argtypes = [C.c_void_p]
        L.orc_max_map_size_estimate.restype = C.c_double
        L.orc_max_map_size_estimate.argtypes = [C.c_double, C.c_uint]
        L.orc_convert_alignment.restype = C.c_long
        L.orc_convert_alignment.argtypes = [C.c_char_p, C.c_long, C.c_char_p, _u8_p]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_quirk_carry.argtypes = [C.c_int]
        L.orc_updates.restype = C.c_long
        L.orc_updates.argtypes = [C.c_void_p]
        L.orc_reset_updates.argtypes = [C.c_void_p]
        L.orc_get_partial.argtypes = [C.c_void_p, C.c_int, _dbl_p, C.POINTER(C.c_float)]
        L.orc_get_topology.argtypes = [C.c_void_p, _int_p, _int_p, _int_p, _int_p]
        L.orc_taus_stream.argtypes = [C.c_ulong, C.c_int, C.POINTER(C.c_ulong)]
        self.threads = threads
        self.quirk_carry = quirk_carry
        self._shape = {}

    # ------------------------------------------------------------------ the 11 entry points
    def getConvertedAlignment(self, equivVector, alignmentAlphaMat) -> np.ndarray:
        a = np.asarray(alignmentAlphaMat)
        if a.dtype != np.uint8:
            a = np.char.encode(a.astype(str), "ascii").view(np.uint8).reshape(a.shape)
        a = np.ascontiguousarray(a)
        out = np.empty(a.shape, np.uint8)
        states = "".join(equivVector).encode()
        assert len(states) == 4, "4-state DNA only"
        bad = self.lib.orc_convert_alignment(a.ctypes.data_as(C.c_char_p), a.size, states, _p(out, _u8_p))
        if bad:
            raise ValueError(f"{bad} alignment cells hold a symbol outside the DNA/IUPAC alphabet")
        return out

    def logLikCpp(self, edgeMat, clusterMRCAs, limProbsVec, withinTransMatList, betweenTransMatList, numOpenMP,
                  alignmentBin, numTips, numLoci, withinMatListIndex, betweenMatListIndex):
        e = _edge(edgeMat)
        m = np.asarray(clusterMRCAs, dtype=np.float64).ravel()
        pi = np.asarray(limProbsVec, dtype=np.float64)
        w, b = _mats(withinTransMatList), _mats(betweenTransMatList)
        al = np.ascontiguousarray(alignmentBin, dtype=np.uint8)
        assert al.shape == (numTips, numLoci)
        ll = C.c_double()
        self.lib.orc_set_threads(self.threads)
        self.lib.orc_set_quirk_carry(int(self.quirk_carry))
        h = self.lib.orc_loglik_create(_p(e, _int_p), len(e) // 2, _p(m, _dbl_p), len(m), _p(pi, _dbl_p),
                                       _p(w, _dbl_p), _p(b, _dbl_p), len(w) // 16, _p(al, _u8_p), numTips, numLoci,
                                       withinMatListIndex, betweenMatListIndex, C.byref(ll))
        if not h:
            raise ValueError("oracle: invalid tree (edge matrix out of range or not bifurcating)")
        self._shape[h] = (numTips, numLoci, len(w) // 16)
        return {"logLik": ll.value, "solutionPointer": h, "alignmentBinPointer": None}

    def newBetweenTransProbsLogLik(self, ptr, withinTransProbs, newBetweenTransProbs, numOpenMP,
                                   newBetweenMatListIndex, limProbs) -> float:
        w, b, pi = _mats(withinTransProbs), _mats(newBetweenTransProbs), np.asarray(limProbs, np.float64)
        return self.lib.orc_new_between(ptr, _p(w, _dbl_p), _p(b, _dbl_p), newBetweenMatListIndex, _p(pi, _dbl_p))

    def newWithinTransProbsLogLik(self, ptr, newWithinTransProbs, betweenTransProbs, numOpenMP,
                                  newWithinMatListIndex, limProbs) -> float:
        w, b, pi = _mats(newWithinTransProbs), _mats(betweenTransProbs), np.asarray(limProbs, np.float64)
        return self.lib.orc_new_within(ptr, _p(w, _dbl_p), _p(b, _dbl_p), newWithinMatListIndex, _p(pi, _dbl_p))

    def withinClusNNIlogLik(self, ptr, withinTransProbs, betweenTransProbs, MRCAofClusForNNI, numMovesNNI,
                            numOpenMP, limProbs):
        w, b, pi = _mats(withinTransProbs), _mats(betweenTransProbs), np.asarray(limProbs, np.float64)
        n = self._shape[ptr][0]
        edge = np.zeros(2 * (2 * n - 2), np.int32)
        ll = self.lib.orc_within_nni(ptr, _p(w, _dbl_p), _p(b, _dbl_p), int(MRCAofClusForNNI), int(numMovesNNI),
                                     _p(pi, _dbl_p), _p(edge, _int_p))
        if np.isnan(ll) and not edge.any():      # (a NaN log-likelihood WITH an edge matrix is a result: total underflow)
            raise ValueError("no NNI move is possible in this cluster")
        return {"logLik": ll, "edge": edge.reshape(2, -1).T.copy()}

    def betweenClusNNIlogLik(self, ptr, withinTransProbs, betweenTransProbs, clusterMRCAs, numMovesNNI,
                             numOpenMP, limProbs):
        w, b, pi = _mats(withinTransProbs), _mats(betweenTransProbs), np.asarray(limProbs, np.float64)
        m = np.asarray(clusterMRCAs, dtype=np.float64).ravel()
        n = self._shape[ptr][0]
        edge = np.zeros(2 * (2 * n - 2), np.int32)
        ll = self.lib.orc_between_nni(ptr, _p(w, _dbl_p), _p(b, _dbl_p), _p(m, _dbl_p), len(m), int(numMovesNNI),
                                      _p(pi, _dbl_p), _p(edge, _int_p))
        if np.isnan(ll) and not edge.any():
            raise ValueError("no NNI move is possible in the supporting tree")
        return {"logLik": ll, "edge": edge.reshape(2, -1).T.copy()}

    def clusSplitMergeLogLik(self, ptr, withinTransProbs, betweenTransProbs, clusMRCAsToSplitOrMerge, numOpenMP,
                             limProbs) -> float:
        w, b, pi = _mats(withinTransProbs), _mats(betweenTransProbs), np.asarray(limProbs, np.float64)
        m = np.ascontiguousarray(np.asarray(clusMRCAsToSplitOrMerge).ravel(), dtype=np.int32)
        return self.lib.orc_split_merge(ptr, _p(w, _dbl_p), _p(b, _dbl_p), _p(m, _int_p), len(m), _p(pi, _dbl_p))

    def RestorePreviousConfig(self, ptr, edgeMat, withinMatListIndex, betweenMatListIndex, NNImove, clusterMRCAs,
                              splitMergeMove) -> None:
        e = _edge(edgeMat)
        m = np.ascontiguousarray(np.asarray(clusterMRCAs).ravel(), dtype=np.int32)
        self.lib.orc_restore(ptr, _p(e, _int_p), len(e) // 2, int(withinMatListIndex), int(betweenMatListIndex),
                             int(bool(NNImove)), _p(m, _int_p), len(m), int(bool(splitMergeMove)))

    def finalDeallocate(self, ptr) -> None:
        self._shape.pop(ptr, None)
        self.lib.orc_final_deallocate(ptr)

    def getMaxMapSizeEstimate(self, allowedSizeInMegs, numRateCats) -> float:
        return self.lib.orc_max_map_size_estimate(float(allowedSizeInMegs), int(numRateCats))

    def checkAndCullMap(self, ptr, allowedNumberOfElements, cullProportion, withinTransProbs, betweenTransProbs,
                        limProbs) -> None:
        """The dense oracle has no dictionary; culling is a no-op (results do not depend on it)."""

    # ------------------------------------------------------------------ introspection (tests)
    def partial(self, ptr, node):
        n, s, r = self._shape[ptr]
        sol = np.empty((s, r, 4), np.float64)
        ex = np.empty((s, r), np.float32)
        self.lib.orc_get_partial(ptr, node, _p(sol, _dbl_p), ex.ctypes.data_as(C.POINTER(C.c_float)))
        return sol, ex

    def topology(self, ptr):
        n = 2 * self._shape[ptr][0] - 1
        arrs = [np.empty(n, np.int32) for _ in range(4)]
        self.lib.orc_get_topology(ptr, *[_p(a, _int_p) for a in arrs])
        return dict(zip(("parent", "child0", "child1", "within"), arrs))

    def updates(self, ptr) -> int:
        return self.lib.orc_updates(ptr)

    def reset_updates(self, ptr) -> None:
        self.lib.orc_reset_updates(ptr)

    def taus_stream(self, seed: int, n: int) -> np.ndarray:
        out = (C.c_ulong * n)()
        self.lib.orc_taus_stream(seed, n, out)
        return np.array(list(out), dtype=np.uint64)


def build_ref() -> str | None:
    """Compile oracle/_ref/libref.so from the reference's own sources where /root/reference exists;
    on the GPU box only the prebuilt file (shipped with the snapshot) can be used."""
    lib = os.path.join(HERE, "_ref", "libref.so")
    if os.path.isdir(os.environ.get("DMPC_REFERENCE_SRC", "/root/reference/src")):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    return lib if os.path.exists(lib) else None


def build_shimtest() -> str | None:
    """oracle/_ref/libshimtest.so: rpkg/src/clusterFunArmadillo.cpp (the drop-in Rcpp shim) compiled against the Rcpp
    stand-in and linked to libdmpc.so, behind the same harness as the reference TUs.  Needs libdmpc.so built."""
    lib = os.path.join(HERE, "_ref", "libshimtest.so")
    try:
        subprocess.check_call(["make", "-C", HERE, "shimtest"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        pass
    return lib if os.path.exists(lib) else None


class Reference:
    """The reference's own translation units (compiled against oracle/shim) behind the same
    interface as `Oracle`.  Single-threaded, memoised through its `std::map` dictionary."""

    name = "reference"

    def __init__(self, path: str | None = None):
        if path is None:
            path = build_ref()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libref.so is absent and /root/reference is not available to build it")
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_last_error.restype = C.c_char_p
        L.ref_loglik_create.restype = C.c_void_p
        L.ref_loglik_create.argtypes = [_int_p, C.c_int, _dbl_p, C.c_int, _dbl_p, _dbl_p, _dbl_p, C.c_int,
                                        _u8_p, C.c_int, C.c_int, C.c_int, C.c_int, _dbl_p]
        for f in (L.ref_new_between, L.ref_new_within):
            f.restype = C.c_double
            f.argtypes = [C.c_void_p, _dbl_p, _dbl_p, C.c_int, C.c_int, _dbl_p]
        L.ref_within_nni.restype = C.c_double
        L.ref_within_nni.argtypes = [C.c_void_p, _dbl_p, _dbl_p, C.c_int, C.c_int, C.c_int, _dbl_p, _int_p]
        L.ref_between_nni.restype = C.c_double
        L.ref_between_nni.argtypes = [C.c_void_p, _dbl_p, _dbl_p, C.c_int, _dbl_p, C.c_int, C.c_int, _dbl_p, _int_p]
        L.ref_split_merge.restype = C.c_double
        L.ref_split_merge.argtypes = [C.c_void_p, _dbl_p, _dbl_p, C.c_int, _int_p, C.c_int, _dbl_p]
        L.ref_restore.restype = None
        L.ref_restore.argtypes = [C.c_void_p, _int_p, C.c_int, C.c_int, C.c_int, C.c_int, _int_p, C.c_int, C.c_int]
        L.ref_final_deallocate.argtypes = [C.c_void_p]
        L.ref_max_map_size_estimate.restype = C.c_double
        L.ref_max_map_size_estimate.argtypes = [C.c_double, C.c_uint]
        L.ref_check_and_cull_map.argtypes = [C.c_void_p, C.c_uint, C.c_float, _dbl_p, _dbl_p, C.c_int, _dbl_p]
        L.ref_convert_alignment.restype = C.c_long
        L.ref_convert_alignment.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, _u8_p]
        L.ref_get_partial.argtypes = [C.c_void_p, C.c_int, _dbl_p, C.POINTER(C.c_float)]
        L.ref_get_topology.argtypes = [C.c_void_p, _int_p, _int_p, _int_p, _int_p]
        L.ref_dictionary_size.restype = C.c_long
        L.ref_dictionary_size.argtypes = [C.c_void_p]
        self._shape = {}

    def getConvertedAlignment(self, equivVector, alignmentAlphaMat, allow_unknown=False) -> np.ndarray:
        a = np.asarray(alignmentAlphaMat)
        if a.dtype != np.uint8:
            a = np.char.encode(a.astype(str), "ascii").view(np.uint8).reshape(a.shape)
        a = np.ascontiguousarray(a)
        out = np.empty(a.shape, np.uint8)
        bad = self.lib.ref_convert_alignment(a.ctypes.data_as(C.c_char_p), a.shape[0], a.shape[1],
                                             "".join(equivVector).encode(), _p(out, _u8_p))
        if bad and not allow_unknown:
            raise ValueError(f"{bad} alignment cells hold a symbol outside the DNA/IUPAC alphabet")
        return out

    def logLikCpp(self, edgeMat, clusterMRCAs, limProbsVec, withinTransMatList, betweenTransMatList, numOpenMP,
                  alignmentBin, numTips, numLoci, withinMatListIndex, betweenMatListIndex):
        e = _edge(edgeMat)
        m = np.asarray(clusterMRCAs, dtype=np.float64).ravel()
        pi = np.asarray(limProbsVec, dtype=np.float64)
        w, b = _mats(withinTransMatList), _mats(betweenTransMatList)
        al = np.ascontiguousarray(alignmentBin, dtype=np.uint8)
        assert al.shape == (numTips, numLoci)
        ll = C.c_double()
        h = self.lib.ref_loglik_create(_p(e, _int_p), len(e) // 2, _p(m, _dbl_p), len(m), _p(pi, _dbl_p),
                                       _p(w, _dbl_p), _p(b, _dbl_p), len(w) // 16, _p(al, _u8_p), numTips, numLoci,
                                       withinMatListIndex, betweenMatListIndex, C.byref(ll))
        if not h:
            raise ValueError("reference: " + self.lib.ref_last_error().decode())
        self._shape[h] = (numTips, numLoci, len(w) // 16)
        return {"logLik": ll.value, "solutionPointer": h, "alignmentBinPointer": None}

    def newBetweenTransProbsLogLik(self, ptr, withinTransProbs, newBetweenTransProbs, numOpenMP,
                                   newBetweenMatListIndex, limProbs) -> float:
        w, b, pi = _mats(withinTransProbs), _mats(newBetweenTransProbs), np.asarray(limProbs, np.float64)
        return self.lib.ref_new_between(ptr, _p(w, _dbl_p), _p(b, _dbl_p), len(w) // 16, newBetweenMatListIndex, _p(pi, _dbl_p))

    def newWithinTransProbsLogLik(self, ptr, newWithinTransProbs, betweenTransProbs, numOpenMP,
                                  newWithinMatListIndex, limProbs) -> float:
        w, b, pi = _mats(newWithinTransProbs), _mats(betweenTransProbs), np.asarray(limProbs, np.float64)
        return self.lib.ref_new_within(ptr, _p(w, _dbl_p), _p(b, _dbl_p), len(w) // 16, newWithinMatListIndex, _p(pi, _dbl_p))

    def withinClusNNIlogLik(self, ptr, withinTransProbs, betweenTransProbs, MRCAofClusForNNI, numMovesNNI,
                            numOpenMP, limProbs):
        w, b, pi = _mats(withinTransProbs), _mats(betweenTransProbs), np.asarray(limProbs, np.float64)
        n = self._shape[ptr][0]
        edge = np.zeros(2 * (2 * n - 2), np.int32)
        ll = self.lib.ref_within_nni(ptr, _p(w, _dbl_p), _p(b, _dbl_p), len(w) // 16, int(MRCAofClusForNNI),
                                     int(numMovesNNI), _p(pi, _dbl_p), _p(edge, _int_p))
        return {"logLik": ll, "edge": edge.reshape(2, -1).T.copy()}

    def betweenClusNNIlogLik(self, ptr, withinTransProbs, betweenTransProbs, clusterMRCAs, numMovesNNI,
                             numOpenMP, limProbs):
        w, b, pi = _mats(withinTransProbs), _mats(betweenTransProbs), np.asarray(limProbs, np.float64)
        m = np.asarray(clusterMRCAs, dtype=np.float64).ravel()
        n = self._shape[ptr][0]
        edge = np.zeros(2 * (2 * n - 2), np.int32)
        ll = self.lib.ref_between_nni(ptr, _p(w, _dbl_p), _p(b, _dbl_p), len(w) // 16, _p(m, _dbl_p), len(m),
                                      int(numMovesNNI), _p(pi, _dbl_p), _p(edge, _int_p))
        return {"logLik": ll, "edge": edge.reshape(2, -1).T.copy()}

    def clusSplitMergeLogLik(self, ptr, withinTransProbs, betweenTransProbs, clusMRCAsToSplitOrMerge, numOpenMP,
                             limProbs) -> float:
        w, b, pi = _mats(withinTransProbs), _mats(betweenTransProbs), np.asarray(limProbs, np.float64)
        m = np.ascontiguousarray(np.asarray(clusMRCAsToSplitOrMerge).ravel(), dtype=np.int32)
        return self.lib.ref_split_merge(ptr, _p(w, _dbl_p), _p(b, _dbl_p), len(w) // 16, _p(m, _int_p), len(m), _p(pi, _dbl_p))

    def RestorePreviousConfig(self, ptr, edgeMat, withinMatListIndex, betweenMatListIndex, NNImove, clusterMRCAs,
                              splitMergeMove) -> None:
        e = _edge(edgeMat)
        m = np.ascontiguousarray(np.asarray(clusterMRCAs).ravel(), dtype=np.int32)
        self.lib.ref_restore(ptr, _p(e, _int_p), len(e) // 2, int(withinMatListIndex), int(betweenMatListIndex),
                             int(bool(NNImove)), _p(m, _int_p), len(m), int(bool(splitMergeMove)))

    def finalDeallocate(self, ptr) -> None:
        self._shape.pop(ptr, None)
        self.lib.ref_final_deallocate(ptr)

    def getMaxMapSizeEstimate(self, allowedSizeInMegs, numRateCats) -> float:
        return self.lib.ref_max_map_size_estimate(float(allowedSizeInMegs), int(numRateCats))

    def checkAndCullMap(self, ptr, allowedNumberOfElements, cullProportion, withinTransProbs, betweenTransProbs,
                        limProbs) -> None:
        w, b, pi = _mats(withinTransProbs), _mats(betweenTransProbs), np.asarray(limProbs, np.float64)
        self.lib.ref_check_and_cull_map(ptr, int(allowedNumberOfElements), float(cullProportion), _p(w, _dbl_p),
                                        _p(b, _dbl_p), len(w) // 16, _p(pi, _dbl_p))

    def partial(self, ptr, node):
        n, s, r = self._shape[ptr]
        sol = np.empty((s, r, 4), np.float64)
        ex = np.empty((s, r), np.float32)
        self.lib.ref_get_partial(ptr, node, _p(sol, _dbl_p), ex.ctypes.data_as(C.POINTER(C.c_float)))
        return sol, ex

    def topology(self, ptr):
        n = 2 * self._shape[ptr][0] - 1
        arrs = [np.empty(n, np.int32) for _ in range(4)]
        self.lib.ref_get_topology(ptr, *[_p(a, _int_p) for a in arrs])
        return dict(zip(("parent", "child0", "child1", "within"), arrs))

    def dictionary_size(self, ptr) -> int:
        return self.lib.ref_dictionary_size(ptr)
