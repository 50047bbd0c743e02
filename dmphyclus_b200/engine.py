"""Host binding of the CUDA engine (`libdmpc.so`, C-ABI in `include/dmpc.h`) under the reference's own
entry-point names (`/root/reference/R/RcppExports.R:4-46`, `src/clusterFunArmadillo.cpp:36-332`).

This is the product path: it loads the hand-written sm_100a library and nothing else — no oracle, no CPU
fallback.  A missing library or a machine without a B200-class GPU raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DMPC_LIB_PATH", os.path.join(HERE, "libdmpc.so"))   # the override is for A/B builds of the kernels
_dbl_p = C.POINTER(C.c_double)
_i32_p = C.POINTER(C.c_int32)
_u8_p = C.POINTER(C.c_uint8)
_f32_p = C.POINTER(C.c_float)


class DmpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdmpc status {code}: {msg}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("quirk_carry", C.c_int), ("site_begin", C.c_int), ("sites_total", C.c_int),
                ("fuse_tips", C.c_int), ("fuse_exact", C.c_int), ("comm", C.c_void_p), ("deferred", C.c_int),
                ("chain_cert", C.c_int), ("n_devices", C.c_int), ("devices", C.c_int * 8)]


class Stats(C.Structure):
    _fields_ = [("updates", C.c_ulonglong), ("kernel_launches", C.c_ulonglong), ("evaluations", C.c_ulonglong),
                ("last_eval_ms", C.c_double), ("last_prune_ms", C.c_double), ("last_updates", C.c_ulonglong),
                ("last_prune_launches", C.c_ulonglong), ("last_levels", C.c_int), ("exponent_capacity", C.c_int),
                ("device_bytes", C.c_ulonglong), ("last_active_tiles", C.c_int),
                ("num_tips", C.c_int), ("num_loci", C.c_int), ("num_rate_cats", C.c_int),
                ("last_algorithmic_bytes", C.c_ulonglong), ("last_stored_vertices", C.c_int),
                ("last_fused_vertices", C.c_int), ("exact_mode", C.c_int),
                ("last_host_plan_us", C.c_double), ("last_host_submit_us", C.c_double), ("last_host_total_us", C.c_double),
                ("create_setup_us", C.c_double), ("create_upload_ms", C.c_double),
                ("last_chain_sites", C.c_int), ("last_chain_rows", C.c_int), ("chain_group", C.c_int),
                ("plan_cache_hits", C.c_ulonglong), ("graph_replays", C.c_ulonglong),
                ("last_walk_steps", C.c_int), ("last_materialised", C.c_int),
                ("last_cert", C.c_int), ("last_cert_rate", C.c_int), ("prune_ms_total", C.c_double), ("last_fp64_ops", C.c_ulonglong)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/dmpc.h declares: name -> (restype, argtypes)
_P = C.c_void_p
ABI = {
    "dmpc_default_options": (None, [C.POINTER(Options)]),
    "dmpc_last_error": (C.c_char_p, []),
    "dmpc_version": (C.c_char_p, []),
    "dmpc_convert_alignment": (C.c_long, [C.c_char_p, C.c_long, C.c_char_p, _u8_p]),
    "dmpc_loglik_create": (C.c_int, [_i32_p, C.c_int, _dbl_p, C.c_int, _dbl_p, _dbl_p, _dbl_p, C.c_int, _u8_p, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.POINTER(Options), C.POINTER(_P), _dbl_p]),
    "dmpc_new_between_transprobs_loglik": (C.c_int, [_P, _dbl_p, _dbl_p, C.c_int, _dbl_p, _dbl_p]),
    "dmpc_new_within_transprobs_loglik": (C.c_int, [_P, _dbl_p, _dbl_p, C.c_int, _dbl_p, _dbl_p]),
    "dmpc_within_clus_nni_loglik": (C.c_int, [_P, _dbl_p, _dbl_p, C.c_int, C.c_int, _dbl_p, _dbl_p, _i32_p]),
    "dmpc_between_clus_nni_loglik": (C.c_int, [_P, _dbl_p, _dbl_p, _dbl_p, C.c_int, C.c_int, _dbl_p, _dbl_p, _i32_p]),
    "dmpc_clus_split_merge_loglik": (C.c_int, [_P, _dbl_p, _dbl_p, _i32_p, C.c_int, _dbl_p, _dbl_p]),
    "dmpc_restore_previous_config": (C.c_int, [_P, _i32_p, C.c_int, C.c_int, C.c_int, C.c_int, _i32_p, C.c_int, C.c_int]),
    "dmpc_final_deallocate": (C.c_int, [_P]),
    "dmpc_get_max_map_size_estimate": (C.c_double, [C.c_double, C.c_uint]),
    "dmpc_check_and_cull_map": (C.c_int, [_P, C.c_uint, C.c_float, _dbl_p, _dbl_p, _dbl_p]),
    "dmpc_recompute_all": (C.c_int, [_P, _dbl_p, _dbl_p, _dbl_p, _dbl_p]),
    "dmpc_eval_proposals": (C.c_int, [_P, _dbl_p, _dbl_p, _dbl_p, C.c_void_p, C.c_int, _dbl_p, C.POINTER(C.c_int)]),
    "dmpc_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "dmpc_get_root_timeline": (C.c_int, [_P, _dbl_p]),
    "dmpc_timer_begin": (C.c_int, [_P]),
    "dmpc_timer_end": (C.c_int, [_P, _dbl_p]),
    "dmpc_plan_selfcheck": (C.c_int, [_i32_p, C.c_int, C.c_int, _dbl_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "dmpc_hp_create": (C.c_int, [_i32_p, C.c_int, C.c_int, _dbl_p, C.c_int, C.c_int, C.POINTER(_P)]),
    "dmpc_hp_propose": (C.c_int, [_P, C.c_int, _i32_p, C.c_int]),
    "dmpc_hp_plan": (C.c_int, [_P, _i32_p, C.c_int, _i32_p, C.c_int, _i32_p, C.c_int, _i32_p, C.c_int, _i32_p]),
    "dmpc_hp_path": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _i32_p, C.c_int, _i32_p, C.c_int, _i32_p, C.c_int, _i32_p]),
    "dmpc_hp_restore": (C.c_int, [_P, _i32_p, C.c_int, C.c_int, _i32_p, C.c_int, C.c_int]),
    "dmpc_hp_state": (C.c_int, [_P, _u8_p, _u8_p, _i32_p]),
    "dmpc_hp_destroy": (C.c_int, [_P]),
    "dmpc_get_partial": (C.c_int, [_P, C.c_int, _dbl_p, _f32_p]),
    "dmpc_get_topology": (C.c_int, [_P, _i32_p, _i32_p, _i32_p, _i32_p]),
    "dmpc_get_site_logliks": (C.c_int, [_P, _dbl_p]),
    "dmpc_set_deferred": (C.c_int, [_P, C.c_int]),
    "dmpc_shard_chain": (C.c_int, [_P, _f32_p, _f32_p]),
    "dmpc_shard_active": (C.c_int, [_P]),
    "dmpc_shard_num_chunks": (C.c_int, [_P]),
    "dmpc_shard_finish": (C.c_int, [_P, _dbl_p]),
    "dmpc_reduce_chunks": (C.c_double, [_dbl_p, C.c_int]),
    "dmpc_shard_range": (None, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dmpc_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P), C.c_void_p]),
    "dmpc_comm_connect": (C.c_int, [_P, C.c_void_p]),
    "dmpc_comm_destroy": (C.c_int, [_P]),
    "dmpc_shard_fast_result": (C.c_int, [_P, C.POINTER(C.c_int), _dbl_p]),
}


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: build it with `python -m dmphyclus_b200.build` "
                                "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    return lib


def _mats(m) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    if m.ndim != 3 or m.shape[1:] != (4, 4):
        raise ValueError(f"transition matrix list must be [R][4][4], got {m.shape}")
    return np.ascontiguousarray(m.transpose(0, 2, 1)).ravel()      # column-major per matrix, as R passes them


def _edge(e) -> np.ndarray:
    e = np.asarray(e)
    if e.ndim != 2 or e.shape[1] != 2:
        raise ValueError("edge matrix must be [E][2]")
    return np.ascontiguousarray(e.T, dtype=np.int32).ravel()


def _p(a, t):
    return a.ctypes.data_as(t)


class Engine:
    """The eleven entry points of the reference, evaluated on the GPU."""

    name = "cuda"

    def __init__(self, device: int = -1, quirk_carry: bool = True, site_begin: int = 0, sites_total: int = 0,
                 fuse_tips: int | None = None, fuse_exact: bool = False, chain_cert: bool = True, devices=None):
        self.lib = load_library()
        self.opts = Options()
        self.lib.dmpc_default_options(C.byref(self.opts))
        self.opts.device = device
        self.opts.quirk_carry = int(quirk_carry)
        self.opts.site_begin, self.opts.sites_total = site_begin, sites_total
        self.quirk_carry = bool(quirk_carry)
        if fuse_tips is not None:
            self.opts.fuse_tips = int(fuse_tips)
        self.opts.fuse_exact = int(fuse_exact)
        self.opts.chain_cert = int(chain_cert)
        if devices is not None and len(devices) > 1:
            # single-process multi-GPU: one handle whose sites are sharded over these devices (a device may repeat)
            if len(devices) > 8:
                raise ValueError("at most 8 devices per handle")
            self.opts.n_devices = len(devices)
            for i, d in enumerate(devices):
                self.opts.devices[i] = int(d)
        self._shape = {}
        self._mcache = {}
        self._ecache = {}
        self._lcache = {}

    def _check(self, rc):
        if rc != 0:
            raise DmpcError(rc, self.lib.dmpc_last_error().decode())

    def _mp(self, m):
        """Pointer to a transition-matrix list in the layout R passes (column-major 4x4 each).  Read-only arrays handed
        in again (the MCMC driver proposes from a fixed set of lists) are converted once; the returned ctypes pointer
        keeps its array alive."""
        ent = self._mcache.get(id(m))
        if ent is not None and ent[0] is m:
            return ent[1]
        ptr = _p(_mats(m), _dbl_p)
        if isinstance(m, np.ndarray) and not m.flags.writeable:
            if len(self._mcache) > 256:
                self._mcache.clear()
            self._mcache[id(m)] = (m, ptr)
        return ptr

    def _ep(self, e):
        """Pointer + row count of an edge matrix in column-major int32 (what R passes).  A read-only array handed in again
        (the committed phylogeny, on every rejected NNI) is converted once."""
        ent = self._ecache.get(id(e))
        if ent is not None and ent[0] is e:
            return ent[1], ent[2]
        arr = _edge(e)
        ptr = _p(arr, _i32_p)
        if isinstance(e, np.ndarray) and not e.flags.writeable:
            if len(self._ecache) >= 4:
                self._ecache.clear()
            self._ecache[id(e)] = (e, ptr, len(arr) // 2)
        return ptr, len(arr) // 2

    def _pp(self, pi):
        ent = self._mcache.get(id(pi))
        if ent is not None and ent[0] is pi:
            return ent[1]
        a = np.ascontiguousarray(pi, dtype=np.float64)
        if a.shape != (4,):
            raise ValueError("limProbs must have 4 entries")
        ptr = _p(a, _dbl_p)
        if isinstance(pi, np.ndarray) and not pi.flags.writeable:
            self._mcache[id(pi)] = (pi, ptr)
        return ptr

    def _lp(self, m, dtype, ptype):
        """Pointer + length of an id list (cluster MRCAs) as a contiguous `dtype` array.  The MCMC driver hands the same
        Python list to every call until a split / merge is accepted (paraValues$clusterNodeIndices): a list seen before is
        converted once — recognised by identity AND by comparing it with a copy taken then (lists are mutable)."""
        if type(m) is list:
            ent = self._lcache.get((id(m), dtype))
            if ent is not None and ent[0] is m and ent[1] == m:
                return ent[2], ent[3]
        arr = np.ascontiguousarray(np.asarray(m).ravel(), dtype=dtype)
        ptr = _p(arr, ptype)
        if type(m) is list:
            if len(self._lcache) >= 64:
                self._lcache.clear()
            self._lcache[(id(m), dtype)] = (m, list(m), ptr, len(arr), arr)
        return ptr, len(arr)

    @staticmethod
    def _edge_out(n_tips):
        """Output buffer of an NNI entry point: (what to pass as int32*, the [E][2] view of it the result owns).  Small
        matrices come as a ctypes array (no pointer object to build: 0.5 us against 2.7), large ones as numpy memory (not
        zero-filled first)."""
        n = 2 * (2 * n_tips - 2)
        if n <= 16384:
            buf = (C.c_int32 * n)()
            return buf, np.frombuffer(buf, np.int32).reshape(2, -1).T
        edge = np.empty(n, np.int32)
        return _p(edge, _i32_p), edge.reshape(2, -1).T

    # ------------------------------------------------------------------ the 11 entry points
    def getConvertedAlignment(self, equivVector, alignmentAlphaMat) -> np.ndarray:
        a = np.asarray(alignmentAlphaMat)
        if a.dtype != np.uint8:
            a = np.char.encode(a.astype(str), "ascii").view(np.uint8).reshape(a.shape)
        a = np.ascontiguousarray(a)
        states = "".join(equivVector).encode()
        if len(states) != 4:
            raise ValueError("4-state DNA only (names(limProbs) must have 4 entries)")
        out = np.empty(a.shape, np.uint8)
        bad = self.lib.dmpc_convert_alignment(a.ctypes.data_as(C.c_char_p), a.size, states, _p(out, _u8_p))
        if bad:
            raise ValueError(f"{bad} alignment cells hold a symbol outside the DNA/IUPAC alphabet")
        return out

    def logLikCpp(self, edgeMat, clusterMRCAs, limProbsVec, withinTransMatList, betweenTransMatList, numOpenMP,
                  alignmentBin, numTips, numLoci, withinMatListIndex, betweenMatListIndex, deferred=False):
        # (read-only inputs handed in again — logLikFromClusInd called in a loop — are converted once: _ep, _mp, _pp)
        e, n_rows = self._ep(edgeMat)
        m, n_m = self._lp(clusterMRCAs, np.float64, _dbl_p)
        pi = self._pp(limProbsVec)
        w, b = self._mp(withinTransMatList), self._mp(betweenTransMatList)
        n_rates = len(withinTransMatList)
        al = np.ascontiguousarray(alignmentBin, dtype=np.uint8)
        if al.shape != (numTips, numLoci) or n_rates != len(betweenTransMatList):
            raise ValueError("alignmentBin must be [numTips][numLoci]; matrix lists must have equal length")
        h, ll = _P(), C.c_double()
        self.opts.deferred = int(deferred)
        self._check(self.lib.dmpc_loglik_create(e, n_rows, m, n_m, pi,
                                                w, b, n_rates, _p(al, _u8_p), numTips,
                                                numLoci, int(withinMatListIndex), int(betweenMatListIndex),
                                                C.byref(self.opts), C.byref(h), C.byref(ll)))
        self._shape[h.value] = (numTips, numLoci, n_rates)
        return {"logLik": ll.value, "solutionPointer": h.value, "alignmentBinPointer": None}

    def newBetweenTransProbsLogLik(self, ptr, withinTransProbs, newBetweenTransProbs, numOpenMP,
                                   newBetweenMatListIndex, limProbs) -> float:
        w, b, pi = self._mp(withinTransProbs), self._mp(newBetweenTransProbs), self._pp(limProbs)
        ll = C.c_double()
        self._check(self.lib.dmpc_new_between_transprobs_loglik(ptr, w, b,
                                                                int(newBetweenMatListIndex), pi, C.byref(ll)))
        return ll.value

    def newWithinTransProbsLogLik(self, ptr, newWithinTransProbs, betweenTransProbs, numOpenMP,
                                  newWithinMatListIndex, limProbs) -> float:
        w, b, pi = self._mp(newWithinTransProbs), self._mp(betweenTransProbs), self._pp(limProbs)
        ll = C.c_double()
        self._check(self.lib.dmpc_new_within_transprobs_loglik(ptr, w, b,
                                                               int(newWithinMatListIndex), pi, C.byref(ll)))
        return ll.value

    def withinClusNNIlogLik(self, ptr, withinTransProbs, betweenTransProbs, MRCAofClusForNNI, numMovesNNI,
                            numOpenMP, limProbs):
        w, b, pi = self._mp(withinTransProbs), self._mp(betweenTransProbs), self._pp(limProbs)
        buf, edge = self._edge_out(self._shape[ptr][0])
        ll = C.c_double()
        self._check(self.lib.dmpc_within_clus_nni_loglik(ptr, w, b, int(MRCAofClusForNNI),
                                                         int(numMovesNNI), pi, C.byref(ll), buf))
        return {"logLik": ll.value, "edge": edge}      # [E][2] view of the column-major buffer (owned by the result)

    def betweenClusNNIlogLik(self, ptr, withinTransProbs, betweenTransProbs, clusterMRCAs, numMovesNNI,
                             numOpenMP, limProbs):
        w, b, pi = self._mp(withinTransProbs), self._mp(betweenTransProbs), self._pp(limProbs)
        m, n_m = self._lp(clusterMRCAs, np.float64, _dbl_p)
        buf, edge = self._edge_out(self._shape[ptr][0])
        ll = C.c_double()
        self._check(self.lib.dmpc_between_clus_nni_loglik(ptr, w, b, m, n_m,
                                                          int(numMovesNNI), pi, C.byref(ll), buf))
        return {"logLik": ll.value, "edge": edge}      # [E][2] view of the column-major buffer (owned by the result)

    def clusSplitMergeLogLik(self, ptr, withinTransProbs, betweenTransProbs, clusMRCAsToSplitOrMerge, numOpenMP,
                             limProbs) -> float:
        w, b, pi = self._mp(withinTransProbs), self._mp(betweenTransProbs), self._pp(limProbs)
        m = np.ascontiguousarray(np.asarray(clusMRCAsToSplitOrMerge).ravel(), dtype=np.int32)
        ll = C.c_double()
        self._check(self.lib.dmpc_clus_split_merge_loglik(ptr, w, b, _p(m, _i32_p), len(m),
                                                          pi, C.byref(ll)))
        return ll.value

    def RestorePreviousConfig(self, ptr, edgeMat, withinMatListIndex, betweenMatListIndex, NNImove, clusterMRCAs,
                              splitMergeMove) -> None:
        e, n_rows = self._ep(edgeMat)
        m, n_m = self._lp(clusterMRCAs, np.int32, _i32_p)
        self._check(self.lib.dmpc_restore_previous_config(ptr, e, n_rows, int(withinMatListIndex),
                                                          int(betweenMatListIndex), int(bool(NNImove)), m,
                                                          n_m, int(bool(splitMergeMove))))

    def finalDeallocate(self, ptr) -> None:
        self._shape.pop(ptr, None)
        self._check(self.lib.dmpc_final_deallocate(ptr))

    def getMaxMapSizeEstimate(self, allowedSizeInMegs, numRateCats) -> float:
        return self.lib.dmpc_get_max_map_size_estimate(float(allowedSizeInMegs), int(numRateCats))

    def checkAndCullMap(self, ptr, allowedNumberOfElements, cullProportion, withinTransProbs, betweenTransProbs,
                        limProbs) -> None:
        w, b, pi = self._mp(withinTransProbs), self._mp(betweenTransProbs), self._pp(limProbs)
        self._check(self.lib.dmpc_check_and_cull_map(ptr, int(allowedNumberOfElements), float(cullProportion),
                                                     w, b, pi))

    # ------------------------------------------------------------------ beyond the reference
    def recomputeAll(self, ptr, withinTransProbs, betweenTransProbs, limProbs) -> float:
        w, b, pi = self._mp(withinTransProbs), self._mp(betweenTransProbs), self._pp(limProbs)
        ll = C.c_double()
        self._check(self.lib.dmpc_recompute_all(ptr, w, b, pi, C.byref(ll)))
        return ll.value

    def evalProposals(self, ptr, withinTransProbs, betweenTransProbs, limProbs, proposals):
        """Batch of split / merge proposals against the committed state (left untouched).  `proposals`: iterable of
        ("split", mrca) or ("merge", mrcaA, mrcaB), 1-based ids.  Returns (logLik array, number evaluated by the batched kernel)."""
        w, b, pi = self._mp(withinTransProbs), self._mp(betweenTransProbs), self._pp(limProbs)
        arr = np.zeros((len(proposals), 3), np.int32)
        for i, q in enumerate(proposals):
            arr[i] = (0, q[1], 0) if q[0] == "split" else (1, q[1], q[2])
        out = np.empty(len(proposals), np.float64)
        nb = C.c_int()
        self._check(self.lib.dmpc_eval_proposals(ptr, w, b, pi, arr.ctypes.data_as(C.c_void_p),
                                                 len(proposals), _p(out, _dbl_p), C.byref(nb)))
        return out, nb.value

    def stats(self, ptr) -> dict:
        s = Stats()
        self._check(self.lib.dmpc_get_stats(ptr, C.byref(s)))
        return s.as_dict()

    def root_timeline(self, ptr) -> dict:
        """Where the last evaluation's root stage spent its time on the device (microseconds; None = phase did not run)."""
        out = np.zeros(6)
        self._check(self.lib.dmpc_get_root_timeline(ptr, _p(out, _dbl_p)))
        keys = ("gather", "exchange_or_reduce", "certificate", "gap_to_final", "final_terms", "final_exchange_or_reduce")
        return {k: (None if v < 0 else float(v)) for k, v in zip(keys, out)}

    def timer_begin(self, ptr) -> None:
        self._check(self.lib.dmpc_timer_begin(ptr))

    def timer_end(self, ptr) -> float:
        ms = C.c_double()
        self._check(self.lib.dmpc_timer_end(ptr, C.byref(ms)))
        return ms.value

    def partial(self, ptr, node):
        n, s, r = self._shape[ptr]
        sol = np.empty((s, r, 4), np.float64)
        ex = np.empty((s, r), np.float32)
        self._check(self.lib.dmpc_get_partial(ptr, int(node), _p(sol, _dbl_p), _p(ex, _f32_p)))
        return sol, ex

    def topology(self, ptr):
        n = 2 * self._shape[ptr][0] - 1
        arrs = [np.empty(n, np.int32) for _ in range(4)]
        self._check(self.lib.dmpc_get_topology(ptr, *[_p(a, _i32_p) for a in arrs]))
        return dict(zip(("parent", "child0", "child1", "within"), arrs))

    def site_logliks(self, ptr) -> np.ndarray:
        out = np.empty(self._shape[ptr][1], np.float64)
        self._check(self.lib.dmpc_get_site_logliks(ptr, _p(out, _dbl_p)))
        return out

    # multi-GPU site sharding (see dmphyclus_b200/sharded.py)
    def set_deferred(self, ptr, on: bool) -> None:
        self._check(self.lib.dmpc_set_deferred(ptr, int(on)))

    def shard_chain(self, ptr, carry_in) -> np.ndarray:
        r = self._shape[ptr][2]
        cin = np.zeros(8, np.float32)
        cin[:r] = np.asarray(carry_in, np.float32)[:r]
        cout = np.zeros(8, np.float32)
        self._check(self.lib.dmpc_shard_chain(ptr, _p(cin, _f32_p), _p(cout, _f32_p)))
        return cout[:r].copy()

    def shard_active(self, ptr) -> bool:
        """True when a site of this handle rescaled in the last evaluation (its chain is not the identity)."""
        return self.lib.dmpc_shard_active(ptr) > 0

    def shard_finish(self, ptr) -> np.ndarray:
        n = self.lib.dmpc_shard_num_chunks(ptr)
        out = np.empty(n, np.float64)
        self._check(self.lib.dmpc_shard_finish(ptr, _p(out, _dbl_p)))
        return out

    def shard_fast_result(self, ptr):
        """(True, logLik) when the reduction kernel's peer exchange finished the evaluation on the device."""
        ok, ll = C.c_int(), C.c_double()
        self._check(self.lib.dmpc_shard_fast_result(ptr, C.byref(ok), C.byref(ll)))
        return bool(ok.value), ll.value

    def comm_setup(self, rank: int, world: int, exchange, max_chunks: int = 4096) -> None:
        """Create this rank's NVLink mailbox, swap IPC handles with the peers through `exchange` (a callable
        bytes -> list of every rank's bytes, e.g. over torch.distributed) and attach it to later handles."""
        comm = _P()
        mine = C.create_string_buffer(64)
        device = self.opts.device if self.opts.device >= 0 else 0
        self._check(self.lib.dmpc_comm_create(device, rank, world, max_chunks, C.byref(comm), mine))
        allh = b"".join(exchange(mine.raw))
        if len(allh) != 64 * world:
            raise RuntimeError("IPC handle exchange returned the wrong amount of data")
        self._check(self.lib.dmpc_comm_connect(comm, allh))
        self._comm = comm
        self.opts.comm = comm.value

    def comm_close(self) -> None:
        if getattr(self, "_comm", None):
            self.opts.comm = None
            self._check(self.lib.dmpc_comm_destroy(self._comm))
            self._comm = None

    def reduce_chunks(self, chunks) -> float:
        c = np.ascontiguousarray(chunks, dtype=np.float64)
        return self.lib.dmpc_reduce_chunks(_p(c, _dbl_p), len(c))
