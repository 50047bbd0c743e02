"""The C-ABI library: it loads without a GPU, exports every symbol include/dmpc.h declares (and the ctypes table of
engine.py covers exactly those), host-only entry points work, and compute entry points FAIL LOUDLY without a
device — there is no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dmphyclus_b200 import build
    from dmphyclus_b200.engine import load_library
    build.build()
    return load_library()


def declared_symbols(header="dmpc.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(dmpc_[a-z0-9_]+)\s*\(", src))


def test_header_symbols_exported_and_bound(lib):
    from dmphyclus_b200.engine import ABI
    syms = declared_symbols()
    hooks = declared_symbols("dmpc_test.h")            # host-only test hooks: same library, separate header
    assert len(syms) >= 25 and hooks and not (syms & hooks)
    assert all(h.startswith(("dmpc_hp_", "dmpc_plan_selfcheck")) for h in hooks), hooks
    assert not any(s.startswith("dmpc_hp_") for s in syms)        # the product header carries no test hook
    assert syms | hooks == set(ABI), (syms | hooks) ^ set(ABI)
    for s in syms | hooks:
        assert getattr(lib, s) is not None


def test_reference_entry_points_all_present():
    """One C symbol per Rcpp-exported routine of the reference (src/RcppExports.cpp:179-192)."""
    want = {"logLikCpp": "dmpc_loglik_create", "getConvertedAlignment": "dmpc_convert_alignment",
            "newBetweenTransProbsLogLik": "dmpc_new_between_transprobs_loglik",
            "newWithinTransProbsLogLik": "dmpc_new_within_transprobs_loglik",
            "withinClusNNIlogLik": "dmpc_within_clus_nni_loglik", "betweenClusNNIlogLik": "dmpc_between_clus_nni_loglik",
            "clusSplitMergeLogLik": "dmpc_clus_split_merge_loglik", "RestorePreviousConfig": "dmpc_restore_previous_config",
            "finalDeallocate": "dmpc_final_deallocate", "getMaxMapSizeEstimate": "dmpc_get_max_map_size_estimate",
            "checkAndCullMap": "dmpc_check_and_cull_map"}
    syms = declared_symbols()
    from dmphyclus_b200.engine import Engine
    for r_name, c_name in want.items():
        assert c_name in syms
        assert callable(getattr(Engine, r_name))


def test_host_only_entry_points(lib, oracle):
    chars = np.frombuffer(b"atcgryswkmbdhv-n.", np.uint8).copy()
    out = np.empty_like(chars)
    bad = lib.dmpc_convert_alignment(chars.ctypes.data_as(C.c_char_p), chars.size, b"atcg",
                                     out.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert bad == 0
    assert np.array_equal(out, oracle.getConvertedAlignment(list("atcg"), chars.reshape(1, -1))[0])
    assert lib.dmpc_get_max_map_size_estimate(100.0, 3) == oracle.getMaxMapSizeEstimate(100.0, 3)
    x = np.random.default_rng(0).standard_normal(1000)
    got = lib.dmpc_reduce_chunks(x.ctypes.data_as(C.POINTER(C.c_double)), len(x))
    assert abs(got - x.sum()) < 1e-10
    assert b"sm_100a" in lib.dmpc_version()


def test_no_cpu_fallback(lib, packaged):
    """Without a GPU the compute entry points return DMPC_ERR_CUDA — they never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the failure path is exercised on CPU-only boxes")
    from dmphyclus_b200.engine import DmpcError, Engine
    from dmphyclus_b200.workloads import Problem
    p = Problem(packaged, 8, 10, 2, 3, "iid", seed=1)
    with pytest.raises(DmpcError) as e:
        p.create(Engine())
    assert e.value.code == -2


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under dmphyclus_b200/ may import, load or call it."""
    pkg = os.path.join(ROOT, "dmphyclus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liborc" not in src and "libref" not in src, f
