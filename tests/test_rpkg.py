"""The R package skeleton (rpkg/): what a maintainer installs so that DMphyClusChain & co. run on libdmpc with the R
code unchanged.  There is no R in this image, so the checks are: the SEXP-level glue (rpkg/src/RcppExports.cpp) compiles
against the Rcpp stand-in together with the shim, registers the reference's eleven routines with the reference's
argument counts (/root/reference/src/RcppExports.cpp:179-192), and a call through a registered entry point works; the R
stubs' formal arguments are the reference's (R/RcppExports.R:4-46); configure and assemble.sh do what they say."""
import ctypes as C
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RPKG = os.path.join(ROOT, "rpkg")

# /root/reference/src/RcppExports.cpp:179-192
ROUTINES = {"logLikCpp": 11, "getConvertedAlignment": 2, "newBetweenTransProbsLogLik": 6, "newWithinTransProbsLogLik": 6,
            "withinClusNNIlogLik": 7, "betweenClusNNIlogLik": 7, "clusSplitMergeLogLik": 6, "RestorePreviousConfig": 7,
            "finalDeallocate": 1, "getMaxMapSizeEstimate": 2, "checkAndCullMap": 6}


@pytest.fixture(scope="module")
def shimlib():
    from dmphyclus_b200 import build
    from oracle.oracle import build_shimtest
    build.build()
    path = build_shimtest()
    if path is None:
        pytest.skip("libshimtest.so not built")
    return C.CDLL(path)


def test_glue_registers_the_reference_routine_table(shimlib):
    names = C.create_string_buffer(4096)
    nargs = (C.c_int * 32)()
    n = shimlib.ref_registered_routines(names, 4096, nargs, 32)
    got = [x for x in names.value.decode().split(";") if x]
    assert n == 11 and got == ["_DMphyClus_" + k for k in ROUTINES]          # same symbols, same order as the reference's table
    assert list(nargs[:n]) == list(ROUTINES.values())
    for k in ROUTINES:
        assert hasattr(shimlib, "_DMphyClus_" + k)                            # exported with C linkage, as .Call looks them up
    assert hasattr(shimlib, "R_init_DMphyClus")


def test_call_through_a_registered_entry_point(shimlib):
    f = shimlib.ref_glue_max_map_size_estimate
    f.restype, f.argtypes = C.c_double, [C.c_double, C.c_uint]
    # getMaxMapSizeEstimate, clusterFunArmadillo.cpp:298-302: floor(MB * 1e6 / (sizeof(S) + numRateCats * sizeof(pair) + 16))
    assert f(100.0, 3) == float(int(100e6 // (32 + 24 * 3 + 16)))


def _stub_table():
    src = open(os.path.join(RPKG, "R", "RcppExports.R")).read()
    body = src[src.index(".dmpcEntryPoints <- list("):src.index("# routines that return nothing")]
    table = {}
    for name, args in re.findall(r"(\w+)\s*=\s*c\(([^)]*)\)", body):
        table[name] = re.findall(r'"(\w+)"', args)
    return table


def test_r_stubs_have_the_reference_formals():
    table = _stub_table()
    assert {k: len(v) for k, v in table.items()} == ROUTINES
    ref = "/root/reference/R/RcppExports.R"
    if os.path.exists(ref):                                                    # here; absent on the GPU box
        want = {m.group(1): [a.strip() for a in m.group(2).split(",")]
                for m in re.finditer(r"^(\w+) <- function\(([^)]*)\)", open(ref).read(), re.M)}
        assert table == want


def test_skeleton_files_and_configure(tmp_path):
    desc = open(os.path.join(RPKG, "DESCRIPTION")).read()
    assert re.search(r"^LinkingTo: Rcpp$", desc, re.M) and "RcppArmadillo" not in desc and "RcppGSL" not in desc
    ns = open(os.path.join(RPKG, "NAMESPACE")).read()
    for f in ("DMphyClusChain", "logLikFromClusInd", "clusIndLogPrior", "outputTransMatList"):     # NAMESPACE:3-6 of the reference
        assert f"export({f})" in ns
    assert "useDynLib(DMphyClus" in ns
    work = tmp_path / "rpkg"
    shutil.copytree(RPKG, work)
    env = dict(os.environ, DMPC_HOME=ROOT)
    subprocess.check_call(["sh", str(work / "configure")], env=env, stdout=subprocess.DEVNULL)
    mk = (work / "src" / "Makevars").read_text()
    assert f"-I{ROOT}/include" in mk and f"-L{ROOT}/dmphyclus_b200 -ldmpc" in mk
    env = dict(os.environ, DMPC_HOME=str(tmp_path / "nowhere"), DMPC_INCLUDE="", DMPC_LIBDIR="")
    bad = subprocess.run(["sh", str(work / "configure")], env=env, capture_output=True, text=True, cwd=str(tmp_path))
    # (falls back to the checkout the skeleton sits in — work/.. has no include/, so it must fail loudly)
    assert bad.returncode != 0 and "cannot find" in bad.stderr


def test_assemble_overlays_the_upstream_package(tmp_path):
    up = "/root/reference"
    if not os.path.isdir(os.path.join(up, "R")):
        pytest.skip("upstream package tree not available on this box")
    out = tmp_path / "DMphyClus.b200"
    subprocess.check_call(["sh", os.path.join(RPKG, "assemble.sh"), up, str(out)], stdout=subprocess.DEVNULL)
    assert sorted(os.listdir(out / "src")) == ["Makevars", "Makevars.in", "RcppExports.cpp", "clusterFunArmadillo.cpp"]
    for f in ("DMphyClusInterface.R", "DMphyClusSource.R", "RcppExports.R"):
        assert (out / "R" / f).exists()
    assert (out / "R" / "DMphyClusSource.R").read_bytes() == open(os.path.join(up, "R", "DMphyClusSource.R"), "rb").read()
    assert ".dmpcEntryPoints" in (out / "R" / "RcppExports.R").read_text()
