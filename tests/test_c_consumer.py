"""include/dmpc.h is a C header, not a C++ one: it compiles as strict C99, and a plain-C program (examples/loglik_c99.c:
create, propose, reject, release — no C++, Python or torch in sight) links against libdmpc.so and runs.  Without a GPU the
library must refuse loudly (DMPC_ERR_CUDA, exit status 3 of the example): there is no CPU path to fall back to."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gcc():
    exe = shutil.which("gcc")
    if exe is None:
        pytest.skip("no gcc")
    return exe


@pytest.mark.parametrize("header", ["dmpc.h", "dmpc_test.h"])
def test_headers_are_strict_c99(gcc, header):
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c",
                          os.path.join(ROOT, "include", header)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_c_program_links_and_fails_loudly_without_a_gpu(gcc, tmp_path):
    import torch
    from dmphyclus_b200 import build
    build.build()
    libdir = os.path.join(ROOT, "dmphyclus_b200")
    exe = str(tmp_path / "loglik_c99")
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                          os.path.join(ROOT, "examples", "loglik_c99.c"), "-L" + libdir, "-ldmpc", "-Wl,-rpath," + libdir,
                          "-o", exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert "masks of tip 3: 2248924f" in run.stdout            # t t c g r t c n  ->  2 2 4 8 (a|g = 9) 2 4 f
    if torch.cuda.is_available():
        assert run.returncode == 0, run.stdout + run.stderr
        assert "logLik" in run.stdout
    else:
        assert run.returncode == 3, run.stdout + run.stderr
        assert "no CUDA device" in run.stderr and "logLik" not in run.stdout
