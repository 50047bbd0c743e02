"""In-tree build of the CUDA engine: `dmphyclus_b200/libdmpc.so` (sm_100a only).

    python -m dmphyclus_b200.build [--force]

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", f) for f in ("dmpc_engine.cu", "kernels.cuh", "host_tree.hpp", "task.hpp")]
HDR = os.path.join(ROOT, "include", "dmpc.h")
HDR_TEST = os.path.join(ROOT, "include", "dmpc_test.h")
LIB = os.path.join(HERE, "libdmpc.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"]


def build(force: bool = False, verbose: bool = False) -> str:
    deps = SRC + [HDR, HDR_TEST]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC[0]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


def build_fp64_peak(force: bool = False) -> str:
    """tools/libfp64peak.so: the DMUL / DADD issue-rate microbenchmark bench.py measures its roofline peak with."""
    src = os.path.join(ROOT, "tools", "fp64_peak.cu")
    lib = os.path.join(ROOT, "tools", "libfp64peak.so")
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= os.path.getmtime(src):
        return lib
    res = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
                          "-o", lib, src], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return lib


if __name__ == "__main__":
    build_fp64_peak(force="--force" in sys.argv)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
