"""The engine's host side alone (dmphyclus_b200/csrc/host_tree.hpp: topology, proposal state machine, plan builder,
NNI undo log, incremental edge export) compiled with AddressSanitizer + UBSan and driven through random proposal
sequences and damaged edge matrices (tools/host_bench.cpp --fuzz).  What it guards: the flat arrays that replace the
reference's pointer graph (/root/reference/src/TreeNode.h:80-125, AugTree.cpp:59-101, 144-271, 328-374) are never read
or written out of range, whatever R hands in, and every plan keeps the invariants the kernels rely on."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "host_bench.cpp")


@pytest.fixture(scope="module")
def fuzzer(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("hostfuzz") / "host_fuzz")
    res = subprocess.run([gxx, "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                          "-o", exe, SRC], capture_output=True, text=True)
    if res.returncode != 0 and "sanitize" in res.stderr and "cannot find" in res.stderr:
        pytest.skip("sanitizer runtimes are not installed")
    assert res.returncode == 0, res.stderr
    return exe


@pytest.mark.parametrize("seed,threads", [(1, 1), (2, 3)])
def test_host_state_machine_under_sanitizers(fuzzer, seed, threads):
    """threads > 1: several state machines at once on one recycled-storage pool (PooledVec, host_tree.hpp), as the shards
    of a multi-device handle are created and destroyed on their worker threads."""
    res = subprocess.run([fuzzer, "--fuzz", "250", str(seed), str(threads)], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("fuzz ok") == threads, res.stdout + res.stderr


def test_host_timings_run(tmp_path):
    """The timing mode (what DESIGN.md §8 quotes) builds without sanitizers and runs to the end."""
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "host_bench")
    subprocess.run([gxx, "-O2", "-std=c++17", "-pthread", "-o", exe, SRC], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "config5 100,000 tips, caterpillar" in res.stdout
