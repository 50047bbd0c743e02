"""The N > 1 path on CPU: world_size-2 gloo process group, each rank evaluating its site block (oracle-backed
stand-in for the CUDA handle) and meeting the other rank only in the root reduction — the fp32 exponent chain hops
rank 0 -> rank 1 and the chunk sums are all-gathered (dmphyclus_b200/sharded.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, case, ret):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    from cpu_shard import OracleShard
    from dmphyclus_b200.sharded import ShardedEngine, shard_range
    from dmphyclus_b200.workloads import Problem, load_packaged
    from oracle.oracle import Oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, S, k, R, kind, seed = case
        p = Problem(load_packaged(), n, S, k, R, kind, seed=seed)
        full = Oracle()
        eng = ShardedEngine(OracleShard(Oracle()), S)
        assert (eng.begin, eng.end) == shard_range(S, rank, world)
        al = full.getConvertedAlignment(list("atcg"), p.chars)
        W, B = p.W, p.B
        got, want = [], []
        r1 = eng.logLikCpp(p.edge, p.mrcas, p.pi, W[4], B[4], 1, al, n, S, 5, 5)
        r2 = full.logLikCpp(p.edge, p.mrcas, p.pi, W[4], B[4], 1, al, n, S, 5, 5)
        h1, h2 = r1["solutionPointer"], r2["solutionPointer"]
        got.append(r1["logLik"]); want.append(r2["logLik"])
        got.append(eng.newBetweenTransProbsLogLik(h1, W[4], B[1], 1, 2, p.pi))
        want.append(full.newBetweenTransProbsLogLik(h2, W[4], B[1], 1, 2, p.pi))
        for be, h in ((eng, h1), (full, h2)):
            be.RestorePreviousConfig(h, np.zeros((1, 2), np.int32), 5, 5, False, p.mrcas, False)
        m = [x for x in p.mrcas if x > n][0]
        got.append(eng.clusSplitMergeLogLik(h1, W[4], B[4], [m], 1, p.pi))
        want.append(full.clusSplitMergeLogLik(h2, W[4], B[4], [m], 1, p.pi))
        a, b = eng.withinClusNNIlogLik(h1, W[4], B[4], m, 1, 1, p.pi), full.withinClusNNIlogLik(h2, W[4], B[4], m, 1, 1, p.pi)
        assert np.array_equal(a["edge"], b["edge"])
        got.append(a["logLik"]); want.append(b["logLik"])
        eng.finalDeallocate(h1); full.finalDeallocate(h2)
        ret[rank] = (got, want, eng.collectives)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [(300, 600, 5, 3, "iid", 31), (40, 300, 4, 1, "evolved", 32), (700, 520, 6, 3, "diverged", 33)])
def test_two_rank_site_sharding_matches_single_process(case):
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() + case[-1]) % 2000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, case, ret), nprocs=world, join=True)
        assert len(ret) == world
        for rank in range(world):
            got, want, ncoll = ret[rank]
            assert ncoll > 0
            # fp32 chain reproduced across the rank boundary; only the fp64 summation order over sites differs
            assert np.allclose(got, want, rtol=1e-12, atol=0), (rank, got, want)
        assert ret[0][0] == ret[1][0]          # every rank returns the same value (SPMD drivers stay in lock-step)


def test_shard_ranges_are_chunk_aligned_and_cover():
    from dmphyclus_b200.sharded import CHUNK, shard_range
    from dmphyclus_b200.workloads import CONFIGS
    assert CHUNK == 64          # RCH of dmphyclus_b200/csrc/kernels.cuh: decoupled from the 256-site prune tile
    for S in (1, 255, 256, 257, 10000, 30000, 1500):
        for world in (1, 2, 4, 8):
            rs = [shard_range(S, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == S
            for (b0, e0), (b1, e1) in zip(rs, rs[1:]):
                assert e0 == b1 and (b1 % CHUNK == 0 or b1 == S)
    # every BASELINE config can run at 1, 2, 4 and 8 GPUs: no rank is left without sites (config 5 has only 1,500)
    for name, (_, S, _, _) in CONFIGS.items():
        for world in (1, 2, 4, 8):
            for r in range(world):
                b, e = shard_range(S, r, world)
                assert e > b, (name, world, r)


def test_library_and_python_agree_on_the_shard_layout():
    """dmpc_shard_range (the layout of a multi-device handle inside libdmpc) == sharded.shard_range (process per GPU)."""
    import ctypes as C
    from dmphyclus_b200 import build
    from dmphyclus_b200.engine import load_library
    from dmphyclus_b200.sharded import shard_range
    build.build()
    lib = load_library()
    b, e = C.c_int(), C.c_int()
    for S in list(range(1, 70)) + [127, 128, 129, 1499, 1500, 1501, 4095, 4096, 10000, 30000, 123457]:
        for world in range(1, 9):
            for r in range(world):
                lib.dmpc_shard_range(S, r, world, C.byref(b), C.byref(e))
                assert (b.value, e.value) == shard_range(S, r, world), (S, world, r)
