"""Site sharding over real GPUs (needs >= 2 on the box; skipped otherwise): one process per GPU, NCCL for the host
protocol, NVLink peer mailboxes for the in-kernel exchange.  Every sharded value must equal the single-GPU value BIT
FOR BIT, on the fast path (no rescaled site: the reduction kernel finishes by itself) and on the slow one."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, case, ret):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    from dmphyclus_b200.engine import Engine
    from dmphyclus_b200.sharded import ShardedEngine, shard_range
    from dmphyclus_b200.workloads import Problem, load_packaged
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n, S, k, R, kind, seed = case
        p = Problem(load_packaged(), n, S, k, R, kind, seed=seed)
        single = Engine(device=rank)
        local = Engine(device=rank, site_begin=shard_range(S, rank, world)[0], sites_total=S)
        eng = ShardedEngine(local, S, device=torch.device("cuda", rank))
        al = single.getConvertedAlignment(list("atcg"), p.chars)
        W, B = p.W, p.B
        got, want = [], []
        r1 = eng.logLikCpp(p.edge, p.mrcas, p.pi, W[4], B[4], 1, al, n, S, 5, 5)
        r2 = single.logLikCpp(p.edge, p.mrcas, p.pi, W[4], B[4], 1, al, n, S, 5, 5)
        h1, h2 = r1["solutionPointer"], r2["solutionPointer"]
        got.append(r1["logLik"]); want.append(r2["logLik"])
        for be, h, out in ((eng, h1, got), (single, h2, want)):
            out.append(be.newBetweenTransProbsLogLik(h, W[4], B[1], 1, 2, p.pi))
            be.RestorePreviousConfig(h, np.zeros((1, 2), np.int32), 5, 5, False, p.mrcas, False)
            out.append(be.newWithinTransProbsLogLik(h, W[6], B[4], 1, 7, p.pi))
            m = [x for x in p.mrcas if x > n][0]
            out.append(be.clusSplitMergeLogLik(h, W[6], B[4], [m], 1, p.pi))
            out.append(be.withinClusNNIlogLik(h, W[6], B[4], m, 1, 1, p.pi)["logLik"])
            out.append(be.recomputeAll(h, W[6], B[4], p.pi))
        eng.finalDeallocate(h1); single.finalDeallocate(h2)
        ret[rank] = (got, want, eng.fast, eng.collectives)
        local.comm_close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [(300, 2600, 5, 3, "evolved", 41), (700, 2600, 6, 3, "iid", 42), (200, 2700, 4, 1, "iid", 43),
                                  (400, 2600, 5, 3, "tiny", 44), (1024, 2600, 6, 3, "diverged", 45)])
def test_sharded_matches_single_gpu_bitwise(case):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() + case[-1]) % 2000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, case, ret), nprocs=world, join=True)
        assert len(ret) == world
        for rank in range(world):
            got, want, fast, ncoll = ret[rank]
            assert got == want, (rank, got, want)
            # every evaluation finished on the device: chunk sums (and, for the iid cases, the fp32 exponent vector of
            # the cross-locus chain) travelled through the NVLink mailboxes, no host-side collective ran
            assert fast == len(got) and ncoll == 0
