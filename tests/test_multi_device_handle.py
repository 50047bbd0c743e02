"""Single-process multi-device handles (dmpc_options.n_devices / DMPC_DEVICES): ONE handle whose sites are sharded over
several devices — what an R session, being one process, needs to use the 8 GPUs of a box with the R API unchanged
(/root/reference/R/DMphyClusInterface.R:71).  The shards meet inside the root kernels through the same peer mailboxes
as the process-per-GPU path, so a device may be listed twice: on a one-GPU box these tests still drive the IN-KERNEL
exchange (chunk sums, certificate records, the literal chain's hand-off).  Every value must equal the single-device
handle's BIT FOR BIT."""
import numpy as np
import pytest

from helpers import Problem, compare_to_golden, rel, scripted_moves

pytestmark = pytest.mark.gpu
TOL = 1e-9

CASES = [
    # tips, sites, clusters, rate cats, alignment, seed, chain_cert
    (60, 700, 4, 3, "evolved", 31, True),        # nothing rescales: the gather's exchange finishes the evaluation
    (10000, 1200, 100, 3, "evolved", 32, True),  # config 3's tree: certificate, literal prefix inside shard 0
    (700, 900, 6, 3, "iid", 33, True),           # rescale-heavy
    (700, 900, 6, 3, "iid", 33, False),          # ... with the literal chain handed from shard to shard
    (1024, 900, 6, 3, "diverged", 34, True),
    (300, 1000, 4, 1, "iid", 35, True),          # one rate category: no chain at all
    (400, 520, 5, 3, "tiny", 36, True),          # fused clades that rescale: every shard re-runs in exact mode
]


def _devices(parts):
    import torch
    nd = torch.cuda.device_count()
    return [i % nd for i in range(parts)]


@pytest.mark.parametrize("n,S,k,R,kind,seed,cert", CASES)
@pytest.mark.parametrize("parts", [2, 3])
def test_shards_of_one_handle_equal_single_device_bitwise(packaged, n, S, k, R, kind, seed, cert, parts):
    from dmphyclus_b200.engine import Engine
    p = Problem(packaged, n, S, k, R, kind, seed=seed)
    single, multi = Engine(chain_cert=cert), Engine(chain_cert=cert, devices=[0] * parts)
    hs, ls = p.create(single)
    hm, lm = p.create(multi)
    assert lm == ls, (lm, ls)
    assert np.array_equal(single.site_logliks(hs), multi.site_logliks(hm))
    for keep in (False, True):
        a, b = scripted_moves(single, hs, p, keep), scripted_moves(multi, hm, p, keep)
        assert a == b, (a, b)
    sa, sb = single.partial(hs, p.n), multi.partial(hm, p.n)
    assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1])
    st = multi.stats(hm)
    assert st["num_loci"] == S and st["num_tips"] == n
    single.finalDeallocate(hs); multi.finalDeallocate(hm)


def test_more_devices_than_site_blocks(packaged):
    """A short alignment leaves trailing shards without sites: they are simply not created."""
    from dmphyclus_b200.engine import Engine
    p = Problem(packaged, 200, 150, 4, 3, "evolved", seed=41)     # 150 sites = 3 chunks of 64
    single, multi = Engine(), Engine(devices=[0] * 8)
    hs, ls = p.create(single)
    hm, lm = p.create(multi)
    assert lm == ls
    assert scripted_moves(single, hs, p, False) == scripted_moves(multi, hm, p, False)
    single.finalDeallocate(hs); multi.finalDeallocate(hm)


def test_real_devices_when_present(packaged):
    import torch
    from dmphyclus_b200.engine import Engine
    nd = torch.cuda.device_count()
    if nd < 2:
        pytest.skip("one GPU on this box (the same-device tests above cover the exchange)")
    for n, S, k, R, kind, seed in ((10000, 2000, 100, 3, "evolved", 51), (700, 2100, 6, 3, "iid", 52)):
        p = Problem(packaged, n, S, k, R, kind, seed=seed)
        single, multi = Engine(device=0), Engine(devices=list(range(min(nd, 8))))
        hs, ls = p.create(single)
        hm, lm = p.create(multi)
        assert lm == ls
        assert scripted_moves(single, hs, p, True) == scripted_moves(multi, hm, p, True)
        single.finalDeallocate(hs); multi.finalDeallocate(hm)


def test_rcpp_shim_uses_dmpc_devices(packaged, oracle, monkeypatch):
    """The drop-in shim (rpkg/src/clusterFunArmadillo.cpp) reads DMPC_DEVICES: same scripted sequence as the oracle,
    every call going shim -> C-ABI -> two shards -> in-kernel exchange."""
    from oracle.oracle import Reference, build_shimtest
    from script import run_script
    path = build_shimtest()
    if path is None:
        pytest.skip("libshimtest.so not built")
    monkeypatch.setenv("DMPC_DEVICES", "0,0")
    shim = Reference(path)
    for n, S, k, R, kind in ((40, 300, 5, 3, "evolved"), (500, 200, 6, 3, "iid")):
        p = Problem(packaged, n, S, k, R, kind, seed=n)
        compare_to_golden(run_script(shim, p), run_script(oracle, p), ll_tol=TOL, exact_partials=True)
