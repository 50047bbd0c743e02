"""The chain certificate (k_cert, dmphyclus_b200/csrc/kernels.cuh): where the reference's sequential fp32 exponent
chain (/root/reference/src/AugTree.cpp:306-323) is provably in its one-dominant-category regime, the per-site terms
are computed in parallel from a closed form.  The result must equal the literal walk BIT FOR BIT (log-likelihood and
every per-site term) and the oracle within 1e-9; where the certificate does not hold the literal chain must run."""
import numpy as np
import pytest

from helpers import Problem, rel, scripted_moves as _moves

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _pair():
    from dmphyclus_b200.engine import Engine
    return Engine(chain_cert=True), Engine(chain_cert=False)


@pytest.mark.parametrize("n,S,k,R,kind,seed,expect", [
    (10000, 1200, 100, 3, "evolved", 11, 1),  # config 3's tree: a third of the sites rescale, one category takes over
    (1000, 2100, 8, 3, "iid", 12, None),       # ~12 rescales per site and rate
    (1024, 900, 6, 3, "diverged", 13, None),   # the dominant category differs from site to site
    (800, 700, 6, 2, "iid", 14, None),
    (900, 520, 5, 4, "iid", 15, None),
    (700, 300, 5, 5, "iid", 16, None),         # RP = 8 variants of the kernels
])
def test_certificate_equals_literal_chain(packaged, oracle, n, S, k, R, kind, seed, expect):
    p = Problem(packaged, n, S, k, R, kind, seed=seed)
    cert, lit = _pair()
    hc, lc = p.create(cert)
    hl, ll = p.create(lit)
    ho, lo = p.create(oracle)
    sc, sl = cert.stats(hc), lit.stats(hl)
    assert sl["last_cert"] in (0, 2)
    if expect is not None:
        assert sc["last_cert"] == expect, sc
        assert 0 < sc["last_chain_sites"] < sl["last_chain_sites"]      # only the prefix was walked
    assert lc == ll, (lc, ll, sc["last_cert"])
    assert rel(lc, lo) < TOL
    assert np.array_equal(cert.site_logliks(hc), lit.site_logliks(hl))
    for keep in (False, True):
        a, b, c = _moves(cert, hc, p, keep), _moves(lit, hl, p, keep), _moves(oracle, ho, p, keep)
        assert a == b, (a, b)
        assert all(rel(x, y) < TOL for x, y in zip(a, c)), (a, c)
    for be, h in ((cert, hc), (lit, hl), (oracle, ho)):
        be.finalDeallocate(h)


def test_certificate_status_is_reported(packaged):
    """dmpc_stats tells how the chain was resolved: 0 = no rescaled site, 1 = certificate, 2 = literal."""
    from dmphyclus_b200.engine import Engine
    eng = Engine()
    quiet = Problem(packaged, 60, 300, 4, 3, "evolved", seed=21)            # small tree: nothing rescales
    h, _ = quiet.create(eng)
    assert eng.stats(h)["last_cert"] == 0 and eng.stats(h)["last_cert_rate"] == -1
    eng.finalDeallocate(h)
    big = Problem(packaged, 10000, 600, 100, 3, "evolved", seed=22)
    h, _ = big.create(eng)
    st = eng.stats(h)
    assert st["last_cert"] == 1 and 0 <= st["last_cert_rate"] < 3
    # the second evaluation queues gather + certificate + final at once (no host round trip): same value
    l1 = eng.recomputeAll(h, big.W[4], big.B[4], big.pi)
    l2 = eng.recomputeAll(h, big.W[4], big.B[4], big.pi)
    assert l1 == l2 and eng.stats(h)["last_cert"] == 1
    eng.finalDeallocate(h)


def test_textbook_mode_unaffected(packaged, oracle):
    from dmphyclus_b200.engine import Engine
    from oracle.oracle import Oracle
    p = Problem(packaged, 1000, 600, 6, 3, "iid", seed=23)
    eng, orc = Engine(quirk_carry=False), Oracle(quirk_carry=False)
    h, l = p.create(eng)
    ho, lo = p.create(orc)
    assert rel(l, lo) < TOL and eng.stats(h)["last_cert"] == 0
    eng.finalDeallocate(h); orc.finalDeallocate(ho)


def test_root_timeline_and_kernel_time_counters(packaged):
    """dmpc_get_root_timeline (globaltimer stamps of the root kernels) and dmpc_stats.prune_ms_total (the pruning launches'
    device time, accumulated without synchronisation): what bench.py builds its roofline and root-stage breakdown from."""
    from dmphyclus_b200.engine import Engine
    eng = Engine()
    p = Problem(packaged, 10000, 600, 100, 3, "evolved", seed=31)
    h, _ = p.create(eng)
    s0 = eng.stats(h)
    for _ in range(3):
        eng.recomputeAll(h, p.W[4], p.B[4], p.pi)
    s1 = eng.stats(h)
    assert s1["prune_ms_total"] > s0["prune_ms_total"] > 0 and s1["last_prune_ms"] > 0
    assert s1["last_fp64_ops"] > 20 * (p.n - 1) * p.S * p.R          # ~32 DMUL + DADD per update
    tl = eng.root_timeline(h)
    assert all(tl[k] is not None and 0 < tl[k] < 5000 for k in ("gather", "certificate", "final_terms")), tl
    quiet = Problem(packaged, 60, 300, 4, 3, "evolved", seed=32)
    hq, _ = quiet.create(eng)
    tq = eng.root_timeline(hq)
    assert tq["gather"] is not None and tq["final_terms"] is None       # nothing rescaled: the gather finished alone
    eng.finalDeallocate(h); eng.finalDeallocate(hq)
