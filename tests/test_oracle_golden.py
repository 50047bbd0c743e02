"""The CPU oracle (oracle/dmpc_oracle.cpp) against the committed golden vectors that the reference's own
translation units produced (tests/golden/make_ref_vectors.py).  Runs anywhere; no GPU."""
import numpy as np
import pytest

from helpers import GoldenCase, compare_to_golden, load_golden
from script import GOLDEN_CASES, VIGNETTE_CASES, run_script, vignette_loglik


@pytest.fixture(scope="module")
def golden():
    return load_golden()


@pytest.mark.parametrize("name", [c[0] for c in GOLDEN_CASES])
def test_oracle_replays_reference_script(oracle, golden, name):
    case = GoldenCase(golden, name)
    got = run_script(oracle, case)
    # same compiler, same libm, same operation order: the oracle reproduces the reference bit for bit
    compare_to_golden(got, case.expected, ll_tol=0.0, exact_partials=True)
    assert np.array_equal(got["exp_root"], case.expected["exp_root"])


def test_oracle_vignette_known_answers(oracle, packaged, golden):
    """Config 1 data: the reference TUs' log-likelihoods == BASELINE.md §5's independently derived values."""
    derived = [-3072.967109235698, -3148.266004654241, -3038.497338364486, -3130.306408470979, -3510.074064616059,
               -3259.751644946885]
    for (m, wi, bi), ref_ll, d in zip(VIGNETTE_CASES, golden["vignette/logliks"], derived):
        got = vignette_loglik(oracle, packaged, m, wi, bi)
        assert got == ref_ll
        assert abs(got - d) < 1e-10 * abs(d)


def test_taus_generator(oracle, golden):
    """gsl_rng_taus: GSL's own test-suite value (rng/test.c: seed 1, 10000th output) and the default-seed stream."""
    assert int(oracle.taus_stream(1, 10000)[-1]) == 2733957125
    assert np.array_equal(oracle.taus_stream(0, 64), golden["taus/seed0"])
    assert np.array_equal(oracle.taus_stream(0, 8), oracle.taus_stream(1, 8))     # gsl: seed 0 means seed 1


def test_textbook_switch_differs_only_with_carry(oracle, packaged):
    """quirk_carry=False resets the exponent vector per locus: identical when nothing rescales (small trees),
    different on rescale-heavy multi-rate data."""
    from dmphyclus_b200.workloads import Problem
    from oracle.oracle import Oracle
    tb = Oracle(quirk_carry=False)
    p = Problem(packaged, 40, 64, 5, 3, "evolved", seed=1)
    (h1, l1), (h2, l2) = p.create(oracle), p.create(tb)
    assert l1 == l2
    oracle.finalDeallocate(h1); tb.finalDeallocate(h2)
    p = Problem(packaged, 1024, 60, 6, 3, "diverged", seed=11)
    (h1, l1), (h2, l2) = p.create(oracle), p.create(tb)
    assert abs(l1 - l2) > 1e-6 * abs(l1)
    oracle.finalDeallocate(h1); tb.finalDeallocate(h2)
