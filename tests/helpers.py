"""Shared helpers of the parity tests."""
import numpy as np
from dmphyclus_b200.workloads import Problem  # noqa: F401


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


class GoldenCase:
    """A case of tests/golden/ref_vectors.npz: the stored inputs in `Problem`'s shape plus the reference's outputs."""

    def __init__(self, data, name):
        import numpy as np
        from dmphyclus_b200 import synth
        g = {k.split("/", 1)[1]: data[k] for k in data.files if k.startswith(name + "/")}
        self.edge, self.mrcas, self.chars = g["in_edge"], [int(x) for x in g["in_mrcas"]], g["in_chars"]
        self.W, self.B = g["in_W"], g["in_B"]
        self.n, self.S = self.chars.shape
        self.R = self.W.shape[1]
        self.pi = synth.LIM_PROBS.copy()
        self.w_idx = self.b_idx = 5
        self.expected = {k: v for k, v in g.items() if not k.startswith("in_")}
        self.np = np


def load_golden():
    import os
    import numpy as np
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))


def compare_to_golden(got, expected, ll_tol, exact_partials=True):
    """Integer outputs bit-exact; log-likelihoods within ll_tol relative; partial vectors bit-exact (or 1e-13
    relative) and fp32 exponents to their last bit (glibc vs CUDA log differ by <= 1 ulp before the float rounding)."""
    import numpy as np
    for k, e in expected.items():
        g = got[k]
        if k == "logliks":
            assert g.shape == e.shape
            assert np.all(np.abs(g - e) <= ll_tol * np.abs(e)), (k, g, e)
        elif k.startswith("sol_"):
            if exact_partials:
                assert np.array_equal(g, e), k
            else:
                assert np.allclose(g, e, rtol=1e-13, atol=0), k
        elif k.startswith("exp_"):
            assert np.allclose(g, e, rtol=2e-7, atol=0), k
        else:
            assert np.array_equal(g, e), k


def recompute(be, h, W, B, w_idx, pi):
    """InvalidateAll + ComputeLoglik: the engine's recomputeAll, or the same thing through the reference's entry point."""
    if hasattr(be, "recomputeAll"):
        return be.recomputeAll(h, W, B, pi)
    return be.newWithinTransProbsLogLik(h, W, B, 1, w_idx, pi)


def scripted_moves(be, h, p, keep):
    """A short scripted sequence over every kind of proposal; returns the log-likelihoods."""
    W, B = p.W, p.B
    out = []
    none = np.zeros((1, 2), np.int32)
    out.append(recompute(be, h, W[4], B[4], 5, p.pi))
    out.append(be.newBetweenTransProbsLogLik(h, W[4], B[1], 1, 2, p.pi))
    if not keep:
        be.RestorePreviousConfig(h, none, 5, 5, False, p.mrcas, False)
    out.append(be.newWithinTransProbsLogLik(h, W[6], B[1 if keep else 4], 1, 7, p.pi))
    m = [x for x in p.mrcas if x > p.n][0]
    out.append(be.clusSplitMergeLogLik(h, W[6], B[1 if keep else 4], [m], 1, p.pi))
    be.RestorePreviousConfig(h, none, 7, 2 if keep else 5, False, p.mrcas, True)
    r = be.withinClusNNIlogLik(h, W[6], B[1 if keep else 4], m, 1, 1, p.pi)
    out.append(r["logLik"])
    out.append(recompute(be, h, W[6], B[1 if keep else 4], 7, p.pi))
    return out
