"""Shared builders for the parity tests: seeded synthetic problems in the reference's input conventions."""
import numpy as np

from dmphyclus_b200 import synth


class Problem:
    def __init__(self, packaged, n, n_sites, k, n_rates, kind="iid", shape="random", seed=0, w_idx=5, b_idx=5):
        rng = np.random.default_rng(seed)
        self.n, self.S, self.R = n, n_sites, n_rates
        self.W = packaged["within"][:, :n_rates].copy()
        self.B = packaged["between"][:, :n_rates].copy()
        if kind == "diverged":
            # longer branches (powers of the packaged between-cluster matrices): sites then differ in which rate
            # category dominates, which is what makes the reference's cross-locus exponent carry visible
            #  and exercises the sequential fp32 chain
            def power(m, k):
                out = np.stack([[np.linalg.matrix_power(m[i, r], k) for r in range(n_rates)] for i in range(len(m))])
                return out / out.sum(-1, keepdims=True)
            self.W, self.B = power(self.B, 4), power(self.B, 8)
        self.pi = synth.LIM_PROBS.copy()
        self.edge = synth.random_tree(n, rng, shape)
        self.mrcas, self.clus = synth.pick_clusters(self.edge, n, k)
        if kind == "iid":
            self.chars = synth.iid_alignment(n, n_sites, rng)
        elif kind == "diverged":
            self.chars = synth.evolve_alignment(self.edge, n, n_sites, self.W[9], self.B[9], self.mrcas, self.pi, rng)
        elif kind == "evolved":
            self.chars = synth.evolve_alignment(self.edge, n, n_sites, self.W[4], self.B[9], self.mrcas, self.pi, rng)
        else:
            raise ValueError(kind)
        self.w_idx, self.b_idx = w_idx, b_idx

    def create(self, backend):
        al = backend.getConvertedAlignment(list(synth.STATES), self.chars)
        res = backend.logLikCpp(self.edge, self.mrcas, self.pi, self.W[self.w_idx - 1], self.B[self.b_idx - 1], 1, al,
                                self.n, self.S, self.w_idx, self.b_idx)
        return res["solutionPointer"], res["logLik"]


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)
