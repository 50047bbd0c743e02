"""Seeded synthetic workloads in the reference's input conventions: the shapes
`BASELINE.json.configs` names, built from `synth` and the reference's packaged transition matrices
(fixture `tests/golden/packaged_data.npz`, made by `tests/golden/make_packaged_data.py`).  Shared by the
parity tests, `__graft_entry__.smoke()` and `bench.py`."""
from __future__ import annotations

import os

import numpy as np

from . import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name -> (tips, sites, clusters, rate categories): BASELINE.json configs 2-5
CONFIGS = {
    "config2": (1000, 10000, 20, 3),      # logLikFromClusInd only, 1 GPU
    "config3": (10000, 30000, 100, 3),    # chain, site-sharded
    "config4": (5000, 5000, 50, 3),       # proposal-batch sweep
    "config5": (100000, 1500, 100, 3),    # wide-short, deep trees
    "config3_shard8": (10000, 3776, 100, 3),   # what ONE rank of config 3 over 8 GPUs holds (profiling a rank's kernels on one GPU)
}


def load_packaged() -> dict:
    d = np.load(os.path.join(ROOT, "tests", "golden", "packaged_data.npz"))
    return {k: d[k] for k in d.files}


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
        if kind == "tiny":
            # near-identity matrices with 1e-80 off-diagonals: three different bases under a vertex already push its
            # partial below 1e-150, so even the smallest clades rescale (exercises the exact mode of fused clades)
            def tiny(m, eps):
                out = np.empty_like(m)
                out[...] = eps
                idx = np.arange(4)
                out[..., idx, idx] = 1.0
                return out
            self.W = np.stack([tiny(self.W[i], 10.0 ** -(80 + i)) for i in range(len(self.W))])
            self.B = np.stack([tiny(self.B[i], 10.0 ** -(70 + i)) for i in range(len(self.B))])
        self.pi = synth.LIM_PROBS.copy()
        self.edge = synth.random_tree(n, rng, shape)
        self.mrcas, self.clus = synth.pick_clusters(self.edge, n, k)
        if kind in ("iid", "tiny"):
            self.chars = synth.iid_alignment(n, n_sites, rng)
        elif kind == "diverged":
            self.chars = synth.evolve_alignment(self.edge, n, n_sites, self.W[9], self.B[9], self.mrcas, self.pi, rng)
        elif kind == "matched":      # evolved under exactly the matrices it is evaluated with
            self.chars = synth.evolve_alignment(self.edge, n, n_sites, self.W[w_idx - 1], self.B[b_idx - 1], self.mrcas, self.pi, rng)
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
