"""Golden vectors of the likelihood path, produced by the REFERENCE'S OWN translation units
(`/root/reference/src/{clusterFunArmadillo,AugTree,IntermediateNode,InputNode}.cpp`, compiled unmodified against
the header shim in `oracle/shim` -> `oracle/_ref/libref.so`).  Run here, where /root/reference exists:

    python tests/golden/make_ref_vectors.py

Writes `tests/golden/ref_vectors.npz`: for every case of `tests/script.py:GOLDEN_CASES` the INPUTS (edge matrix,
cluster MRCAs, alignment characters — so the fixture does not depend on numpy's generator staying stable) and the
reference's outputs for the scripted call sequence (log-likelihoods of every proposal, partial-likelihood vectors
and fp32 exponents of three vertices, edge matrices after NNI moves, final topology and within-flags); plus the
six vignette-data log-likelihoods (config 1) and the first 64 draws of the NNI generator (gsl_rng_taus, default
seed) — the latter from the shim's restatement, itself pinned by GSL's own test-suite value
(rng/test.c: taus, seed 1, 10000th output = 2733957125).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from dmphyclus_b200.workloads import Problem, load_packaged  # noqa: E402
from oracle.oracle import Oracle, Reference  # noqa: E402
from script import GOLDEN_CASES, VIGNETTE_CASES, run_script, vignette_loglik  # noqa: E402


def main():
    ref = Reference()
    packaged = load_packaged()
    out = {}
    for name, n, s, k, r, kind, shape, seed in GOLDEN_CASES:
        p = Problem(packaged, n, s, k, r, kind, shape, seed=seed)
        res = run_script(ref, p)
        out[f"{name}/in_edge"] = p.edge
        out[f"{name}/in_mrcas"] = np.array(p.mrcas)
        out[f"{name}/in_chars"] = p.chars
        out[f"{name}/in_W"] = p.W
        out[f"{name}/in_B"] = p.B
        for key, val in res.items():
            out[f"{name}/{key}"] = val
        print(name, res["logliks"])
    out["vignette/logliks"] = np.array([vignette_loglik(ref, packaged, m, wi, bi) for m, wi, bi in VIGNETTE_CASES])
    print("vignette", out["vignette/logliks"])
    out["taus/seed0"] = Oracle().taus_stream(0, 64)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_vectors.npz"), os.path.getsize(os.path.join(HERE, "ref_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
