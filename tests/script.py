"""A fixed sequence of calls over the eleven entry points (every proposal type, accepted and rejected) that any
backend — the reference's own TUs, the CPU oracle, the CUDA engine — can replay; the recorded outputs are what
the golden fixtures hold (tests/golden/make_ref_vectors.py) and what the parity tests compare."""
import numpy as np

from dmphyclus_b200.phylo import Phylo

NO_EDGE = np.zeros((1, 2), np.int32)


def run_script(be, p, partials=True):
    out = {}
    pi, W, B = p.pi, p.W, p.B
    n = p.n
    al = be.getConvertedAlignment(list("atcg"), p.chars)
    out["masks"] = np.asarray(al).copy()
    w, b = p.w_idx, p.b_idx
    res = be.logLikCpp(p.edge, p.mrcas, pi, W[w - 1], B[b - 1], 1, al, n, p.S, w, b)
    h = res["solutionPointer"]
    lls = [res["logLik"]]
    edge = np.array(p.edge, np.int32)
    mrcas = list(p.mrcas)
    if partials:
        for tag, v in (("root", n), ("mid", n + (n - 1) // 2), ("last", 2 * n - 2)):
            s, e = be.partial(h, v)
            out[f"sol_{tag}"], out[f"exp_{tag}"] = s.copy(), e.copy()
    phy = Phylo(edge, n)
    k = len(mrcas)
    # 1. between-cluster matrices, rejected
    if k > 1:
        lls.append(be.newBetweenTransProbsLogLik(h, W[w - 1], B[1], 1, 2, pi))
        be.RestorePreviousConfig(h, NO_EDGE, w, b, False, mrcas, False)
    # 2. within-cluster matrices (InvalidateAll), accepted
    lls.append(be.newWithinTransProbsLogLik(h, W[6], B[b - 1], 1, 7, pi))
    w = 7
    big = [m for m in mrcas if m > n and len(phy.descendants(m)) >= 3]
    # 3. within-cluster NNI, rejected
    if big:
        r = be.withinClusNNIlogLik(h, W[w - 1], B[b - 1], big[0], 1, 1, pi)
        lls.append(r["logLik"])
        out["edge_nni_rejected"] = np.array(r["edge"], np.int32)
        be.RestorePreviousConfig(h, edge, w, b, True, mrcas, False)
    # 4. between-cluster NNI, accepted
    if k > 2:
        r = be.betweenClusNNIlogLik(h, W[w - 1], B[b - 1], mrcas, 1, 1, pi)
        lls.append(r["logLik"])
        edge = np.array(r["edge"], np.int32)
        phy = Phylo(edge, n)
    # 5. within-cluster NNI with two moves, accepted
    if big:
        r = be.withinClusNNIlogLik(h, W[w - 1], B[b - 1], big[-1], 2, 1, pi)
        lls.append(r["logLik"])
        edge = np.array(r["edge"], np.int32)
        phy = Phylo(edge, n)
    splittable = [m for m in mrcas if m > n]
    if splittable:
        m = splittable[0]
        # 6. split, rejected
        lls.append(be.clusSplitMergeLogLik(h, W[w - 1], B[b - 1], [m], 1, pi))
        be.RestorePreviousConfig(h, NO_EDGE, w, b, False, mrcas, True)
        # 7. split, accepted; 8. merge of the two halves, accepted
        m = splittable[-1]
        lls.append(be.clusSplitMergeLogLik(h, W[w - 1], B[b - 1], [m], 1, pi))
        kids = phy.children[m]
        split_mrcas = [x for x in mrcas if x != m] + kids
        lls.append(be.newBetweenTransProbsLogLik(h, W[w - 1], B[8], 1, 9, pi))
        b = 9
        lls.append(be.clusSplitMergeLogLik(h, W[w - 1], B[b - 1], kids, 1, pi))
        del split_mrcas
    # 9. a final full re-evaluation through new within-cluster matrices: checks the whole committed state
    lls.append(be.newWithinTransProbsLogLik(h, W[2], B[b - 1], 1, 3, pi))
    out["logliks"] = np.array(lls)
    out["edge_final"] = edge
    topo = be.topology(h)
    for kk in ("parent", "child0", "child1", "within"):
        out[f"topo_{kk}"] = np.asarray(topo[kk]).copy()
    if partials:
        s, e = be.partial(h, n)
        out["sol_root_final"], out["exp_root_final"] = s.copy(), e.copy()
    be.finalDeallocate(h)
    return out


# (name, tips, sites, clusters, rate categories, alignment kind, tree shape, seed)
GOLDEN_CASES = [
    ("tiny", 8, 60, 3, 3, "evolved", "random", 101),
    ("small", 40, 64, 5, 3, "evolved", "random", 102),
    ("balanced_r1", 64, 48, 6, 1, "iid", "balanced", 103),
    ("caterpillar_r2", 300, 24, 5, 2, "iid", "caterpillar", 104),
    ("rescale_r3", 700, 40, 8, 3, "iid", "random", 105),
    ("carry_r3", 1024, 30, 6, 3, "diverged", "random", 106),
]

VIGNETTE_EDGE = np.array([[7, 8], [8, 9], [9, 1], [9, 2], [8, 3], [7, 10], [10, 11], [11, 4], [11, 5], [10, 6]])
# (cluster MRCAs, within index, between index)
VIGNETTE_CASES = [([8, 10], 5, 5), ([8, 10], 1, 1), ([8, 10], 10, 10), ([8, 10], 5, 1), ([7], 5, 5),
                  ([1, 2, 3, 4, 5, 6], 5, 5)]


def vignette_loglik(be, packaged, mrcas, wi, bi):
    from dmphyclus_b200 import synth
    al = be.getConvertedAlignment(list("atcg"), packaged["seqs"])
    r = be.logLikCpp(VIGNETTE_EDGE, mrcas, synth.LIM_PROBS, packaged["within"][wi - 1], packaged["between"][bi - 1],
                     1, al, 6, 1617, wi, bi)
    be.finalDeallocate(r["solutionPointer"])
    return r["logLik"]
