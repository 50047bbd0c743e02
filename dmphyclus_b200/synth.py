"""Seeded synthetic inputs for the likelihood path: trees as `phylo$edge` matrices,
cluster MRCAs as disjoint clades, and alignments either evolved down the tree with the
within/between transition matrices or drawn iid (rescale stress).  numpy only — no R, no reference.
"""
from __future__ import annotations

import numpy as np

from .phylo import Phylo

STATES = "atcg"                      # names(limProbs) order of the packaged data
LIM_PROBS = np.array([0.39, 0.22, 0.17, 0.22])   # vignette: limProbsPosada2001


def _relabel_preorder(left, right, root, n):
    """temp ids (tips 0..n-1, internal n..2n-2) -> ape numbering (tips 1..n, root n+1, internal nodes
    numbered in pre-order) and the pre-order edge list `AugTree::BuildEdgeMatrix` would emit."""
    new_id = np.zeros(2 * n - 1, np.int64)
    new_id[:n] = np.arange(1, n + 1)
    edge = np.zeros((2 * n - 2, 2), np.int64)
    nxt, row = n + 1, 0
    stack = [(root, 0)]
    while stack:
        v, par = stack.pop()
        if v >= n:
            new_id[v] = nxt
            nxt += 1
        if par:
            edge[row] = (par, new_id[v])
            row += 1
        if v >= n:
            stack.append((right[v - n], new_id[v]))
            stack.append((left[v - n], new_id[v]))
    return edge.astype(np.int32)


def random_tree(n: int, rng: np.random.Generator, shape: str = "random") -> np.ndarray:
    """Binary rooted tree with n tips as an edge matrix [2n-2][2] (1-based, root = n+1, pre-order rows).

    shape: "random" (uniformly random pairwise joining — coalescent-like), "balanced", or
    "caterpillar" (maximum depth; level-scheduling stress)."""
    assert n >= 2
    left = np.zeros(n - 1, np.int64)
    right = np.zeros(n - 1, np.int64)
    if shape == "random":
        active = list(range(n))
        for k in range(n - 1):
            m = len(active)
            i = int(rng.integers(m))
            j = int(rng.integers(m - 1))
            if j >= i:
                j += 1
            left[k], right[k] = active[i], active[j]
            active[i] = n + k
            active[j] = active[-1]
            active.pop()
        root = active[0]
    elif shape == "balanced":
        level = list(rng.permutation(n))
        k = 0
        while len(level) > 1:
            nxt = []
            for a in range(0, len(level) - 1, 2):
                left[k], right[k] = level[a], level[a + 1]
                nxt.append(n + k)
                k += 1
            if len(level) % 2:
                nxt.append(level[-1])
            level = nxt
        root = level[0]
    elif shape == "caterpillar":
        order = rng.permutation(n)
        cur = int(order[0])
        for k in range(n - 1):
            left[k], right[k] = cur, int(order[k + 1])
            cur = n + k
        root = cur
    else:
        raise ValueError(shape)
    return _relabel_preorder(left, right, root, n)


def pick_clusters(edge, n: int, k: int):
    """K disjoint clades chosen top-down (always opening the largest remaining clade), singletons allowed.
    Returns (clusterMRCAs [K] 1-based node ids, tip ids for singleton clusters; clusInd [n] in 1..K)."""
    phy = Phylo(edge, n)
    size = phy.subtree_sizes()
    frontier = [phy.root]
    while len(frontier) < k:
        cand = [v for v in frontier if v > n]
        if not cand:
            break
        v = max(cand, key=lambda x: (size[x], -x))
        frontier.remove(v)
        frontier.extend(phy.children[v])
    clus = np.zeros(n, np.int64)
    for i, v in enumerate(frontier):
        clus[phy.descendants(v) - 1] = i + 1
    return [int(v) for v in frontier], clus


def within_flags(phy: Phylo, mrcas) -> np.ndarray:
    """`_withinParentBranch` per node id (AugTree::AssociateTransProbMatrices, AugTree.cpp:30-57)."""
    w = np.zeros(len(phy.parent), bool)
    for m in mrcas:
        if m > phy.n_tips:
            stack = list(phy.children[m])
            while stack:
                v = stack.pop()
                w[v] = True
                if v > phy.n_tips:
                    stack.extend(phy.children[v])
    return w


def evolve_alignment(edge, n: int, n_sites: int, within, between, mrcas, lim_probs, rng) -> np.ndarray:
    """Alignment [n][n_sites] of lower-case ASCII bases evolved down the tree: root state ~ lim_probs, one
    rate category per site ~ U{0..R-1}, each branch draws the child state from row P_r[parent, :] of the
    within-cluster matrices below a cluster MRCA and the between-cluster ones above."""
    phy = Phylo(edge, n)
    within, between = np.asarray(within), np.asarray(between)
    n_rates = within.shape[0]
    w = within_flags(phy, mrcas)
    thr = {True: np.cumsum(within, axis=2).reshape(n_rates * 4, 4)[:, :3].copy(),
           False: np.cumsum(between, axis=2).reshape(n_rates * 4, 4)[:, :3].copy()}
    rate = rng.integers(n_rates, size=n_sites)
    state = {phy.root: rng.choice(4, size=n_sites, p=np.asarray(lim_probs) / np.sum(lim_probs)).astype(np.int64)}
    out = np.zeros((n, n_sites), np.uint8)
    code = np.frombuffer(STATES.encode(), np.uint8)
    stack = [phy.root]
    while stack:
        v = stack.pop()
        sv = state.pop(v)
        for c in phy.children[v]:
            t = thr[bool(w[c])][rate * 4 + sv]
            u = rng.random(n_sites)
            sc = (u[:, None] > t).sum(axis=1)
            if c <= n:
                out[c - 1] = code[sc]
            else:
                state[c] = sc
                stack.append(c)
    return out


def iid_alignment(n: int, n_sites: int, rng, p_missing: float = 0.01, p_ambiguous: float = 0.005) -> np.ndarray:
    """Uniform iid bases with a sprinkle of `n` and 2-way IUPAC codes: no phylogenetic signal, so the
    partial likelihoods shrink as fast as they can (rescale stress) and no two site columns agree."""
    code = np.frombuffer(STATES.encode(), np.uint8)
    out = code[rng.integers(4, size=(n, n_sites))]
    u = rng.random((n, n_sites))
    out[u < p_missing] = ord("n")
    amb = np.frombuffer(b"ryswkm", np.uint8)
    sel = (u >= p_missing) & (u < p_missing + p_ambiguous)
    out[sel] = amb[rng.integers(len(amb), size=int(sel.sum()))]
    return out
