"""Host-side mirror of the reference's MCMC driver (`/root/reference/R/DMphyClusSource.R`) over a
pluggable likelihood backend.

The reference's driver is R and stays R in a drop-in deployment; R is absent from this
image, so this module restates `.performStepPhylo` and the five Metropolis-Hastings moves with the same
move order, the same calls into the eleven entry points and the same RNG *draw order*, so that

  * "identical cluster-indicator trajectories for a fixed seed" is testable (CUDA engine vs oracle under
    the same numpy stream), and
  * `MCMC iterations/sec` can be measured.

R's own Mersenne-Twister stream (`sample.int`, `sample`, `runif`) is replaced by `numpy.random.Generator`
— the draw ORDER is the reference's, the VALUES are numpy's.  A backend is any object with the eleven
entry points under the reference's names (`dmphyclus_b200.engine.Engine`; the tests plug in the oracle).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .phylo import Phylo, cluster_mrcas


# ---------------------------------------------------------------- priors (R/DMphyClusInterface.R:147-162)
def _log_big_b(x) -> float:                       # .logBigB, DMphyClusSource.R:329-333
    x = [float(v) for v in x]
    return sum(math.lgamma(v) for v in x) - math.lgamma(sum(x))


def clusIndLogPrior(clusInd, alpha, k=None) -> float:
    """Dirichlet-multinomial log-prior of a cluster assignment vector (clusIndLogPrior, R API)."""
    counts = np.bincount(np.unique(np.asarray(clusInd), return_inverse=True)[1]).astype(np.float64)
    return _log_prior_of_counts(counts, len(np.asarray(clusInd)), alpha, k)


def _log_prior_of_counts(counts, n, alpha, k=None) -> float:
    """clusIndLogPrior from table(clusInd) (counts in label order) and length(clusInd):
    `lgamma(n+1) - sum(lgamma(counts+1)) + .logBigB(alpha+counts) - .logBigB(alpha)`.  Every sum is the builtin `sum` over
    the terms in cluster order, so the value does not depend on how the terms are produced (lists here: this runs five
    times per MCMC iteration over all clusters)."""
    lg = math.lgamma
    counts = [float(c) for c in counts]
    if k is None:
        k = len(counts)
    if len(counts) < k:
        counts = counts + [0.0] * (k - len(counts))
    if np.ndim(alpha) == 0:
        a = float(alpha)
        post = [a + c for c in counts]
        return (lg(n + 1) - sum([lg(c + 1) for c in counts])
                + (sum([lg(v) for v in post]) - lg(sum(post)))
                - (sum([lg(a)] * k) - lg(sum([a] * k))))
    alpha_vec = [float(a) for a in alpha]
    return (lg(n + 1) - sum([lg(c + 1) for c in counts])
            + _log_big_b([a + c for a, c in zip(alpha_vec, counts)]) - _log_big_b(alpha_vec))


def _dgamma_log(x, shape, scale) -> float:
    if x <= 0:
        return -math.inf if not (x == 0 and shape == 1) else -math.log(scale)
    return (shape - 1) * math.log(x) - x / scale - math.lgamma(shape) - shape * math.log(scale)


def _dpois_log(k, lam) -> float:
    return k * math.log(lam) - lam - math.lgamma(k + 1)


def _relabel(x):                                   # .relabel, DMphyClusSource.R:322-327
    x = np.asarray(x)
    _, inv = np.unique(x, return_inverse=True)
    return inv + 1


# ---------------------------------------------------------------- state
@dataclass
class ChainSettings:
    limProbs: np.ndarray
    withinTransMatAll: np.ndarray          # [n_within][R][4][4]
    betweenTransMatAll: np.ndarray         # [n_between][R][4][4]
    shapeForAlpha: float
    scaleForAlpha: float
    alphaMin: float = 0.0
    poisRateNumClus: float = 1.0
    numMovesNNIbetween: int = 1
    numMovesNNIwithin: int = 1
    numLikThreads: int = 1
    clusPhyloUpdateProp: float = 1.0
    numSplitMergeMoves: int = 1
    togglePhyloUpdates: bool = True
    maxMapSize: float | None = None
    cullProportion: float | None = None

    def __post_init__(self):
        # the lists the chain proposes from never change (DMphyClusInterface.R: withinTransMatAll / betweenTransMatAll),
        # so every list is one read-only array object: a backend may convert it once and recognise it afterwards
        def frozen(a):
            a = np.array(a, dtype=np.float64)
            a.setflags(write=False)
            return a
        self.limProbs = frozen(self.limProbs)
        self._within = [frozen(m) for m in self.withinTransMatAll]
        self._between = [frozen(m) for m in self.betweenTransMatAll]

    def within(self, index: int) -> np.ndarray:      # withinTransMatAll[[index]], 1-based
        return self._within[index - 1]

    def between(self, index: int) -> np.ndarray:
        return self._between[index - 1]


@dataclass
class ChainState:
    edge: np.ndarray                       # paraValues$phylogeny$edge
    n_tips: int
    clusInd: np.ndarray                    # paraValues$clusInd, in tip order
    alpha: float
    withinMatListIndex: int
    betweenMatListIndex: int
    clusterNodeIndices: list
    logLik: float = 0.0
    logPostProb: float = 0.0
    extPointer: object = None
    clusterCounts: dict = field(default_factory=dict)   # names(clusterCounts) -> count
    n_loglik_calls: int = 0
    _phylo: object = None                  # view of `edge`, rebuilt only when an accepted NNI replaced it
    _phylo_edge: object = None

    def phylo(self) -> Phylo:
        """paraValues$phylogeny as a parent/children view.  R hands the same phylo object to every move until an NNI
        is accepted; here the view is cached on the identity of the edge array for the same reason."""
        if self._phylo is None or self._phylo_edge is not self.edge:
            self._phylo, self._phylo_edge = Phylo(self.edge, self.n_tips), self.edge
        return self._phylo


def init_current_value(backend, alignmentBin, edge, clusInd, alpha, st: ChainSettings,
                       withinMatListIndex=None, betweenMatListIndex=None) -> ChainState:
    """.initCurrentValue (DMphyClusSource.R:54-98)."""
    n, s = alignmentBin.shape
    if betweenMatListIndex is None:
        betweenMatListIndex = math.ceil(len(st.betweenTransMatAll) / 2)
    if withinMatListIndex is None:
        withinMatListIndex = math.ceil(len(st.withinTransMatAll) / 2)
    phy = Phylo(edge, n)
    clusInd = np.asarray(clusInd, dtype=np.int64)
    mrcas = cluster_mrcas(phy, clusInd)
    res = backend.logLikCpp(edge, mrcas, st.limProbs, st.within(withinMatListIndex),
                            st.between(betweenMatListIndex), st.numLikThreads, alignmentBin, n, s,
                            withinMatListIndex, betweenMatListIndex)
    cv = ChainState(edge=_frozen_edge(edge), n_tips=n, clusInd=clusInd, alpha=float(alpha),
                    withinMatListIndex=withinMatListIndex, betweenMatListIndex=betweenMatListIndex,
                    clusterNodeIndices=list(mrcas), logLik=res["logLik"], extPointer=res["solutionPointer"])
    cv.clusterCounts = {k: int((clusInd == k).sum()) for k in range(1, int(clusInd.max()) + 1)}
    cv.logPostProb = (cv.logLik + clusIndLogPrior(clusInd, cv.alpha)
                      + _dgamma_log(cv.alpha - st.alphaMin, st.shapeForAlpha, st.scaleForAlpha)
                      + _dpois_log(int(clusInd.max()), st.poisRateNumClus))
    cv.n_loglik_calls = 1
    return cv


# ---------------------------------------------------------------- moves
def _update_trans_matrix(backend, cv: ChainState, st: ChainSettings, rng, between: bool):
    """.updateTransMatrix (DMphyClusSource.R:216-255)."""
    all_list = st._between if between else st._within
    new_index = int(rng.integers(1, len(all_list) + 1))           # sample.int(n, 1)
    if between:
        new_ll = backend.newBetweenTransProbsLogLik(cv.extPointer, st.within(cv.withinMatListIndex),
                                                    all_list[new_index - 1], st.numLikThreads, new_index, st.limProbs)
    else:
        new_ll = backend.newWithinTransProbsLogLik(cv.extPointer, all_list[new_index - 1],
                                                   st.between(cv.betweenMatListIndex),
                                                   st.numLikThreads, new_index, st.limProbs)
    cv.n_loglik_calls += 1
    if rng.random() < _exp(new_ll - cv.logLik):
        if between:
            cv.betweenMatListIndex = new_index
        else:
            cv.withinMatListIndex = new_index
        cv.logPostProb = new_ll + (cv.logPostProb - cv.logLik)
        cv.logLik = new_ll
    else:
        backend.RestorePreviousConfig(cv.extPointer, _NO_EDGE, cv.withinMatListIndex,
                                      cv.betweenMatListIndex, False, cv.clusterNodeIndices, False)


_NO_EDGE = np.zeros((1, 2), np.int32)            # matrix(0, 1, 1): what the R code passes when no NNI is rolled back
_NO_EDGE.setflags(write=False)


def _frozen_edge(edge) -> np.ndarray:
    """paraValues$phylogeny$edge: replaced as a whole when an NNI is accepted, never edited in place — kept read-only so
    that a backend (and the cached tree view) may recognise the same matrix when it comes back on a rollback."""
    e = np.array(edge, dtype=np.int32)
    e.setflags(write=False)
    return e


def _exp(x: float) -> float:
    return math.inf if x > 700 else math.exp(x)


def _nni_update(backend, cv, st, rng, call):
    res = call()
    cv.n_loglik_calls += 1
    new_ll = res["logLik"]
    if rng.random() < _exp(new_ll - cv.logLik):
        cv.logPostProb = cv.logPostProb - cv.logLik + new_ll
        cv.logLik = new_ll
        cv.edge = _frozen_edge(res["edge"])
    else:
        backend.RestorePreviousConfig(cv.extPointer, cv.edge, cv.withinMatListIndex, cv.betweenMatListIndex, True,
                                      cv.clusterNodeIndices, False)


def _update_between_phylo(backend, cv, st, rng):
    """.updateBetweenPhylo (DMphyClusSource.R:335-351)."""
    w = st.within(cv.withinMatListIndex)
    b = st.between(cv.betweenMatListIndex)
    _nni_update(backend, cv, st, rng, lambda: backend.betweenClusNNIlogLik(
        cv.extPointer, w, b, cv.clusterNodeIndices, st.numMovesNNIbetween, st.numLikThreads, st.limProbs))


def _update_cluster_phylos(backend, cv, st, rng):
    """.updateClusterPhylos (DMphyClusSource.R:371-415)."""
    if st.clusPhyloUpdateProp < 1:
        k = math.ceil(len(cv.clusterNodeIndices) * st.clusPhyloUpdateProp)
        select = [cv.clusterNodeIndices[i] for i in rng.choice(len(cv.clusterNodeIndices), size=k, replace=False)]
    else:
        select = list(cv.clusterNodeIndices)
    # length(Descendants(originalPhylo, clusterMRCA)) = the size of that cluster: clusters are clades of the tree
    # (DMphyClusInterface.R:228-249) and NNI moves keep every cluster's tip set and MRCA id (DMphyClusSource.R:396)
    size_of = {m: cv.clusterCounts[k + 1] for k, m in enumerate(cv.clusterNodeIndices)}
    for mrca in select:
        if size_of[mrca] < 3:
            continue
        w = st.within(cv.withinMatListIndex)
        b = st.between(cv.betweenMatListIndex)
        _nni_update(backend, cv, st, rng, lambda: backend.withinClusNNIlogLik(
            cv.extPointer, w, b, mrca, st.numMovesNNIwithin, st.numLikThreads, st.limProbs))


def _parents(phy: Phylo, mrcas) -> np.ndarray:
    return phy.parent[np.asarray(mrcas, dtype=np.int64)]          # one vector look-up instead of one access per cluster


def _pairs_and_splits(phy: Phylo, mrcas, counts_gt1):
    """Sibling cluster pairs (mergeable) grouped by parent in increasing parent id, as
    split(clusterNodeIndices, f = parent) orders its factor levels (DMphyClusSource.R:157-165); within a group the
    clusters keep their list order."""
    par = _parents(phy, mrcas)
    order = np.argsort(par, kind="stable")
    sp = par[order]
    same = np.nonzero(sp[1:] == sp[:-1])[0]                        # positions whose successor has the same parent
    pairs, order = [], order.tolist()
    i = 0
    same = same.tolist()
    while i < len(same):
        j = i
        while j + 1 < len(same) and same[j + 1] == same[j] + 1:    # (more than two clusters under one parent: not in a binary tree)
            j += 1
        pairs.append([mrcas[q] for q in order[same[i]:same[j] + 2]])
        i = j + 1
    return pairs, counts_gt1


def _num_pairs(phy: Phylo, mrcas) -> int:
    """How many parents have more than one cluster MRCA hanging off them."""
    return int(np.count_nonzero(np.unique(_parents(phy, mrcas), return_counts=True)[1] > 1))


def _split_join_cluster_move(backend, cv: ChainState, st: ChainSettings, rng):
    """.splitJoinClusterMove / .performSplit / .performMerge (DMphyClusSource.R:100-214)."""
    phy = cv.phylo()
    pairs = _pairs_and_splits(phy, cv.clusterNodeIndices, None)[0] if len(cv.clusterNodeIndices) > 1 else []
    to_split = [k for k, c in cv.clusterCounts.items() if c > 1]
    num_moves = len(pairs) + len(to_split)
    if num_moves == 0:
        raise RuntimeError("No possible move. This is an error. Exiting...")
    selected = int(rng.integers(1, num_moves + 1))                # sample(1:numMoves, 1)
    split_move = selected > len(pairs)
    w = st.within(cv.withinMatListIndex)
    b = st.between(cv.betweenMatListIndex)
    if split_move:
        clus_number = to_split[selected - len(pairs) - 1]
        node = cv.clusterNodeIndices[clus_number - 1]
        if node <= cv.n_tips:
            raise RuntimeError("split of a cluster whose stored MRCA is a tip (single-cluster start): "
                               "the reference dereferences a null child here (AugTree.cpp:332)")
        # Children(phylogeny, node) and the tips below each of them: the cluster IS the clade below its MRCA
        # (DMphyClusInterface.R:228-249; every move keeps it so), hence its tips are the ones to sort by child
        new_nodes, parts = phy.split_below(node, np.flatnonzero(cv.clusInd == clus_number) + 1)
        new_mrcas = list(cv.clusterNodeIndices)
        new_mrcas[clus_number - 1] = new_nodes[0]
        new_mrcas += new_nodes[1:]
        new_ll = backend.clusSplitMergeLogLik(cv.extPointer, w, b, [node], st.numLikThreads, st.limProbs)
        numbers = [clus_number] + list(range(len(cv.clusterNodeIndices) + 1, len(new_mrcas) + 1))
        # table(newClusInd) without building newClusInd: the split cluster's tips are exactly the tips below its MRCA's
        # children, and the labels in use are 1..K (merges relabel), so K+1 is new and the keys stay sorted
        new_counts = dict(cv.clusterCounts)
        for num, tips in zip(numbers, parts):
            new_counts[num] = len(tips)

        def build_clus():
            out = cv.clusInd.copy()
            for num, tips in zip(numbers, parts):
                out[tips - 1] = num
            return out
    else:
        to_merge = pairs[selected - 1]
        new_mrca = int(phy.parent[to_merge[0]])                   # phangorn::mrca.phylo of siblings
        numbers = [cv.clusterNodeIndices.index(m) + 1 for m in to_merge]
        replaced = list(cv.clusterNodeIndices)
        for k in numbers:
            replaced[k - 1] = new_mrca
        new_mrcas = list(dict.fromkeys(replaced))                 # unique(), first occurrence kept
        new_ll = backend.clusSplitMergeLogLik(cv.extPointer, w, b, to_merge, st.numLikThreads, st.limProbs)
        lowest = min(numbers)                                     # the merged cluster takes the lowest label, leaving a gap
        new_counts = {k: c for k, c in cv.clusterCounts.items() if k == lowest or k not in numbers}
        new_counts[lowest] = sum(cv.clusterCounts[k] for k in numbers)

        def build_clus():
            out = cv.clusInd.copy()
            out[np.isin(out, numbers)] = lowest
            return out
    cv.n_loglik_calls += 1
    new_lpp = (new_ll + _log_prior_of_counts(list(new_counts.values()), len(cv.clusInd), cv.alpha)
               + _dgamma_log(cv.alpha - st.alphaMin, st.shapeForAlpha, st.scaleForAlpha)
               + _dpois_log(max(new_counts), st.poisRateNumClus))
    num_pairs_new = _num_pairs(phy, new_mrcas)
    num_splits_new = sum(1 for c in new_counts.values() if c > 1)
    ratio = num_moves / (num_pairs_new + num_splits_new)
    mh = ratio * _exp(new_lpp - cv.logPostProb)
    if rng.random() < mh:
        cv.logLik, cv.logPostProb = new_ll, new_lpp
        cv.clusInd, cv.clusterCounts, cv.clusterNodeIndices = build_clus(), new_counts, new_mrcas   # newClusInd, only if kept
        if not split_move:
            keys = _relabel(np.array(sorted(cv.clusterCounts)))
            cv.clusterCounts = {int(k): cv.clusterCounts[o] for k, o in zip(keys, sorted(cv.clusterCounts))}
            cv.clusInd = _relabel(cv.clusInd)
    else:
        backend.RestorePreviousConfig(cv.extPointer, _NO_EDGE, cv.withinMatListIndex,
                                      cv.betweenMatListIndex, False, cv.clusterNodeIndices, True)


def _update_alpha(cv: ChainState, st: ChainSettings, rng):
    """.updateAlpha (DMphyClusSource.R:38-52)."""
    new_alpha = cv.alpha + (rng.random() - 0.5)
    new_alpha = new_alpha + abs(new_alpha - st.alphaMin) * (new_alpha < st.alphaMin)
    counts = [cv.clusterCounts[k] for k in sorted(cv.clusterCounts)]          # table(clusInd)
    new_lp = _log_prior_of_counts(counts, len(cv.clusInd), new_alpha) + _dgamma_log(new_alpha - st.alphaMin, st.shapeForAlpha, st.scaleForAlpha)
    cur_lp = _log_prior_of_counts(counts, len(cv.clusInd), cv.alpha) + _dgamma_log(cv.alpha - st.alphaMin, st.shapeForAlpha, st.scaleForAlpha)
    pois = cv.logPostProb - cur_lp - cv.logLik
    if rng.random() < _exp(new_lp - cur_lp):
        cv.alpha = new_alpha
        cv.logPostProb = cv.logLik + new_lp + pois


def perform_step_phylo(backend, cv: ChainState, st: ChainSettings, rng) -> ChainState:
    """One MCMC iteration: .performStepPhylo (DMphyClusSource.R:1-26)."""
    if cv.clusInd.max() > 1:
        _update_trans_matrix(backend, cv, st, rng, between=True)
    _update_trans_matrix(backend, cv, st, rng, between=False)
    if st.togglePhyloUpdates:
        if cv.clusInd.max() > 2:
            _update_between_phylo(backend, cv, st, rng)
        _update_cluster_phylos(backend, cv, st, rng)
    for _ in range(st.numSplitMergeMoves):
        _split_join_cluster_move(backend, cv, st, rng)
    _update_alpha(cv, st, rng)
    if st.maxMapSize is not None:
        backend.checkAndCullMap(cv.extPointer, st.maxMapSize, st.cullProportion,
                                st.within(cv.withinMatListIndex),
                                st.between(cv.betweenMatListIndex), st.limProbs)
    return cv


def run_chain(backend, alignmentBin, edge, clusInd, alpha, st: ChainSettings, n_iter: int, seed: int = 0,
              withinMatListIndex=None, betweenMatListIndex=None, on_iter=None):
    """.DMphyClusCore (DMphyClusSource.R:257-313): returns the per-iteration samples."""
    rng = np.random.default_rng(seed)
    cv = init_current_value(backend, alignmentBin, edge, clusInd, alpha, st, withinMatListIndex, betweenMatListIndex)
    out = []
    try:
        for it in range(n_iter):
            perform_step_phylo(backend, cv, st, rng)
            out.append({"clusInd": cv.clusInd.copy(), "edge": cv.edge.copy(), "alpha": cv.alpha,
                        "withinMatListIndex": cv.withinMatListIndex, "betweenMatListIndex": cv.betweenMatListIndex,
                        "clusterNodeIndices": list(cv.clusterNodeIndices), "logPostProb": cv.logPostProb,
                        "logLik": cv.logLik})
            if on_iter is not None:
                on_iter(it, cv)
    finally:
        backend.finalDeallocate(cv.extPointer)
    return out, cv.n_loglik_calls
