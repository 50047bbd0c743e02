"""Host logic of the engine on CPU: the plans its scheduler hands out (dmpc_hp_*, no device involved), executed by the
numpy interpreter of tests/plan_interpreter.py, must reproduce the oracle's partial vectors BIT FOR BIT after every step of
a scripted proposal sequence — full evaluations, new matrices, NNI moves (same Tausworthe picks, same edge matrices),
splits and merges, accepted and rejected — for every fusion threshold and tree shape, including walked chains with
materialised clades and rescaling inside fused clades."""
import numpy as np
import pytest

from dmphyclus_b200.phylo import Phylo
from dmphyclus_b200.workloads import Problem
from plan_interpreter import PlanMachine

NO_EDGE = np.zeros((1, 2), np.int32)


@pytest.fixture(scope="module")
def lib():
    from dmphyclus_b200 import build
    from dmphyclus_b200.engine import load_library
    build.build()
    return load_library()


def _compare(machine, oracle, ho, n, what, every=1):
    cur, within, edge = machine.state()
    topo = oracle.topology(ho)
    assert np.array_equal(within.astype(np.int32), np.asarray(topo["within"], np.int32)), what
    for v in range(n + 1, 2 * n, every):
        (sg, eg), (so, eo) = machine.partial(v), oracle.partial(ho, v - 1)
        # (NaN == NaN here: a partial that underflows to all zeros is divided by its maximum, 0, in the reference too)
        assert np.array_equal(sg, so, equal_nan=True), f"{what}: partial of vertex {v}"
        assert np.allclose(eg, eo, rtol=2e-7, atol=0, equal_nan=True), f"{what}: exponents of vertex {v}"
    return edge


@pytest.mark.parametrize("n,S,k,R,kind,shape,fuse,seed", [
    (9, 40, 2, 3, "evolved", "random", 15, 1), (40, 48, 5, 3, "evolved", "random", 4, 2), (64, 33, 6, 1, "iid", "balanced", 8, 3),
    (120, 24, 5, 2, "iid", "caterpillar", 15, 4), (120, 24, 5, 2, "iid", "caterpillar", 1, 5), (300, 20, 8, 3, "iid", "random", 15, 6),
    (300, 20, 8, 3, "diverged", "random", 2, 7), (400, 16, 6, 3, "tiny", "random", 8, 8), (500, 12, 24, 3, "evolved", "random", 15, 9)])
def test_interpreted_plans_reproduce_the_oracle(lib, oracle, packaged, n, S, k, R, kind, shape, fuse, seed):
    p = Problem(packaged, n, S, k, R, kind, shape, seed=seed)
    W, B, pi = p.W, p.B, p.pi
    masks = oracle.getConvertedAlignment(list("atcg"), p.chars)
    ho = oracle.logLikCpp(p.edge, p.mrcas, pi, W[4], B[4], 1, masks, n, S, 5, 5)["solutionPointer"]
    mc = PlanMachine(lib, p, masks, fuse)
    every = 1 if n <= 130 else 7
    try:
        mc.set_matrices(W[4], B[4])
        st = mc.evaluate()                                         # logLikCpp
        assert st["tasks"] > 0 and (fuse > 1 or st["tokens"] == 0)
        edge = _compare(mc, oracle, ho, n, "create", every)
        assert np.array_equal(edge, np.asarray(p.edge, np.int32))          # BuildEdgeMatrix order = the synthetic tree's own
        mrcas, w, b = list(p.mrcas), 5, 5
        phy = Phylo(edge, n)
        chains = mats = 0

        def step(what, kind_, args, call, keep, nni=False, split_merge=False, new_w=None, new_b=None):
            nonlocal edge, w, b, chains, mats
            mc.propose(kind_, args)
            mc.set_matrices(W[(new_w or w) - 1], B[(new_b or b) - 1])
            res = call()
            st = mc.evaluate()
            chains += st["chain"] > 0
            mats += st["materialised"]
            new_edge = _compare(mc, oracle, ho, n, what, every)
            if nni:
                assert np.array_equal(new_edge, np.asarray(res["edge"], np.int32)), what
            if keep:
                edge, w, b = new_edge, new_w or w, new_b or b
            else:
                oracle.RestorePreviousConfig(ho, edge if nni else NO_EDGE, w, b, nni, mrcas, split_merge)
                mc.restore(edge if nni else NO_EDGE, nni, mrcas, split_merge)
                mc.set_matrices(W[w - 1], B[b - 1])
                _compare(mc, oracle, ho, n, what + " (rolled back)", every)

        if len(mrcas) > 1:
            step("between matrices", 1, (), lambda: oracle.newBetweenTransProbsLogLik(ho, W[w - 1], B[1], 1, 2, pi), False, new_b=2)
        step("within matrices", 0, (), lambda: oracle.newWithinTransProbsLogLik(ho, W[6], B[b - 1], 1, 7, pi), True, new_w=7)
        big = [m for m in mrcas if m > n and len(phy.descendants(m)) >= 3]
        for m in big[:3]:
            step(f"within NNI at {m}", 4, (m, 1), lambda m=m: oracle.withinClusNNIlogLik(ho, W[w - 1], B[b - 1], m, 1, 1, pi), False, nni=True)
        if len(mrcas) > 2:
            step("between NNI", 5, [1] + mrcas, lambda: oracle.betweenClusNNIlogLik(ho, W[w - 1], B[b - 1], mrcas, 1, 1, pi), True, nni=True)
        if big:
            m = big[-1]
            step(f"within NNI x2 at {m}", 4, (m, 2), lambda: oracle.withinClusNNIlogLik(ho, W[w - 1], B[b - 1], m, 2, 1, pi), True, nni=True)
        phy = Phylo(edge, n)
        for m in [x for x in mrcas if x > n][:4]:
            step(f"split {m}", 2, (m,), lambda m=m: oracle.clusSplitMergeLogLik(ho, W[w - 1], B[b - 1], [m], 1, pi), False, split_merge=True)
        by_parent = {}
        for m in mrcas:
            by_parent.setdefault(int(phy.parent[m]), []).append(m)
        for g in [g for g in by_parent.values() if len(g) == 2][:3]:
            step(f"merge {g}", 3, g, lambda g=g: oracle.clusSplitMergeLogLik(ho, W[w - 1], B[b - 1], g, 1, pi), False, split_merge=True)
        splittable = [x for x in mrcas if x > n]
        if splittable:                                             # a split that is kept, then the two halves merged again
            m = splittable[-1]
            kids = phy.kids(m)
            step(f"split {m} kept", 2, (m,), lambda: oracle.clusSplitMergeLogLik(ho, W[w - 1], B[b - 1], [m], 1, pi), True)
            step("between matrices kept", 1, (), lambda: oracle.newBetweenTransProbsLogLik(ho, W[w - 1], B[8], 1, 9, pi), True, new_b=9)
            step(f"merge {kids} kept", 3, kids, lambda: oracle.clusSplitMergeLogLik(ho, W[w - 1], B[b - 1], kids, 1, pi), True)
        step("final within matrices", 0, (), lambda: oracle.newWithinTransProbsLogLik(ho, W[2], B[b - 1], 1, 3, pi), True, new_w=3)
        if shape == "caterpillar":
            assert chains > 0
        if n >= 300 and fuse > 2:
            assert chains > 0 and mats > 0            # root paths were walked, with fused clades materialised off them
    finally:
        mc.close()
        oracle.finalDeallocate(ho)


@pytest.mark.parametrize("seed", range(48))
def test_random_proposal_sequences(lib, oracle, packaged, seed):
    """Random shapes, sizes, fusion thresholds and 14 random moves each (matrices, within / between NNI with 1-3 swaps,
    splits, merges; kept or rolled back at random): the interpreted plans track the oracle through the whole sequence."""
    rng = np.random.default_rng(1000 + seed)
    n, S = int(rng.integers(5, 260)), int(rng.integers(3, 20))
    k, R = int(rng.integers(1, min(12, n // 2) + 1)), int(rng.integers(1, 4))
    kind = str(rng.choice(["evolved", "iid", "diverged", "tiny"]))
    shape, fuse = str(rng.choice(["random", "balanced", "caterpillar"])), int(rng.choice([1, 2, 3, 5, 8, 15]))
    p = Problem(packaged, n, S, k, R, kind, shape, seed=seed)
    W, B, pi = p.W, p.B, p.pi
    masks = oracle.getConvertedAlignment(list("atcg"), p.chars)
    ho = oracle.logLikCpp(p.edge, p.mrcas, pi, W[4], B[4], 1, masks, n, S, 5, 5)["solutionPointer"]
    mc = PlanMachine(lib, p, masks, fuse)
    every = max(1, n // 40)
    try:
        mc.set_matrices(W[4], B[4])
        mc.evaluate()
        edge = _compare(mc, oracle, ho, n, "create", every)
        mrcas, w, b = list(p.mrcas), 5, 5
        for it in range(14):
            phy = Phylo(edge, n)
            move, keep = int(rng.integers(0, 6)), bool(rng.integers(0, 2))
            nni = sm = False
            new_w = new_b = new_mrcas = res = None
            if move == 0:
                new_w = int(rng.integers(1, 11))
                mc.propose(0)
                oracle.newWithinTransProbsLogLik(ho, W[new_w - 1], B[b - 1], 1, new_w, pi)
            elif move == 1:
                new_b = int(rng.integers(1, 11))
                mc.propose(1)
                oracle.newBetweenTransProbsLogLik(ho, W[w - 1], B[new_b - 1], 1, new_b, pi)
            elif move == 2:
                big = [m for m in mrcas if m > n and len(phy.descendants(m)) >= 3]
                if not big:
                    continue
                m, nm, nni = int(rng.choice(big)), int(rng.integers(1, 4)), True
                mc.propose(4, (m, nm))
                res = oracle.withinClusNNIlogLik(ho, W[w - 1], B[b - 1], m, nm, 1, pi)
            elif move == 3:
                nm = int(rng.integers(1, 3))
                try:
                    mc.propose(5, [nm] + mrcas)
                except RuntimeError:                                  # no candidate above the clusters: nothing was drawn
                    continue
                nni = True
                res = oracle.betweenClusNNIlogLik(ho, W[w - 1], B[b - 1], mrcas, nm, 1, pi)
            elif move == 4:
                sp = [m for m in mrcas if m > n]
                if not sp:
                    continue
                m, sm = int(rng.choice(sp)), True
                mc.propose(2, (m,))
                oracle.clusSplitMergeLogLik(ho, W[w - 1], B[b - 1], [m], 1, pi)
                new_mrcas = [x for x in mrcas if x != m] + phy.kids(m)
            else:
                by_parent = {}
                for m in mrcas:
                    by_parent.setdefault(int(phy.parent[m]), []).append(m)
                pairs = [g for g in by_parent.values() if len(g) == 2]
                if not pairs:
                    continue
                g, sm = pairs[int(rng.integers(0, len(pairs)))], True
                mc.propose(3, g)
                oracle.clusSplitMergeLogLik(ho, W[w - 1], B[b - 1], g, 1, pi)
                new_mrcas = [x for x in mrcas if x not in g] + [int(phy.parent[g[0]])]
            mc.set_matrices(W[(new_w or w) - 1], B[(new_b or b) - 1])
            mc.evaluate()
            new_edge = _compare(mc, oracle, ho, n, f"move {it} kind {move}", every)
            if nni:
                assert np.array_equal(new_edge, np.asarray(res["edge"], np.int32))
            if keep:
                edge, w, b = new_edge, new_w or w, new_b or b
                mrcas = new_mrcas if new_mrcas is not None else mrcas
            else:
                oracle.RestorePreviousConfig(ho, edge if nni else NO_EDGE, w, b, nni, mrcas, sm)
                mc.restore(edge if nni else NO_EDGE, nni, mrcas, sm)
                mc.set_matrices(W[w - 1], B[b - 1])
                _compare(mc, oracle, ho, n, f"move {it} rolled back", every)
    finally:
        mc.close()
        oracle.finalDeallocate(ho)


@pytest.mark.parametrize("n,S,k,R,kind,shape,fuse,seed", [
    (64, 24, 9, 1, "evolved", "random", 8, 21), (300, 20, 24, 3, "evolved", "random", 15, 22), (300, 20, 24, 3, "iid", "random", 1, 23),
    (200, 16, 12, 2, "tiny", "random", 8, 24), (150, 16, 10, 3, "iid", "caterpillar", 4, 25), (256, 12, 16, 3, "diverged", "balanced", 15, 26)])
def test_proposal_runs_of_a_batch(lib, oracle, packaged, n, S, k, R, kind, shape, fuse, seed):
    """dmpc_eval_proposals plans every split / merge of a batch as a run (HostTree::proposalRun, the function the product
    calls): interpreted in the batch kernel's way — path partial forwarded, off-path operands from the committed state,
    nothing written — every vertex of every proposal's path equals the oracle's after the same proposal, and the
    committed state is untouched."""
    p = Problem(packaged, n, S, k, R, kind, shape, seed=seed)
    W, B, pi = p.W[4], p.B[4], p.pi
    masks = oracle.getConvertedAlignment(list("atcg"), p.chars)
    ho = oracle.logLikCpp(p.edge, p.mrcas, pi, W, B, 1, masks, n, S, 5, 5)["solutionPointer"]
    mc = PlanMachine(lib, p, masks, fuse)
    try:
        mc.set_matrices(W, B)
        mc.evaluate()
        phy = Phylo(p.edge, n)
        props = [(0, m, 0) for m in p.mrcas if m > n]
        by_parent = {}
        for m in p.mrcas:
            by_parent.setdefault(int(phy.parent[m]), []).append(m)
        props += [(1, g[0], g[1]) for g in by_parent.values() if len(g) == 2]
        assert len(props) >= 2
        for kind_, a, b in props:
            run = mc.run_path(kind_, a, b)
            oracle.clusSplitMergeLogLik(ho, W, B, [a] if kind_ == 0 else [a, b], 1, pi)
            assert run and run[-1][0] == 0                                    # the run ends at the root
            for idx, sol, e in run:
                so, eo = oracle.partial(ho, idx + n)
                assert np.array_equal(sol, so) and np.allclose(e, eo, rtol=2e-7, atol=0), (kind_, a, b, idx)
            oracle.RestorePreviousConfig(ho, NO_EDGE, 5, 5, False, p.mrcas, True)
        _compare(mc, oracle, ho, n, "after the batch", max(1, n // 50))      # the machine's committed state never moved
    finally:
        mc.close()
        oracle.finalDeallocate(ho)


@pytest.mark.parametrize("group_cost,chain_min", [(0, 4), (6, 4), (1000, 4), (128, 1000000)])
def test_scheduler_variants_give_the_same_partials(lib, oracle, packaged, monkeypatch, group_cost, chain_min):
    """The plan builder's knobs — the cap on a path's serial work (0 = one launch per level, the round-1 schedule) and the
    length from which a trailing chain goes to the walk kernel — only change HOW the dirty vertices are scheduled: a few
    random proposal sequences under each setting must still track the oracle bit for bit."""
    monkeypatch.setenv("DMPC_GROUP_COST", str(group_cost))
    monkeypatch.setenv("DMPC_CHAIN_MIN", str(chain_min))
    for seed in (3, 11, 29):
        test_random_proposal_sequences(lib, oracle, packaged, seed)


def test_edge_matrix_in_arbitrary_row_order(lib, oracle, packaged):
    """The incremental edge export and the undo-log rollback of NNI moves rely on the committed edge matrix being the
    engine's own pre-order export — which it checks, not assumes: with the rows of R's matrix in another order the first
    NNI move re-exports the whole tree and a rollback relinks it, and everything still matches the oracle."""
    rng = np.random.default_rng(5)
    p = Problem(packaged, 90, 12, 6, 3, "iid", "random", seed=77)
    perm = rng.permutation(len(p.edge))
    p.edge = np.ascontiguousarray(p.edge[perm])
    W, B, pi = p.W, p.B, p.pi
    masks = oracle.getConvertedAlignment(list("atcg"), p.chars)
    ho = oracle.logLikCpp(p.edge, p.mrcas, pi, W[4], B[4], 1, masks, p.n, p.S, 5, 5)["solutionPointer"]
    mc = PlanMachine(lib, p, masks, 8)
    try:
        mc.set_matrices(W[4], B[4])
        mc.evaluate()
        _compare(mc, oracle, ho, p.n, "create")
        edge = p.edge
        big = [m for m in p.mrcas if m > p.n and len(Phylo(p.edge, p.n).descendants(m)) >= 3]
        for it in range(8):
            m, keep = int(big[it % len(big)]), it % 3 == 2
            mc.propose(4, (m, 1 + it % 2))
            res = oracle.withinClusNNIlogLik(ho, W[4], B[4], m, 1 + it % 2, 1, pi)
            mc.evaluate()
            new_edge = _compare(mc, oracle, ho, p.n, f"nni {it}")
            assert np.array_equal(new_edge, np.asarray(res["edge"], np.int32))
            if keep:
                edge = new_edge
            else:
                oracle.RestorePreviousConfig(ho, edge, 5, 5, True, p.mrcas, False)
                mc.restore(edge, True, p.mrcas, False)
                _compare(mc, oracle, ho, p.n, f"nni {it} rolled back")
    finally:
        mc.close()
        oracle.finalDeallocate(ho)
