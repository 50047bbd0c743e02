"""Edge-matrix tree utilities: the few `ape` / `phangorn` functions the reference's R driver leans on
(`ape::getMRCA`, `phangorn::Children`, `Descendants`, `Ancestors(type="parent")`;
`/root/reference/R/DMphyClusSource.R:78-88,104,112,146,159,380-386`), restated on the `phylo$edge`
convention the C++ side consumes (`/root/reference/src/AugTree.cpp:59-86`): int matrix [E][2], 1-based,
parents in column 0, tips are 1..N, the root is N+1.
"""
from __future__ import annotations

import numpy as np


class Phylo:
    """Read-only view of an edge matrix with O(1) parent/children lookups (all ids 1-based)."""

    def __init__(self, edge, n_tips: int):
        self.edge = np.ascontiguousarray(edge, dtype=np.int64)
        self.n_tips = int(n_tips)
        n_nodes = int(self.edge.max())
        self.parent = np.zeros(n_nodes + 1, np.int64)
        self.parent[self.edge[:, 1]] = self.edge[:, 0]
        self.root = self.n_tips + 1
        self._children = None
        self._table = None

    def _child_table(self):
        """[node][0..1] = its two children in edge-row order (0, 0 for tips), built with a handful of vector operations;
        None when the tree is not strictly bifurcating.  The MCMC driver rebuilds this view after every accepted
        topology change, and most moves look at a few vertices only."""
        if self._table is None:
            n_nodes = len(self.parent) - 1
            key = self.edge[:, 0]
            if n_nodes < 65536:
                key = key.astype(np.uint16)           # same stable order, by radix sort (10,000 tips: 290 -> 40 us)
            order = np.argsort(key, kind="stable")
            par_sorted, kid_sorted = self.edge[order, 0], self.edge[order, 1]
            if len(order) and len(order) % 2 == 0 and np.all(par_sorted[0::2] == par_sorted[1::2]) \
                    and (len(order) < 4 or np.all(par_sorted[0:-2:2] != par_sorted[2::2])):
                ch = np.zeros((n_nodes + 1, 2), np.int64)
                ch[par_sorted[0::2], 0] = kid_sorted[0::2]
                ch[par_sorted[1::2], 1] = kid_sorted[1::2]
                self._table = ch
            else:
                self._table = False
        return self._table if self._table is not False else None

    def kids(self, v: int) -> list:
        """phangorn::Children(phy, v) in edge-row order."""
        t = self._child_table()
        if t is None:
            return self.children[v]
        return [] if v <= self.n_tips else [int(t[v, 0]), int(t[v, 1])]

    @property
    def children(self):
        """children[v] for every vertex, in edge-row order (AddChild appends), as lists."""
        if self._children is None:
            t = self._child_table()
            if t is not None:
                out = t.tolist()
                for tip in range(self.n_tips + 1):
                    out[tip] = []
            else:                             # not bifurcating: generic path
                out = [[] for _ in range(len(self.parent))]
                for p, c in self.edge:
                    out[p].append(int(c))
            self._children = out
        return self._children

    def split_below(self, node: int, tips) -> tuple[list, list]:
        """phangorn::Children(phy, node) in edge-row order and, for each child, which of `tips` lie below it
        (phangorn::Descendants(phy, child, "tips") restricted to `tips`, as an unordered set).  `tips` must all lie below
        `node` — the tips of a cluster and its MRCA (R/DMphyClusSource.R:104-119) — which is checked.  Every tip climbs
        towards `node` at once (a handful of vector operations per level of the clade), which needs no child table; small
        trees, and views whose table exists anyway, take the plain traversal."""
        if self._table is not None or len(self.parent) <= 8192:
            # the child table exists already, or is cheap to build (it costs a sort of the edge rows: 0.4 ms at 10,000
            # tips, after every accepted topology change): walk the two clades
            kids = self.kids(node)
            return kids, [self.descendants(k) for k in kids]
        tips = np.asarray(tips, dtype=np.int64)
        kids = self.edge[np.flatnonzero(self.edge[:, 0] == node), 1].tolist()
        cur = tips.copy()
        par = self.parent[cur]
        for _ in range(len(self.parent)):
            up = par != node
            if not up.any():
                break
            cur = np.where(up, par, cur)
            par = self.parent[cur]
            if not par.all():                         # parent[root] = 0: a tip climbed past the root
                raise ValueError(f"split_below: a tip does not lie below vertex {node}")
        return kids, [tips[cur == k] for k in kids]

    def is_tip(self, v: int) -> bool:
        return v <= self.n_tips

    def descendants(self, v: int) -> np.ndarray:
        """Tip ids under v (phangorn::Descendants(type='tips')), in traversal order."""
        out, stack, nt = [], [int(v)], self.n_tips
        t = self._child_table()
        if t is not None and self._children is None:      # a few vertices: index the table instead of listing every vertex
            while stack:
                x = stack.pop()
                if x <= nt:
                    out.append(x)
                else:
                    stack.append(int(t[x, 1]))
                    stack.append(int(t[x, 0]))
            return np.array(out, dtype=np.int64)
        ch = self.children
        while stack:
            x = stack.pop()
            if x <= nt:
                out.append(x)
            else:
                stack.extend(reversed(ch[x]))
        return np.array(out, dtype=np.int64)

    def subtree_sizes(self) -> np.ndarray:
        """Number of tips under every node."""
        size = np.zeros(len(self.parent), np.int64)
        size[1:self.n_tips + 1] = 1
        ch = self.children
        for v in self.postorder():
            if v > self.n_tips:
                size[v] = sum(size[c] for c in ch[v])
        return size

    def postorder(self):
        out, stack, ch = [], [(self.root, False)], self.children
        while stack:
            v, done = stack.pop()
            if done or v <= self.n_tips:
                out.append(v)
            else:
                stack.append((v, True))
                for c in reversed(ch[v]):
                    stack.append((c, False))
        return out

    def depth(self) -> np.ndarray:
        d = np.zeros(len(self.parent), np.int64)
        for v in reversed(self.postorder()):
            if v != self.root:
                d[v] = d[self.parent[v]] + 1
        return d

    def mrca(self, tips) -> int:
        """ape::getMRCA for a set of tip (or node) ids."""
        tips = list(tips)
        path = []
        v = tips[0]
        while v:
            path.append(v)
            v = int(self.parent[v])
        pos = {v: i for i, v in enumerate(path)}
        best = 0
        for t in tips[1:]:
            v = t
            while v not in pos:
                v = int(self.parent[v])
            best = max(best, pos[v])
        return path[best]


def cluster_mrcas(phy: Phylo, clus_ind) -> list[int]:
    """clusterNodeIndices as `.initCurrentValue` builds them (DMphyClusSource.R:78-88): for cluster
    k = 1..max: the MRCA of its tips, or the tip id itself for a singleton.  With a single cluster the
    reference stores `Ntip` (a tip id!), so the whole tree is evaluated with between-cluster matrices."""
    clus_ind = np.asarray(clus_ind)
    k_max = int(clus_ind.max())
    if k_max == 1:
        return [phy.n_tips]
    out = []
    for k in range(1, k_max + 1):
        tips = np.nonzero(clus_ind == k)[0] + 1
        out.append(int(tips[0]) if len(tips) == 1 else phy.mrca(tips))
    return out


def check_clades(phy: Phylo, clus_ind) -> bool:
    """`.checkInput` / logLikFromClusInd's clade test (DMphyClusInterface.R:228-249)."""
    clus_ind = np.asarray(clus_ind)
    for k in np.unique(clus_ind):
        tips = np.nonzero(clus_ind == k)[0] + 1
        if len(tips) > 1 and set(phy.descendants(phy.mrca(tips))) != set(tips):
            return False
    return True
