"""Site-sharded evaluation across the GPUs of one box: one process per GPU (`torch.distributed`), each rank
owning one contiguous block of alignment columns behind its own `libdmpc` handle.

Every per-(vertex, site) quantity of the likelihood is independent across sites, so pruning needs no
data-path collective.  The root reduction is where ranks meet:

  * the reference never resets its fp32 exponent vector between loci (`/root/reference/src/AugTree.cpp:306`;
    SURVEY.md §0.2), so with more than one rate category the per-site terms form ONE sequential fp32 chain over
    all loci.  The vector only moves at sites that rescaled somewhere in the tree; a rank whose block holds no such
    site passes it through unchanged.  Every rank first evaluates its block as if the vector entered as zeros and
    the ranks all-gather (has-rescaled-site flag, chunk sums) in ONE collective; only when a rank before the last
    reports a rescaled site does the chain hop through those ranks in site order (one R-float broadcast per hop)
    before a second all-gather;
  * each rank then reduces its sites to per-64-site chunk sums (`CHUNK`); the chunk sums are all-gathered and added in
    a fixed order (`dmpc_reduce_chunks`), so the total is bit-identical for 1, 2, 4 or 8 ranks — which is what
    keeps MCMC trajectories identical across GPU counts.

With NVLink peer mailboxes attached (`Engine.comm_setup`, the normal case on a multi-GPU box) none of this happens on the
host: the reduction kernels exchange chunk sums and certificate records themselves and `_reduce` only reads the result
(`shard_fast_result`); the protocol above is the fallback, and what the CPU tests drive under gloo.

All ranks run the same driver with the same RNG stream and make the same entry-point calls (SPMD); every
rank gets the same log-likelihood back.  `torch.distributed` is plumbing only: payloads are R floats and
S/64 doubles per evaluation.
"""
from __future__ import annotations

import numpy as np

CHUNK = 64    # sites per deterministic reduction chunk (RCH in dmphyclus_b200/csrc/kernels.cuh)


def shard_range(sites_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous site block [begin, end) of `rank`; block sizes are multiples of CHUNK so that chunk
    boundaries — and hence the summation tree — do not depend on the number of ranks."""
    per = -(-sites_total // world)
    per = -(-per // CHUNK) * CHUNK
    b = min(sites_total, rank * per)
    return b, min(sites_total, b + per)


class ShardedEngine:
    """The reference's entry points (same names as `engine.Engine`) over a process group.

    `backend` is this rank's local evaluator: an `engine.Engine` built with `site_begin`/`sites_total`
    (tests plug in a CPU stand-in with the same shard methods to exercise the exchange under gloo)."""

    name = "cuda-sharded"

    def __init__(self, backend, sites_total: int, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.be, self.group = backend, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.sites_total = sites_total
        self.begin, self.end = shard_range(sites_total, self.rank, self.world)
        if self.end <= self.begin:
            raise ValueError(f"rank {self.rank} owns no site: {sites_total} sites over {self.world} ranks "
                             f"(blocks are multiples of {CHUNK})")
        self.device = device if device is not None else torch.device("cpu")
        self._chunks_per_rank = [-(-(e - b) // CHUNK) for b, e in
                                 (shard_range(sites_total, r, self.world) for r in range(self.world))]
        self._R = {}
        self.collectives = 0
        self.fast = 0                      # evaluations finished by the in-kernel peer exchange
        # NVLink peer mailboxes: the reduction kernel then exchanges the chunk sums itself (no NCCL call per evaluation)
        if self.world > 1 and hasattr(backend, "comm_setup") and self.device.type == "cuda":
            def swap(mine: bytes):
                out = [None] * self.world
                dist.all_gather_object(out, mine, group=self.group)
                return out
            backend.comm_setup(self.rank, self.world, swap)

    # ------------------------------------------------------------------ the exchange
    def _gather(self, active: bool, mine: np.ndarray):
        """One all-gather of every rank's (active flag, chunk sums)."""
        torch, dist = self.torch, self.dist
        width = max(self._chunks_per_rank) + 1
        send = torch.zeros(width, dtype=torch.float64)
        send[0] = float(active)
        send[1:1 + len(mine)] = torch.from_numpy(mine)
        send = send.to(self.device)
        recv = torch.empty(width * self.world, dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        self.collectives += 1
        allc = recv.cpu().numpy().reshape(self.world, width)
        flags = [bool(allc[r, 0]) for r in range(self.world)]
        chunks = np.concatenate([allc[r, 1:1 + n] for r, n in enumerate(self._chunks_per_rank)])
        return flags, chunks

    def _reduce(self, ptr) -> float:
        torch, dist = self.torch, self.dist
        if hasattr(self.be, "shard_fast_result"):
            ok, ll = self.be.shard_fast_result(ptr)
            if ok:                                   # every rank sees the same flags, so all take this branch together
                self.fast += 1
                return ll
        R = self._R[ptr]
        chain = R > 1 and getattr(self.be, "quirk_carry", True)
        zero = np.zeros(R, np.float32)
        # optimistic pass: every rank assumes the exponent vector enters its block as zeros.  That is exact
        # unless an EARLIER rank holds a rescaled site (then the chain has to hop through the ranks in order).
        active = False
        if chain:
            self.be.shard_chain(ptr, zero)
            active = self.be.shard_active(ptr)
        flags, chunks = self._gather(active, self.be.shard_finish(ptr))
        if chain and any(flags[:-1]):
            carry = zero
            ran = False
            buf = torch.zeros(R, dtype=torch.float32, device=self.device)
            for a in [r for r in range(self.world) if flags[r]]:
                if a == self.rank:
                    out = self.be.shard_chain(ptr, carry)
                    ran = True
                    buf.copy_(torch.from_numpy(out))
                if a + 1 < self.world:
                    dist.broadcast(buf, src=dist.get_global_rank(self.group, a) if self.group else a, group=self.group)
                    self.collectives += 1
                    if self.rank > a:
                        carry = buf.cpu().numpy().copy()
            if not ran:
                self.be.shard_chain(ptr, carry)        # no rescaled site here: the vector passes through unchanged
            flags, chunks = self._gather(True, self.be.shard_finish(ptr))
        return float(self.be.reduce_chunks(chunks))

    # ------------------------------------------------------------------ entry points
    def getConvertedAlignment(self, equivVector, alignmentAlphaMat):
        return self.be.getConvertedAlignment(equivVector, alignmentAlphaMat)

    def logLikCpp(self, edgeMat, clusterMRCAs, limProbsVec, withinTransMatList, betweenTransMatList, numOpenMP,
                  alignmentBin, numTips, numLoci, withinMatListIndex, betweenMatListIndex):
        """`alignmentBin` is the FULL [numTips][numLoci] matrix or this rank's block [numTips][end-begin]."""
        al = np.asarray(alignmentBin)
        if al.shape[1] == self.sites_total:
            al = al[:, self.begin:self.end]
        if al.shape[1] != self.end - self.begin:
            raise ValueError("alignmentBin must hold all sites or exactly this rank's block")
        res = self.be.logLikCpp(edgeMat, clusterMRCAs, limProbsVec, withinTransMatList, betweenTransMatList,
                                numOpenMP, al, numTips, al.shape[1], withinMatListIndex, betweenMatListIndex,
                                deferred=True)
        ptr = res["solutionPointer"]
        self._R[ptr] = len(np.asarray(withinTransMatList))
        res["logLik"] = self._reduce(ptr)
        return res

    def newBetweenTransProbsLogLik(self, ptr, *a):
        self.be.newBetweenTransProbsLogLik(ptr, *a)
        return self._reduce(ptr)

    def newWithinTransProbsLogLik(self, ptr, *a):
        self.be.newWithinTransProbsLogLik(ptr, *a)
        return self._reduce(ptr)

    def withinClusNNIlogLik(self, ptr, *a):
        res = self.be.withinClusNNIlogLik(ptr, *a)
        res["logLik"] = self._reduce(ptr)
        return res

    def betweenClusNNIlogLik(self, ptr, *a):
        res = self.be.betweenClusNNIlogLik(ptr, *a)
        res["logLik"] = self._reduce(ptr)
        return res

    def clusSplitMergeLogLik(self, ptr, *a):
        self.be.clusSplitMergeLogLik(ptr, *a)
        return self._reduce(ptr)

    def recomputeAll(self, ptr, *a):
        self.be.recomputeAll(ptr, *a)
        return self._reduce(ptr)

    def RestorePreviousConfig(self, ptr, *a):
        self.be.RestorePreviousConfig(ptr, *a)

    def finalDeallocate(self, ptr):
        self._R.pop(ptr, None)
        self.be.finalDeallocate(ptr)

    def getMaxMapSizeEstimate(self, *a):
        return self.be.getMaxMapSizeEstimate(*a)

    def checkAndCullMap(self, ptr, *a):
        self.be.checkAndCullMap(ptr, *a)

    def stats(self, ptr):
        return self.be.stats(ptr)
