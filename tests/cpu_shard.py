"""CPU stand-in for one rank's evaluator in the site-sharded path: the oracle prunes this rank's site block, and the
shard protocol of include/dmpc.h (chain from a carry, per-64-site chunk sums) is restated in numpy from
AugTree.cpp:296-323.  Lets the rank-to-rank exchange of dmphyclus_b200/sharded.py run under gloo without a GPU."""
import numpy as np

CHUNK = 64


class OracleShard:
    name = "oracle-shard"

    def __init__(self, oracle):
        self.o = oracle
        self.quirk_carry = oracle.quirk_carry
        self._site = {}

    def __getattr__(self, k):          # every other entry point goes straight to the oracle
        return getattr(self.o, k)

    def logLikCpp(self, *a, deferred=False):
        return self.o.logLikCpp(*a)

    def _per_site(self, ptr):
        n, s, r = self.o._shape[ptr]
        root, _ = self.o.partial(ptr, n)
        pi = np.array([0.39, 0.22, 0.17, 0.22])
        dot = (root[..., 0] * pi[0] + root[..., 2] * pi[2]) + (root[..., 1] * pi[1] + root[..., 3] * pi[3])
        expo = np.stack([self.o.partial(ptr, v)[1] for v in range(n, 2 * n - 1)])     # [vertex][site][rate], id order
        return dot, expo

    def shard_chain(self, ptr, carry_in):
        dot, expo = self._per_site(ptr)
        s, r = dot.shape
        ev = np.array(carry_in, np.float32).copy()
        ll = np.zeros(s)
        for l in range(s):
            for rr in range(r):
                acc = ev[rr]
                for e in expo[:, l, rr][expo[:, l, rr] != 0]:
                    acc = np.float32(acc + e)
                ev[rr] = acc
            mx = ev.max()
            ev = (ev - mx).astype(np.float32)
            w = np.exp(ev).astype(np.float32).astype(np.float64)
            ll[l] = np.log((dot[l] * w).sum() / r) + float(mx)
        self._site[ptr] = ll
        return ev

    def shard_active(self, ptr):
        _, expo = self._per_site(ptr)
        return bool((expo != 0).any())

    def shard_finish(self, ptr):
        if ptr not in self._site:         # one rate category or textbook mode: no chain ran
            self.shard_chain_reset(ptr)
        ll = self._site.pop(ptr)
        return np.array([ll[i:i + CHUNK].sum() for i in range(0, len(ll), CHUNK)])

    def shard_chain_reset(self, ptr):
        dot, expo = self._per_site(ptr)
        s, r = dot.shape
        ll = np.zeros(s)
        for l in range(s):
            ev = expo[:, l, :].astype(np.float64).sum(0).astype(np.float32)
            mx = ev.max()
            w = np.exp((ev - mx).astype(np.float32)).astype(np.float64)
            ll[l] = np.log((dot[l] * w).sum() / r) + float(mx)
        self._site[ptr] = ll

    def reduce_chunks(self, chunks):
        return float(np.sum(chunks))
