"""TEST INFRASTRUCTURE — a numpy interpreter of the engine's evaluation plans.

`libdmpc.so` exposes its host-side scheduler without a device (`dmpc_hp_*`, include/dmpc.h): the same topology /
proposal state machine a GPU handle owns, handing out the plan of every pending evaluation — stored tasks level by
level, token programs of the fused clades, the walked chain with its forwarding flags, materialised clades, slot bits.
This module executes such plans under the kernels' semantics (k_prune / k_walk in dmphyclus_b200/csrc/kernels.cuh:
operands in the reference's child order, matrix set from the task's flag, products and sums rounded separately in
gemv_emul_tinysq order, the two rescale checks of IntermediateNode.cpp:53-66, two slots per vertex with rollback by slot
flip), vectorised over sites and rate categories, so that the CPU suite can check the HOST logic — what gets recomputed,
from which operands, in which order, into which slot — against the oracle's partial vectors, bit for bit, without a GPU.
It is a checker for tests/ only: nothing under dmphyclus_b200/ imports it, and it is far too slow to be a fallback.
"""
import ctypes as C

import numpy as np

# task.hpp
TF_OUTSLOT, TF_C0TIP, TF_C1TIP, TF_C0SLOT, TF_C1SLOT, TF_WITHIN = 1, 2, 4, 8, 16, 32
TF_C0FUSED, TF_C1FUSED, TF_C0FWD, TF_C1FWD, TF_HEAD = 64, 128, 256, 512, 1024
OP_CHERRY, OP_REGTIP, OP_SPILL, OP_SLOTREG = 0, 1, 2, 3
TOK_SET, TOK_SLOT, TOK_FIRST = 8, 16, 32
_i32_p, _u8_p, _dbl_p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_double)


def _i32(a):
    return a.ctypes.data_as(_i32_p)


def matvec(P, v):
    """y = P v per (site, rate): ((P[i,0] v0 + P[i,1] v1) + P[i,2] v2) + P[i,3] v3 — Armadillo's gemv_emul_tinysq order,
    every product and sum rounded on its own (row4 in kernels.cuh).  P: [R][4][4], v: [S][R or 1][4]."""
    out = np.empty(np.broadcast_shapes(v.shape, (1, P.shape[0], 4)))
    for i in range(4):
        out[..., i] = ((P[:, i, 0] * v[..., 0] + P[:, i, 1] * v[..., 1]) + P[:, i, 2] * v[..., 2]) + P[:, i, 3] * v[..., 3]
    return out


def combine_exact(first, second):
    """IntermediateNode.cpp:53-66 for one vertex: first child, check, second child, check (the second OVERWRITES)."""
    s = first.copy()
    e = np.zeros(s.shape[:-1], np.float32)
    for term in (None, second):
        if term is not None:
            s = s * term
        low = np.all(s < 1e-150, axis=-1)
        if low.any():
            m = s.max(axis=-1)
            s[low] = s[low] / m[low][:, None]
            e[low] = np.log(m[low]).astype(np.float32)
    return s, e


class PlanMachine:
    def __init__(self, lib, p, masks, fuse):
        self.lib, self.N, self.S, self.R = lib, p.n, p.S, p.W.shape[1]
        self.masks = np.asarray(masks, np.uint8)
        ind = ((self.masks[..., None] >> np.arange(4)) & 1).astype(np.float64)      # [N][S][4]
        self.ind = ind[:, :, None, :]                                                # [N][S][1][4]
        edge = np.ascontiguousarray(np.asarray(p.edge).T, dtype=np.int32).ravel()
        m = np.asarray(p.mrcas, np.float64)
        self.h = C.c_void_p()
        self._ok(lib.dmpc_hp_create(_i32(edge), len(edge) // 2, p.n, m.ctypes.data_as(_dbl_p), len(m), fuse, C.byref(self.h)))
        n_int = p.n
        self.val = [dict(), dict()]          # slot -> {internal index: partial [S][R][4]}   (stored AND fused vertices)
        self.exp = [dict(), dict()]          # slot -> {internal index: exponent [S][R] float32}
        self.cap = (n_int, 2 * n_int + 8, 2 * n_int + 8, n_int + 2)
        self.last = {}

    def _ok(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.dmpc_last_error().decode())

    def close(self):
        self.lib.dmpc_hp_destroy(self.h)

    def set_matrices(self, within, between):
        self.P = [np.asarray(between, np.float64), np.asarray(within, np.float64)]   # set 0 = between, 1 = within

    def propose(self, kind, args=()):
        a = np.asarray(list(args), np.int32)
        self._ok(self.lib.dmpc_hp_propose(self.h, kind, _i32(a), len(a)))

    def restore(self, edge, nni, mrcas, split_merge):
        e = np.ascontiguousarray(np.asarray(edge).T, dtype=np.int32).ravel()
        m = np.asarray(mrcas, np.int32)
        self._ok(self.lib.dmpc_hp_restore(self.h, _i32(e), len(e) // 2, int(nni), _i32(m), len(m), int(split_merge)))

    def state(self):
        nv = 2 * self.N - 1
        cur, within, edge = np.empty(nv, np.uint8), np.empty(nv, np.uint8), np.empty(2 * (nv - 1), np.int32)
        self._ok(self.lib.dmpc_hp_state(self.h, cur.ctypes.data_as(_u8_p), within.ctypes.data_as(_u8_p), _i32(edge)))
        return cur, within, edge.reshape(2, -1).T.copy()

    def partial(self, v):
        """The committed partial / exponents of vertex v (1-based internal id), stored or fused."""
        cur, _, _ = self.state()
        return self.val[cur[v - 1]][v - 1 - self.N], self.exp[cur[v - 1]][v - 1 - self.N]

    # ------------------------------------------------------------------ plan execution
    def _tip_term(self, tset, tip):
        return matvec(self.P[tset], self.ind[tip])

    def _run_program(self, tokens, tips, flip=0, store=True):
        stack, reg = [], None
        for op_word, a, b, vertex in tokens:
            op = op_word & 7
            if op == OP_SPILL:
                stack.append(reg)
                reg = None
                continue
            tset = 1 if op_word & TOK_SET else 0
            first_flag = bool(op_word & TOK_FIRST)
            if op == OP_CHERRY:
                first, second = self._tip_term(tset, tips[a]), self._tip_term(tset, tips[b])
            elif op == OP_REGTIP:
                x, y = matvec(self.P[tset], reg), self._tip_term(tset, tips[b])
                first, second = (y, x) if first_flag else (x, y)
            elif op == OP_SLOTREG:
                x, y = matvec(self.P[tset], stack.pop()), matvec(self.P[tset], reg)
                first, second = (x, y) if first_flag else (y, x)
            else:
                raise AssertionError("unknown opcode")
            reg, e = combine_exact(first, second)
            if store:
                slot = (1 if op_word & TOK_SLOT else 0) ^ flip
                self.val[slot][vertex], self.exp[slot][vertex] = reg, e
        assert not stack
        return reg

    def evaluate(self):
        """Plans the pending evaluation and executes it.  Returns the counts {tasks, tokens, tips, levels, chain, materialised}."""
        ct, ck, cl, cv = self.cap
        tasks, tokens = np.zeros((ct, 8), np.int32), np.zeros((ck, 4), np.int32)
        tips, levels, counts = np.zeros(cl, np.int32), np.zeros(cv, np.int32), np.zeros(6, np.int32)
        self._ok(self.lib.dmpc_hp_plan(self.h, _i32(tasks), ct, _i32(tokens), ck, _i32(tips), cl, _i32(levels), cv, _i32(counts)))
        n_t, _, _, n_l, chain, _ = [int(x) for x in counts]
        first_chain_level = n_l - chain
        prev = None
        for lv in range(n_l):
            for i in range(levels[lv], levels[lv + 1]):
                out, c0, c1, fl, prog, ntok, tip_off, n_tips = [int(x) for x in tasks[i]]
                tset = 1 if fl & TF_WITHIN else 0
                my_tips = tips[tip_off:tip_off + n_tips]
                n0 = ntok & 0xffff
                terms = []
                for j, (c, is_tip, is_fused, slot_bit, fwd_bit) in enumerate((
                        (c0, fl & TF_C0TIP, fl & TF_C0FUSED, fl & TF_C0SLOT, fl & TF_C0FWD),
                        (c1, fl & TF_C1TIP, fl & TF_C1FUSED, fl & TF_C1SLOT, fl & TF_C1FWD))):
                    if is_tip:
                        terms.append(self._tip_term(tset, c))
                        continue
                    if is_fused:
                        lo = prog + (n0 if j else 0)
                        hi = lo + ((ntok >> 16) & 0xffff if j else n0)
                        value = self._run_program([tuple(int(x) for x in t) for t in tokens[lo:hi]], my_tips)
                    else:
                        value = self.val[1 if slot_bit else 0][c]
                        if fwd_bit:                 # k_walk keeps it in registers: it must be what the previous step wrote
                            assert lv > first_chain_level and prev is not None and prev[0] == c and value is prev[1]
                    terms.append(matvec(self.P[tset], value))
                if lv >= first_chain_level and chain:
                    assert bool(fl & TF_HEAD) == (lv == first_chain_level) and not fl & (TF_C0FUSED | TF_C1FUSED)
                    assert (lv == first_chain_level) or bool(fl & TF_C0FWD) != bool(fl & TF_C1FWD)
                else:
                    assert not fl & (TF_HEAD | TF_C0FWD | TF_C1FWD)
                sol, e = combine_exact(terms[0], terms[1])
                oslot = 1 if fl & TF_OUTSLOT else 0
                self.val[oslot][out], self.exp[oslot][out] = sol, e
                prev = (out, sol)
        self.last = dict(tasks=n_t, levels=n_l, chain=chain, materialised=int(counts[5]), tokens=int(counts[1]))
        return self.last

    def run_path(self, kind, a, b=0):
        """One proposal of a dmpc_eval_proposals batch (k_prune<false, PATHS>): the run of stored vertices from the changed
        vertex to the root, the path partial forwarded from task to task, only off-path operands read, NOTHING written.
        Returns [(internal index, partial, exponents)] bottom-up."""
        ct, ck, cl, _ = self.cap
        tasks, tokens = np.zeros((ct, 8), np.int32), np.zeros((ck, 4), np.int32)
        tips, counts = np.zeros(cl, np.int32), np.zeros(3, np.int32)
        self._ok(self.lib.dmpc_hp_path(self.h, kind, a, b, _i32(tasks), ct, _i32(tokens), ck, _i32(tips), cl, _i32(counts)))
        out, prev = [], None
        for i in range(int(counts[0])):
            vtx, c0, c1, fl, prog, ntok, tip_off, n_tips = [int(x) for x in tasks[i]]
            tset = 1 if fl & TF_WITHIN else 0
            my_tips = tips[tip_off:tip_off + n_tips]
            n0 = ntok & 0xffff
            assert not fl & TF_HEAD and bool(fl & (TF_C0FWD | TF_C1FWD)) == (i > 0) and (fl & (TF_C0FWD | TF_C1FWD)) != (TF_C0FWD | TF_C1FWD)
            terms = []
            for j, (c, is_tip, is_fused, slot_bit, fwd_bit) in enumerate((
                    (c0, fl & TF_C0TIP, fl & TF_C0FUSED, fl & TF_C0SLOT, fl & TF_C0FWD),
                    (c1, fl & TF_C1TIP, fl & TF_C1FUSED, fl & TF_C1SLOT, fl & TF_C1FWD))):
                if is_tip:
                    terms.append(self._tip_term(tset, c))
                elif fwd_bit:
                    assert prev[0] == c
                    terms.append(matvec(self.P[tset], prev[1]))
                elif is_fused:
                    lo = prog + (n0 if j else 0)
                    hi = lo + ((ntok >> 16) & 0xffff if j else n0)
                    terms.append(matvec(self.P[tset], self._run_program([tuple(int(x) for x in t) for t in tokens[lo:hi]],
                                                                        my_tips, store=False)))
                else:
                    terms.append(matvec(self.P[tset], self.val[1 if slot_bit else 0][c]))
            sol, e = combine_exact(terms[0], terms[1])
            prev = (vtx, sol)
            out.append((vtx, sol, e))
        return out
