"""Soundness of the chain certificate, checked on the CPU in numpy, independently of the kernels.

`certify()` (dmphyclus_b200/csrc/kernels.cuh) claims: wherever its per-chunk inequality holds, the reference's literal fp32
exponent chain (/root/reference/src/AugTree.cpp:306-323) is in the regime where the site's term equals a closed form.  The
GPU tests compare the two on a handful of alignments; here the SAME analysis (records per 64-site chunk, entry leader,
rounding bound, threshold -128) is restated in numpy and run against a literal float32 chain on thousands of synthetic
exponent patterns built to sit near the edges: categories that trade the lead, gaps that creep past the threshold, very
long per-site lists, tiny and huge exponents.  Whenever a chunk is certified, every site of it must (a) have all non-leading
entries of the literal vector below the expf underflow point and (b) give the closed-form term BIT FOR BIT."""
import numpy as np
import pytest

RCH = 64
f32 = np.float32


def literal_chain(rows, rootdot):
    """rows[s] = float32 [K_s][R] (vertex order, zeros allowed); returns terms [S] and the vector after every site."""
    S, R = rootdot.shape
    c = np.zeros(R, f32)
    terms = np.empty(S)
    after = np.empty((S, R), f32)
    for s in range(S):
        for row in rows[s]:
            c = (c + row).astype(f32)                       # one fp32 addition per entry, in order
        m = c.max()
        c = (c - m).astype(f32)
        w = np.exp(c.astype(np.float64)).astype(f32).astype(np.float64)      # expf, promoted
        x = rootdot[s] * w
        a1, a2 = x[0::2].sum(), x[1::2].sum()               # two interleaved accumulators (R <= 4: at most two terms each)
        terms[s] = np.log((a1 + a2) / R) + float(m)
        after[s] = c
    return terms, after


def certificate(rows, R):
    """certify(): returns (last uncertified chunk, leader of the certified suffix)."""
    S = len(rows)
    nrec = (S + RCH - 1) // RCH
    E = np.zeros(R)                                          # entry sums
    ops_incl = 0.0
    ug_acc = 0.0
    last_bad, doms = -1, []
    for g in range(nrec):
        sl = range(g * RCH, min(S, (g + 1) * RCH))
        sd = np.array([rows[s].astype(np.float64).sum(0) if len(rows[s]) else np.zeros(R) for s in sl])
        kmax = np.array([int((rows[s] != 0).sum(0).max()) if len(rows[s]) else 0 for s in sl])
        act = kmax > 0
        ops = float((kmax[act] + 1).sum())
        ab = float(np.abs(sd[act]).sum()) if act.any() else 0.0
        pre = np.cumsum(np.where(act[:, None], sd, 0.0), 0)
        dom = int(np.argmax(E))                              # first maximum, like the kernel's strict '>' scan
        emax, emin = E.max(), E.min()
        ops_incl += ops
        ug_acc += 2.0 ** -22 * ops * ((emax - emin) + ab + 1.0)
        B = np.exp(ops_incl * 2.0 ** -22) * (1.0 + 1e-9) * ug_acc
        bad = False
        for r in range(R):
            if r == dom:
                continue
            max_pre = max(0.0, float((pre[:, r] - pre[:, dom]).max())) if len(pre) else 0.0
            if not (E[r] - E[dom] + max_pre + B < -128.0):
                bad = True
        doms.append(dom)
        if bad:
            last_bad = g
        E = E + (pre[-1] if len(pre) else 0.0)
    leader = doms[last_bad + 1] if last_bad + 1 < nrec else 0
    return last_bad, leader


def make_case(rng, S, R, style):
    rows, rootdot = [], rng.random((S, R)) * 10.0 ** rng.integers(-300, 0, (S, 1))
    p_active = rng.choice([0.05, 0.3, 0.9])
    # per-category probability of a rescale per "event": similar values make the lead change hands
    base = rng.random(R) * 0.5 + 0.25
    if style == "neck_and_neck":
        base[:] = base[0] + rng.normal(0, 0.01, R)
    for s in range(S):
        if rng.random() > p_active:
            rows.append(np.zeros((0, R), f32))
            continue
        K = int(rng.integers(1, 4)) if style != "long_lists" else int(rng.integers(20, 60))
        r_ = np.zeros((K, R), f32)
        for k in range(K):
            hit = rng.random(R) < base
            mag = {"creeping": rng.random(R) * 3.0, "huge": 345.4 * (1 + rng.random(R) * 40)}.get(style, 345.4 + rng.random(R) * 5.0)
            r_[k] = np.where(hit, -mag, 0.0).astype(f32)
        rows.append(r_)
    return rows, rootdot


@pytest.mark.parametrize("style", ["plain", "neck_and_neck", "creeping", "long_lists", "huge"])
@pytest.mark.parametrize("R", [2, 3, 4])
def test_certified_chunks_equal_the_literal_chain(style, R):
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{style}-{R}".encode()))
    certified_sites = 0
    for trial in range(12):
        S = int(rng.integers(1, 9)) * RCH + int(rng.integers(0, RCH))
        rows, rootdot = make_case(rng, S, R, style)
        terms, after = literal_chain(rows, rootdot)
        tb, d = certificate(rows, R)
        for s in range((tb + 1) * RCH, S):
            others = [r for r in range(R) if r != d]
            # (a) the regime: the leader's entry is exactly 0, everyone else's expf underflows to exactly 0
            assert after[s][d] == 0.0 and all(np.exp(np.float64(after[s][r])).astype(f32) == 0.0 for r in others), (style, trial, s)
            # (b) the closed form: log(rootdot_d / R) + the fp32 sum, from zero, of the site's own category-d exponents
            m32 = f32(0.0)
            for row in rows[s]:
                m32 = f32(m32 + row[d])
            closed = np.log((rootdot[s][d] + 0.0) / R) + float(m32)
            assert closed == terms[s], (style, trial, s, closed, terms[s])
            certified_sites += 1
    if style in ("plain", "huge", "long_lists"):
        assert certified_sites > 0          # these regimes do establish themselves: the test is not vacuous


def test_certificate_never_certifies_the_first_chunk():
    rng = np.random.default_rng(1)
    rows, _ = make_case(rng, 5 * RCH, 3, "plain")
    tb, _ = certificate(rows, 3)
    assert tb >= 0                           # the chain starts from zeros: chunk 0 is always walked literally
