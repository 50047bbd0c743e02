#!/usr/bin/env python
"""bench.py — throughput of the Felsenstein-pruning likelihood path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config3] [--data evolved|iid]

One "step" = one full evaluation of the likelihood (InvalidateAll + ComputeLoglik: every internal vertex of the
tree re-pruned at every site and rate category, then the root reduction) on one batch of synthetic input.  The
metric is BASELINE.json's: node-site partial updates/sec, one update = one 4-state conditional-likelihood vector
for one (internal vertex, site, rate category); `mcmc` carries its second half, MCMC iterations/sec.

Workload: config 3 (10,000 sequences x 30,000 sites, 100 clusters, 3 rate categories — the shape BASELINE.json's
target is quoted on; it fits one B200: 66 GB).  N > 1 (launched by torchrun, one rank per GPU): the SAME 30,000 sites
are split over the ranks (strong scaling; `--weak` gives every rank the full site count instead); the ranks meet only
in the root reduction, inside the kernel, over NVLink peer mailboxes.  `also_config2` repeats the device-resident and
end-to-end numbers on round 1's headline shape (1,000 x 10,000).

  value      device-resident: the alignment, tree and buffers already live in HBM; K x dmpc_recompute_all, timed
             with CUDA events on the engine's own stream, max over ranks.
  e2e        the same evaluation through the reference-facing call (`logLikCpp` = dmpc_loglik_create, then
             `finalDeallocate`) from pinned HOST buffers: tree build, alignment upload, pruning, reduction and
             the read-back of the log-likelihood all inside the timed region (wall clock around synchronous calls).
  roofline   k_prune (the dominant kernel).  Small-clade fusion keeps ~88 % of the vertices out of HBM, so the kernel is
             bound by the rate at which the fp64 pipe retires separately-rounded DMUL / DADD (the reference's evaluation
             order forbids FMA): achieved = the DMUL + DADD the timed launches had to issue / their device time
             (CUDA events around the level launches), peak = tools/fp64_peak.cu measured on this GPU in this run.
             `hbm_equiv` keeps the algorithmic-bytes view (68 B per update) against MEASURED_PEAKS.json.
  cpu_baseline  the reference's own translation units (oracle/_ref, single-threaded — its OpenMP is commented
             out) on a bounded sample of the same alignment, on this box's host cores; plus the
             reference-driven MCMC rate on a bounded sample.

`--impl reference` times the reference's CPU implementation (oracle/_ref/libref.so, else the oracle port) on the
same workload and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

emit = print
METRIC = "node-site partial updates/sec"
UNIT = "updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--data", default="evolved", choices=["evolved", "iid", "matched"])
    ap.add_argument("--tree", default="random", choices=["random", "balanced", "caterpillar"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (0 = min(steps, 20))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the MCMC / config 2 / other-alignment extra legs")
    ap.add_argument("--all-extras", action="store_true", help="also run the rescale-stress (other alignment kind) leg")
    ap.add_argument("--mcmc-only", action="store_true", help="extra legs: only the MCMC one (skip the other alignment kind)")
    ap.add_argument("--fuse", type=int, default=None, help="small-clade fusion threshold (tips); default = library default")
    ap.add_argument("--sweep-rates", type=int, default=3, help="rate categories of the proposal sweep")
    ap.add_argument("--proposal-sweep", action="store_true", help="config 4: batches of 64..4096 split/merge proposals per launch")
    ap.add_argument("--weak", action="store_true", help="N > 1: give every rank the workload's full site count (weak scaling) instead of splitting the workload's sites across the ranks (strong scaling, the default: BASELINE.json's target is config 3 over 8 GPUs)")
    ap.add_argument("--seed", type=int, default=2001)
    args = ap.parse_args()
    args.strong = not args.weak
    return args


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak(device):
    """DMUL + DADD per second of this GPU, measured now by tools/fp64_peak.cu (built by __graft_entry__.build())."""
    import ctypes
    lib = os.path.join(ROOT, "tools", "libfp64peak.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared",
                               "-Xcompiler", "-fPIC", "-o", lib, os.path.join(ROOT, "tools", "fp64_peak.cu")])
    L = ctypes.CDLL(lib)
    L.fp64_issue_peak.restype = ctypes.c_double
    L.fp64_issue_peak.argtypes = [ctypes.c_int, ctypes.c_int]
    v = L.fp64_issue_peak(device, 5)
    return v if v > 0 else None


def ncu_traffic(workload, data):
    """dram__bytes_read.sum + dram__bytes_write.sum per k_prune launch from the newest committed ncu capture of this
    workload (profiles/*_traffic.json, written by tools/ncu_summary.py), or None."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
        try:
            d = json.load(open(f))
        except Exception:
            continue
        if d.get("workload") == workload and d.get("data") == data:
            best = (d, os.path.basename(f))
    if best is None:
        return None, None
    return best[0]["dram_bytes_per_launch"], best[1]


def algorithmic_bytes(n, s, r):
    """per full evaluation, S*R*[(N-2)*32 + (N-1)*36] + S*N bytes (~68 B per update)."""
    return s * r * ((n - 2) * 32 + (n - 1) * 36) + s * n


def frozen(a):
    """A read-only copy: the binding converts a read-only array it is handed again (matrix lists, limProbs — R hands the
    same objects to every call) once instead of on every call (engine.Engine._mp)."""
    a = np.array(a)
    a.setflags(write=False)
    return a


def make_problem(args, world):
    from dmphyclus_b200.workloads import CONFIGS, Problem, load_packaged
    n, s, k, r = CONFIGS[args.workload]
    if args.strong or world == 1:      # strong scaling: the workload's own site count, sharded
        return Problem(load_packaged(), n, s, k, r, args.data, args.tree, seed=args.seed), (n, s // world, k, r)
    # weak scaling: every rank holds `s` sites of an alignment of s * world sites
    return Problem(load_packaged(), n, s * world, k, r, args.data, args.tree, seed=args.seed), (n, s, k, r)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled through NVML while the GPU works."""

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.stop_flag, self.ok = device, [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:   # the number is still reported; the record says why clocks are missing
            self.err = repr(e)
        self.mark = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.samples.append((time.perf_counter(), sm, rs, util, self.mark))
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable: " + self.err}
        timed = [x for x in self.samples if x[4] == "timed"] or [x for x in self.samples if x[4] is not None]
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        bits = 0
        for x in timed:
            bits |= x[2]
        bits &= ~0x1
        return {"sm_mhz": float(np.median([x[1] for x in timed])) if timed else None, "sm_max_mhz": float(self.max_sm),
                "reasons": [v for k, v in names.items() if bits & k], "samples": len(timed)}


def cpu_reference_backend():
    from oracle.oracle import Oracle, Reference
    try:
        return Reference(), "reference"
    except Exception:
        return Oracle(), "port"


def time_cpu(be, p, n_sites, reps=1):
    """One logLikCpp (tree build, dictionary fill, pruning, reduction) on the first n_sites columns."""
    from dmphyclus_b200 import synth
    al = be.getConvertedAlignment(list(synth.STATES), p.chars[:, :n_sites])
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        res = be.logLikCpp(p.edge, p.mrcas, p.pi, p.W[p.w_idx - 1], p.B[p.b_idx - 1], 1, al, p.n, n_sites, p.w_idx, p.b_idx)
        dt = time.perf_counter() - t0
        if hasattr(be, "dictionary_size"):      # the reference memoises (node, site) solutions: how many were distinct
            time_cpu.dictionary = be.dictionary_size(res["solutionPointer"])
        be.finalDeallocate(res["solutionPointer"])
        best = dt if best is None else min(best, dt)
    return best, res["logLik"]


def reference_sample_sites(be, p, s, budget_s):
    """Sites of a bounded sample: the reference's cost is calibrated on 16 columns and the sample sized for ~budget_s
    of CPU work (/1.5: its std::map look-ups slow down as the dictionary grows)."""
    t_probe, _ = time_cpu(be, p, 16)
    per_site = t_probe / 16
    return int(max(16, min(s, budget_s / per_site / 1.5))), per_site


def reference_mcmc(p, s, per_site, budget_s=16.0):
    """MCMC iterations/sec of the reference's own likelihood code under the same driver (`chain.run_chain`, the mirror of
    `.performStepPhylo`, /root/reference/R/DMphyClusSource.R:1-26), on a bounded sample of sites: a full-size iteration
    of config 3 would take minutes.  The sample is sized for ~budget_s; the rate it would have on all sites is
    extrapolated linearly in the site count (its work per likelihood call is per site; the memo dictionary makes it
    somewhat sub-linear, so the estimate flatters the reference)."""
    from dmphyclus_b200 import chain, synth
    from oracle.oracle import Reference
    try:
        be = Reference()
        kind = "reference"
    except Exception:
        from oracle.oracle import Oracle
        be, kind = Oracle(), "port"
    iters = 2
    m = int(max(16, min(s, budget_s / (iters + 1) / 5.0 / per_site)))   # an iteration costs ~5 full evaluations of the sample
    al = be.getConvertedAlignment(list(synth.STATES), p.chars[:, :m])
    cs = chain.ChainSettings(limProbs=p.pi, withinTransMatAll=p.W, betweenTransMatAll=p.B, shapeForAlpha=1000,
                             scaleForAlpha=0.1, alphaMin=10, poisRateNumClus=2, numSplitMergeMoves=3)
    marks = []
    t0 = time.perf_counter()
    _, calls = chain.run_chain(be, al, p.edge, p.clus, 100.0, cs, iters, seed=2001,
                               on_iter=lambda it, cv: marks.append((time.perf_counter(), cv.n_loglik_calls)))
    rate = (len(marks) - 1) / (marks[-1][0] - marks[0][0]) if len(marks) > 1 else iters / (time.perf_counter() - t0)
    return {"iterations_per_sec": rate, "kind": kind, "cores": 1, "iterations": iters, "likelihood_calls": calls,
            "sample": f"first {m} of {s} sites, all {p.n} tips, {iters} iterations of the .performStepPhylo mirror "
                      f"(rate after the first iteration)",
            "iterations_per_sec_all_sites_estimate": rate * m / s,
            "estimate": "sample rate x sample sites / all sites (work per call is per site)"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    p, (n, s, k, r) = make_problem(args, 1)
    be, kind = cpu_reference_backend()
    # size the per-step sample so that the whole run stays within ~75 s, and a step within ~4 s
    budget = min(4.0, 75.0 / max(1, args.steps + args.warmup))
    ns, per_site = reference_sample_sites(be, p, s, budget)
    for _ in range(args.warmup):
        time_cpu(be, p, ns)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, ll = time_cpu(be, p, ns)
    dt = time.perf_counter() - t0
    upd = (n - 1) * ns * r
    val = upd * args.steps / dt
    sample = (f"first {ns} of {s} sites, all {n} tips, one logLikCpp per step (tree build + dictionary fill + pruning + "
              f"reduction); the reference memoises (node, site) solutions, so its rate per update rises a little with the "
              f"number of sites")
    lookups = (2 * n - 1) * ns
    memo = getattr(time_cpu, "dictionary", None)
    cfg = config_dict(args, n, s, k, r, 1)
    cfg["reference_sample"] = sample
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
           "scaling": "strong" if args.strong else "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": cfg,
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                            "host_cores": os.cpu_count(),
                            "dictionary_entries": memo, "dictionary_hit_rate": (1.0 - memo / lookups) if memo else None,
                            "note": "the reference's likelihood is single-threaded (OpenMP commented out, "
                                    "clusterFunArmadillo.cpp:4,38); it cannot use more host threads"},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    if not args.no_extras:
        out["mcmc"] = reference_mcmc(p, s, per_site)
    emit(json.dumps(out))


def CONFIGS_SITES(args):
    from dmphyclus_b200.workloads import CONFIGS
    return CONFIGS[args.workload][1]


def config_dict(args, n, s, k, r, world):
    return {"workload": f"{args.workload}: logLikFromClusInd-style full evaluation, {n} seqs x {s} sites per GPU, "
                        f"{k} clusters, {r} rate categories",
            "tips": n, "sites_per_gpu": s, "sites_total": (CONFIGS_SITES(args) if args.strong else s * world), "clusters": k, "rate_categories": r,
            "alignment": args.data, "tree": args.tree, "seed": args.seed,
            "matrices": "packaged withinClusTransProbs[[5]] / betweenClusTransProbs[[5]]",
            "l2": "inputs larger than L2 (partials of one evaluation >> 126 MB); no explicit flush",
            "parallelism": f"site-sharded x{world}" if world > 1 else "single GPU"}


def run_ours(args, rank, world, local_rank):
    import torch
    from dmphyclus_b200 import build, synth
    from dmphyclus_b200.engine import Engine
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the likelihood engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p, (n, s, k, r) = make_problem(args, world)
    S_total = p.S
    W, B, pi = frozen(p.W[p.w_idx - 1]), frozen(p.B[p.b_idx - 1]), frozen(p.pi)
    edge = frozen(np.asarray(p.edge, np.int32))
    sampler = ClockSampler(local_rank)
    sampler.start()

    if world > 1:
        from dmphyclus_b200.sharded import ShardedEngine, shard_range
        local = Engine(device=local_rank, site_begin=shard_range(S_total, rank, world)[0], sites_total=S_total,
                       fuse_tips=args.fuse)
        eng = ShardedEngine(local, S_total, device=torch.device("cuda", local_rank))
        b0, b1 = eng.begin, eng.end
    else:
        local = eng = Engine(device=local_rank, fuse_tips=args.fuse)
        b0, b1 = 0, S_total
    s_local = b1 - b0
    # this rank's alignment block in pinned host memory (the e2e leg uploads it every step)
    masks = local.getConvertedAlignment(list(synth.STATES), p.chars[:, b0:b1])
    pinned = torch.empty(masks.shape, dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[...] = masks
    al = pinned.numpy()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- device-resident leg
    res = eng.logLikCpp(edge, p.mrcas, pi, W, B, 1, al, p.n, s_local, p.w_idx, p.b_idx)
    h, ll0 = res["solutionPointer"], res["logLik"]
    single_ll = None
    if world > 1 and rank == 0:
        # the sharded value must be the single-GPU value, bit for bit: one evaluation of the whole alignment on this
        # rank's device (one slot of partials: 33 GB for config 3), outside every timed region
        one = Engine(device=local_rank, fuse_tips=args.fuse)
        full = one.getConvertedAlignment(list(synth.STATES), p.chars)
        r1 = one.logLikCpp(edge, p.mrcas, pi, W, B, 1, full, p.n, S_total, p.w_idx, p.b_idx)
        single_ll = r1["logLik"]
        one.finalDeallocate(r1["solutionPointer"])
        del full
        assert single_ll == ll0, ("sharded log-likelihood differs from the single-GPU one", ll0, single_ll)
    sampler.mark = "warm"
    for _ in range(max(3, args.warmup)):
        ll = eng.recomputeAll(h, W, B, pi)
    assert ll == ll0, (ll, ll0)          # a full re-evaluation is deterministic
    st0 = local.stats(h)
    prune_ms = 0.0
    barrier()
    sampler.mark = "timed"
    local.timer_begin(h)
    for _ in range(args.steps):                     # the timed region: nothing but the K evaluations
        eng.recomputeAll(h, W, B, pi)
    ms = local.timer_end(h)
    # k_prune's device time over exactly these K evaluations: the engine brackets every evaluation's pruning launches with
    # CUDA events on its own stream and adds the elapsed times up as the results come in (no extra synchronisation)
    prune_ms = local.stats(h)["prune_ms_total"] - st0["prune_ms_total"]
    barrier()
    sampler.mark = "after"
    ms = max_over_ranks(ms)
    st1 = local.stats(h)
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    # a second pass of the same evaluations, read out one by one: where the time goes (kernel share, host side, root stage)
    host_us = np.zeros(3)
    tl_keys, tl_sum, tl_n = None, None, 0
    probe_steps = min(args.steps, 10)
    probe_ms = 0.0
    for _ in range(probe_steps):
        eng.recomputeAll(h, W, B, pi)
        sti = local.stats(h)
        tl = local.root_timeline(h)
        if tl_keys is None:
            tl_keys, tl_sum = list(tl), np.zeros(len(tl))
        if all(v is not None for v in tl.values()):
            tl_sum += [tl[k] for k in tl_keys]
            tl_n += 1
        probe_ms += sti["last_eval_ms"]
        host_us += (sti["last_host_plan_us"], sti["last_host_submit_us"], sti["last_host_total_us"])
    host_us /= probe_steps
    prune_launches = st1["last_prune_launches"] * args.steps
    updates_step = (n - 1) * S_total * r
    value = updates_step * args.steps / (ms * 1e-3)
    peak, peak_src = peaks()
    # compulsory HBM bytes of one evaluation's pruning as the engine counted them for the plan it ran (stored
    # partials written and read, mask bytes); with small-clade fusion this is far below the unfused 68 B/update
    bytes_eval = st1["last_algorithmic_bytes"]
    bytes_unfused = algorithmic_bytes(n, s_local, r)
    achieved = bytes_eval * args.steps / (prune_ms * 1e-3) / 1e9
    dev_bytes = st1["device_bytes"]
    # Roofline of the dominant kernel (k_prune).  `achieved` follows the contract: SURVEY.md §8(d)'s per-unit figure
    # (68 B per node-site partial update: children read, partial + exponent written, every vertex stored) x the updates
    # of the timed launches / their device time.  Small-clade fusion removes most of that traffic — about four fifths
    # of the vertices are recomputed in registers and never touch HBM — so `achieved` can exceed the HBM peak; the bytes
    # the kernel really has to move (`moved_*`, counted by the engine for the plan it ran) and the ncu DRAM traffic say
    # by how much.  See DESIGN.md §3.
    traffic, traffic_src = ncu_traffic(args.workload, args.data)
    # ... which is why the roof that binds is the fp64 pipe's DMUL / DADD issue rate (no FMA: the reference's evaluation
    # order rounds products and sums separately), measured on this GPU right now by tools/fp64_peak.cu
    ops_eval = st1["last_fp64_ops"]
    pk64 = fp64_peak(local_rank)
    tfl = ops_eval * args.steps / (prune_ms * 1e-3) / 1e12
    roofline = {"bound": "fp64_issue", "kernel": "k_prune", "achieved": tfl, "peak": (pk64 / 1e12 if pk64 else None),
                "unit": "TFLOP/s", "frac": (tfl / (pk64 / 1e12) if pk64 else None),
                "peak_source": "tools/fp64_peak.cu on this GPU, this run: DMUL + DADD per second, 8 independent chains per "
                               "thread, every SM at full occupancy (FMA is not available to this kernel)",
                "fp64_ops_per_update": ops_eval / ((n - 1) * s_local * r),
                "fp64_ops_per_launch": ops_eval / max(1, st1["last_prune_launches"]),
                "traffic": traffic, "traffic_source": traffic_src,
                "launches_per_step": int(st1["last_prune_launches"]), "kernel_ms_per_step": prune_ms / args.steps,
                "kernel_share_of_step": prune_ms / ms,
                "stored_vertices": int(st1["last_stored_vertices"]), "fused_vertices": int(st1["last_fused_vertices"]),
                "hbm_equiv": {"achieved": bytes_unfused * args.steps / (prune_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": bytes_unfused * args.steps / (prune_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                              "algorithmic_bytes_per_update": bytes_unfused / ((n - 1) * s_local * r),
                              "algorithmic_bytes_per_launch": bytes_unfused / max(1, st1["last_prune_launches"]),
                              "moved_bytes_per_update": bytes_eval / ((n - 1) * s_local * r), "moved_gbs": achieved,
                              "moved_frac_of_peak": achieved / peak,
                              "note": "algorithmic bytes = every vertex stored (68 B per update); small-clade fusion "
                                      "never moves most of them, so frac > 1 is not a bandwidth fraction"}}
    kernel_name = "k_prune"
    eng.finalDeallocate(h)

    # ---------------------------------------------------------------- end-to-end leg (host buffers in, scalar out)
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    for _ in range(3):
        res = eng.logLikCpp(edge, p.mrcas, pi, W, B, 1, al, p.n, s_local, p.w_idx, p.b_idx)
        eng.finalDeallocate(res["solutionPointer"])
    barrier()
    sampler.mark = "e2e"
    t0 = time.perf_counter()
    e2e_parts = np.zeros(4)
    for _ in range(e2e_steps):
        res = eng.logLikCpp(edge, p.mrcas, pi, W, B, 1, al, p.n, s_local, p.w_idx, p.b_idx)
        sti = local.stats(res["solutionPointer"])
        e2e_parts += (sti["create_setup_us"] * 1e-3, sti["create_upload_ms"], sti["last_eval_ms"], sti["last_host_total_us"] * 1e-3)
        eng.finalDeallocate(res["solutionPointer"])
    barrier()
    e2e_parts /= e2e_steps
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    sampler.mark = "after"
    assert res["logLik"] == ll0, (res["logLik"], ll0)
    e2e_value = updates_step * e2e_steps / (e2e_ms * 1e-3)
    h2d = al.nbytes + 2 * (2 * n - 2) * 4 + 2 * r * 16 * 8 + 4 * 8 + 8 * len(p.mrcas) + (n - 1) * 16
    d2h = 8 + 4 + 4

    extras = {}
    if not args.no_extras:
        extras = extra_legs(args, eng, p, al, n, s, k, r, world)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(p, n, s, r, with_mcmc=not args.no_extras)
        if cpu.get("mcmc") and extras.get("mcmc"):
            extras["mcmc"]["vs_reference_all_sites_estimate"] = (extras["mcmc"]["iterations_per_sec"] /
                                                                  cpu["mcmc"]["iterations_per_sec_all_sites_estimate"])

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": config_dict(args, n, s, k, r, world),
               "clocks": sampler.summary(),
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                       "breakdown_ms": {"host_setup": e2e_parts[0], "upload_device": e2e_parts[1], "evaluation_device": e2e_parts[2],
                                        "evaluate_call_host": e2e_parts[3]},
                       "call": "Engine.logLikCpp (dmpc_loglik_create) + finalDeallocate, alignment in pinned host memory"},
               "gpu_launches": int(launches),
               "roofline": roofline,
               "host_us_per_step": {"plan": host_us[0], "submit": host_us[1], "evaluate_total": host_us[2]},
               "root_stage_us": ({k: round(float(v), 2) for k, v in zip(tl_keys, tl_sum / tl_n)} if tl_n else None),
               "sharded_exchange": ({"in_kernel": eng.fast, "host_collectives": eng.collectives} if world > 1 else None),
               "chain": {"resolved_by": {0: "no site rescaled", 1: "certificate", 2: "literal walk"}.get(int(st1["last_cert"])),
                         "walked_sites": int(st1["last_chain_sites"]), "walked_rows": int(st1["last_chain_rows"]),
                         "rescaled_tiles": int(st1["last_active_tiles"]), "dominant_rate": int(st1["last_cert_rate"])},
               "cpu_baseline": cpu, "logLik": ll0, "single_gpu_logLik_equal": (None if single_ll is None else bool(single_ll == ll0)),
               "device_bytes": int(dev_bytes),
               "levels": int(st1["last_levels"])}
        out.update(extras)
        emit(json.dumps(out))
    if dist is not None:
        dist.barrier()
        out_fast = (eng.fast, eng.collectives)
        local.comm_close()
        dist.destroy_process_group()


def extra_legs(args, eng, p, al, n, s, k, r, world=1):
    """Secondary numbers (not the headline): the rescale-stress alignment and MCMC iterations/sec."""
    from dmphyclus_b200 import chain, synth
    from dmphyclus_b200.workloads import Problem, load_packaged
    out = {}
    if world > 1:
        return mcmc_leg(args, eng, p, al, out, "site-sharded: every rank runs the same driver in lock-step")
    if args.mcmc_only:
        return mcmc_leg(args, eng, p, al, out, "single GPU")
    if args.workload != "config2":
        out["also_config2"] = config2_leg(args, eng)
    if not args.all_extras:
        return mcmc_leg(args, eng, p, al, out, "single GPU")
    other = "iid" if args.data == "evolved" else "evolved"
    q = Problem(load_packaged(), n, s, k, r, other, args.tree, seed=args.seed)
    alq = eng.getConvertedAlignment(list(synth.STATES), q.chars)
    res = eng.logLikCpp(q.edge, q.mrcas, q.pi, q.W[4], q.B[4], 1, alq, q.n, s, 5, 5)
    h = res["solutionPointer"]
    for _ in range(3):
        eng.recomputeAll(h, q.W[4], q.B[4], q.pi)
    steps = max(5, min(args.steps, 50))
    eng.timer_begin(h)
    pr = 0.0
    for _ in range(steps):
        eng.recomputeAll(h, q.W[4], q.B[4], q.pi)
        pr += eng.stats(h)["last_prune_ms"]
    ms = eng.timer_end(h)
    st = eng.stats(h)
    out[f"extra_{other}"] = {"value": (n - 1) * s * r * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
                             "prune_ms_per_step": pr / steps, "rescale_rows": st["exponent_capacity"],
                             "logLik": res["logLik"]}
    eng.finalDeallocate(h)
    return mcmc_leg(args, eng, p, al, out, "single GPU")


def config2_leg(args, eng):
    """Round 1's headline shape (1,000 x 10,000, 20 clusters): device-resident and end-to-end, same definitions."""
    import torch
    from dmphyclus_b200 import synth
    from dmphyclus_b200.workloads import CONFIGS, Problem, load_packaged
    n, s, k, r = CONFIGS["config2"]
    q = Problem(load_packaged(), n, s, k, r, args.data, args.tree, seed=args.seed)
    masks = eng.getConvertedAlignment(list(synth.STATES), q.chars)
    pinned = torch.empty(masks.shape, dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[...] = masks
    alq = pinned.numpy()
    W, B, pi = frozen(q.W[4]), frozen(q.B[4]), frozen(q.pi)
    edge = frozen(np.asarray(q.edge, np.int32))
    res = eng.logLikCpp(edge, q.mrcas, pi, W, B, 1, alq, n, s, 5, 5)
    h = res["solutionPointer"]
    for _ in range(5):
        eng.recomputeAll(h, W, B, pi)
    steps = 50
    eng.timer_begin(h)
    for _ in range(steps):
        eng.recomputeAll(h, W, B, pi)
    ms = eng.timer_end(h) / steps
    eng.finalDeallocate(h)
    for _ in range(3):
        eng.finalDeallocate(eng.logLikCpp(edge, q.mrcas, pi, W, B, 1, alq, n, s, 5, 5)["solutionPointer"])
    t0 = time.perf_counter()
    for _ in range(20):
        eng.finalDeallocate(eng.logLikCpp(edge, q.mrcas, pi, W, B, 1, alq, n, s, 5, 5)["solutionPointer"])
    e2e = (time.perf_counter() - t0) / 20 * 1e3
    upd = (n - 1) * s * r
    return {"workload": f"config2: {n} seqs x {s} sites, {k} clusters, {r} rate categories", "value": upd / (ms * 1e-3),
            "unit": UNIT, "ms_per_step": ms, "e2e_value": upd / (e2e * 1e-3), "e2e_ms_per_step": e2e, "logLik": res["logLik"]}


def mcmc_leg(args, eng, p, al, out, how):
    """MCMC iterations/sec: the reference's move schedule (.performStepPhylo) driven from the host mirror, then a short
    second pass that reads the engine's own counters after every likelihood call to say where a call's time goes."""
    from dmphyclus_b200 import chain
    cs = chain.ChainSettings(limProbs=p.pi, withinTransMatAll=p.W, betweenTransMatAll=p.B, shapeForAlpha=1000,
                             scaleForAlpha=0.1, alphaMin=10, poisRateNumClus=2, numSplitMergeMoves=3)
    iters = 30
    marks = []
    t0 = time.perf_counter()
    _, calls = chain.run_chain(eng, al, p.edge, p.clus, 100.0, cs, iters, seed=args.seed,
                               on_iter=lambda it, cv: marks.append((time.perf_counter(), cv.n_loglik_calls)))
    dt = time.perf_counter() - t0
    # steady state: from the end of the first iteration (after handle creation, first plans and graph captures)
    (ta, ca), (tb, cb) = marks[0], marks[-1]
    out["mcmc"] = {"iterations_per_sec": (iters - 1) / (tb - ta), "iterations": iters, "likelihood_calls": calls,
                   "likelihood_calls_per_sec": (cb - ca) / (tb - ta), "us_per_likelihood_call": (tb - ta) / (cb - ca) * 1e6,
                   "iterations_per_sec_with_setup": iters / dt, "how": how,
                   "note": "host driver = Python mirror of .performStepPhylo (R in a deployment); rate taken after the "
                           "first iteration, *_with_setup includes the initial logLikCpp"}
    if how != "single GPU":
        return out

    class Probe:
        """The same backend, reading dmpc_stats after each likelihood entry point (slows the driver: diagnostics only)."""
        KINDS = ("newBetweenTransProbsLogLik", "newWithinTransProbsLogLik", "withinClusNNIlogLik", "betweenClusNNIlogLik",
                 "clusSplitMergeLogLik")

        def __init__(self, be):
            self.be, self.rows = be, {k: [] for k in self.KINDS}

        def __getattr__(self, name):
            f = getattr(self.be, name)
            if name not in self.KINDS:
                return f

            def wrapped(ptr, *a, **kw):
                w0 = time.perf_counter()
                r = f(ptr, *a, **kw)
                w1 = time.perf_counter()
                st = self.be.stats(ptr)
                tl = self.be.root_timeline(ptr)
                self.rows[name].append(((w1 - w0) * 1e6, st["last_host_total_us"], st["last_host_plan_us"], st["last_host_submit_us"],
                                        st["last_eval_ms"] * 1e3, st["last_prune_ms"] * 1e3, st["last_stored_vertices"],
                                        st["last_fused_vertices"], st["last_walk_steps"], st["last_prune_launches"],
                                        tl["gather"] or 0.0, tl["exchange_or_reduce"] or 0.0, tl["certificate"] or 0.0,
                                        tl["gap_to_final"] or 0.0, tl["final_terms"] or 0.0, tl["final_exchange_or_reduce"] or 0.0))
                return r
            return wrapped

    pr = Probe(eng)
    chain.run_chain(pr, al, p.edge, p.clus, 100.0, cs, 6, seed=args.seed)
    cols = ("wall_us", "engine_host_us", "plan_us", "submit_us", "device_eval_us", "device_prune_us", "stored_vertices",
            "fused_vertices", "walk_steps", "prune_launches", "root_gather_us", "root_reduce_us", "root_certificate_us",
            "root_gap_us", "root_final_us", "root_final_reduce_us")
    out["mcmc"]["per_call"] = {k: dict(zip(cols, [round(float(x), 2) for x in np.mean(np.array(v), axis=0)]), n=len(v))
                               for k, v in pr.rows.items() if v}
    return out


def proposal_sweep(args):
    """BASELINE.json config 4: B split / merge proposals per dmpc_eval_proposals call on 5,000 seqs x 5,000 sites."""
    import torch
    from dmphyclus_b200 import build, synth
    from dmphyclus_b200.engine import Engine
    from dmphyclus_b200.phylo import Phylo
    from dmphyclus_b200.workloads import CONFIGS, Problem, load_packaged
    build.build()
    n, s, k, r = CONFIGS["config4"]
    r = args.sweep_rates
    data = args.data
    p = Problem(load_packaged(), n, s, k, r, data, args.tree, seed=args.seed)
    eng = Engine(device=0, fuse_tips=args.fuse)
    al = eng.getConvertedAlignment(list(synth.STATES), p.chars)
    W, B = p.W[4], p.B[4]
    h = eng.logLikCpp(p.edge, p.mrcas, p.pi, W, B, 1, al, n, s, 5, 5)["solutionPointer"]
    phy = Phylo(p.edge, n)
    cand = [("split", m) for m in p.mrcas if m > n]
    groups = {}
    for m in p.mrcas:
        groups.setdefault(int(phy.parent[m]), []).append(m)
    cand += [("merge", g[0], g[1]) for g in groups.values() if len(g) == 2]
    depth = phy.depth()
    rows = []
    for bsz in (64, 128, 256, 512, 1024, 2048, 4096):
        props = [cand[i % len(cand)] for i in range(bsz)]
        upd = sum(int(depth[q[1]] if q[0] == "split" else depth[q[1]] - 1) + 1 for q in props) * s * r
        eng.evalProposals(h, W, B, p.pi, props)
        eng.timer_begin(h)
        reps = 5
        for _ in range(reps):
            ll, nb = eng.evalProposals(h, W, B, p.pi, props)
        ms = eng.timer_end(h) / reps
        rows.append({"batch": bsz, "ms": ms, "proposals_per_s": bsz / ms * 1e3, "path_updates_per_s": upd / ms * 1e3, "batched": nb})
    # the same proposals one at a time through the reference-style entry point
    props = [cand[i % len(cand)] for i in range(64)]
    eng.timer_begin(h)
    for q in props:
        eng.clusSplitMergeLogLik(h, W, B, [q[1]] if q[0] == "split" else [q[1], q[2]], 1, p.pi)
        eng.RestorePreviousConfig(h, np.zeros((1, 2), np.int32), 5, 5, False, p.mrcas, True)
    one = eng.timer_end(h) / len(props)
    active = eng.stats(h)["last_active_tiles"]
    eng.finalDeallocate(h)
    emit(json.dumps({"committed_state_rescaled_tiles": active, "metric": "cluster proposals/sec", "unit": "proposals/s", "workload": f"config4: {n} seqs x {s} sites, {k} clusters, {r} rate categories, "
                     f"{len(cand)} distinct split/merge candidates", "alignment": data, "sweep": rows,
                     "one_by_one_ms_per_proposal": one, "one_by_one_proposals_per_s": 1e3 / one}))


def cpu_baseline(p, n, s, r, with_mcmc=True):
    from oracle.oracle import Oracle
    be, kind = cpu_reference_backend()
    ns, per_site = reference_sample_sites(be, p, s, 12.0)
    dt, _ = time_cpu(be, p, ns)
    memo = getattr(time_cpu, "dictionary", None)
    out = {"value": (n - 1) * ns * r / dt, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"first {ns} of {s} sites, all {n} tips, one logLikCpp ({dt:.1f} s of CPU work)",
           "host_cores": os.cpu_count(),
           "dictionary_entries": memo, "dictionary_hit_rate": (1.0 - memo / ((2 * n - 1) * ns)) if memo else None}
    # the dense oracle port: 1 thread, and OpenMP over sites on every host core (what numLikThreads advertised)
    nd = int(min(s, max(256, 4.0e8 // ((n - 1) * r))))        # ~8 s single-threaded, a few GB of host memory
    dt1, _ = time_cpu(Oracle(threads=1), p, nd)
    nthr = os.cpu_count() or 1
    dtn, _ = time_cpu(Oracle(threads=nthr), p, nd)
    out["port_dense_1thread"] = {"value": (n - 1) * nd * r / dt1, "unit": UNIT, "cores": 1, "sample": f"first {nd} of {s} sites"}
    out["port_dense_openmp"] = {"value": (n - 1) * nd * r / dtn, "unit": UNIT, "cores": nthr, "sample": f"first {nd} of {s} sites"}
    if with_mcmc:
        out["mcmc"] = reference_mcmc(p, s, per_site)
    return out


def main():
    # the ONE JSON line goes to the real stdout; anything libraries print to stdout meanwhile (NCCL's version banner)
    # is diverted to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global emit
    emit = lambda line: (real_stdout.write(line + "\n"), real_stdout.flush())
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.proposal_sweep:
        proposal_sweep(args)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
