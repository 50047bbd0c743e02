#!/usr/bin/env python
"""Condenses what a gpurun call brought back into the small, tracked files under profiles/:

    python tools/ncu_summary.py --tag r01b --launches gpurun_out/launches2.csv --report gpurun_out/prof_prune2.ncu-rep \
        --bench gpurun_out/bench2.json

  <tag>_launches.csv   every launch of the `ncu --metrics gpu__time_duration.sum` pass (id, kernel, grid, ns)
  <tag>_kernel_share.md  per-kernel totals and shares of the profiled run, and of one evaluation
  <tag>_prune_metrics.csv  selected `--set full` metrics per captured launch (needs `ncu` here to read the report)
"""
import argparse
import csv
import json
import os
import subprocess
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("dmpc::", "")


def launches(path):
    hdr, rows = None, []
    for r in csv.reader(open(path)):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) == len(hdr) and r[0].isdigit():
            rows.append(r)
    ki, gi, vi, bi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Block Size")
    return [(int(r[0]), short(r[ki]), r[gi], r[bi], float(r[vi].replace(",", ""))) for r in rows]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--launches")
    ap.add_argument("--report")
    ap.add_argument("--bench")
    ap.add_argument("--note", default="")
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--data", default="evolved")
    a = ap.parse_args()
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)
    md = [f"# {a.tag}: kernel shares", "", a.note, ""]
    if a.launches:
        ls = launches(a.launches)
        with open(os.path.join(out, f"{a.tag}_launches.csv"), "w") as f:
            f.write("id,kernel,grid,block,duration_ns\n")
            for l in ls:
                f.write(f'{l[0]},{l[1]},"{l[2]}","{l[3]}",{l[4]:.0f}\n')
        tot = OrderedDict()
        for _, k, _, _, ns in ls:
            c, t = tot.get(k, (0, 0.0))
            tot[k] = (c + 1, t + ns)
        all_ns = sum(t for _, t in tot.values())
        md += [f"`ncu --metrics gpu__time_duration.sum --clock-control none` over `{os.path.basename(a.launches)}` "
               f"({len(ls)} launches; cold-cache, serialised: compare shares, not absolutes)", "",
               "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
        for k, (c, t) in tot.items():
            md.append(f"| {k} | {c} | {t / 1e3:.1f} | {100 * t / all_ns:.1f}% | {t / c / 1e3:.2f} |")
        # one evaluation = from one root reduction to the next (k_final when something rescaled, else k_root_gather)
        finals = [i for i, l in enumerate(ls) if l[1].startswith("k_final")]
        if len(finals) < 3:
            finals = [i for i, l in enumerate(ls) if l[1].startswith("k_root_gather")]
        if len(finals) >= 3:
            seg = ls[finals[1] + 1:finals[2] + 1]
            md += ["", "One evaluation (between two consecutive root reductions):", "",
                   "| kernel | grid | us |", "|---|---|---|"]
            for _, k, g, _, ns in seg:
                md.append(f"| {k} | {g} | {ns / 1e3:.2f} |")
            st = sum(x[4] for x in seg)
            pr = sum(x[4] for x in seg if x[1].startswith("k_prune"))
            md += ["", f"evaluation total {st / 1e3:.1f} us, of which k_prune {pr / 1e3:.1f} us = {100 * pr / st:.1f}%"]
    if a.bench and os.path.exists(a.bench):
        b = json.loads(open(a.bench).read().strip().splitlines()[-1])
        r = b["roofline"]
        md += ["", "Live (CUDA events, `bench.py`, not under ncu):", "",
               f"* value {b['value']:.4g} {b['unit']}, {b['ms_per_step']:.4f} ms/step; e2e {b['e2e']['value']:.4g} ({b['e2e']['ms_per_step']:.3f} ms/step)",
               f"* k_prune {r['kernel_ms_per_step']:.4f} ms/step = {100 * r['kernel_share_of_step']:.1f}% of the step; "
               f"achieved {r['achieved']:.2f} {r['unit']} = {r['frac']:.3f} of {r['peak']:.2f} ({r['bound']}; {r['peak_source'][:60]}...)",
               f"* clocks {b['clocks']}"]
    if a.report:
        raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr = rows[0]
        cols = ["ID", "Kernel Name", "Grid Size"] + [m for m in METRICS if m in hdr]
        idx = [hdr.index(c) for c in cols]
        with open(os.path.join(out, f"{a.tag}_prune_metrics.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(cols)
            w.writerow([rows[1][i] for i in idx])
            for r in rows[2:]:
                w.writerow([short(r[i]) if c == "Kernel Name" else r[i] for c, i in zip(cols, idx)])
        md += ["", f"`ncu --set full` capture: `{a.tag}_prune_metrics.csv` (per-launch DRAM bytes, throughput, occupancy)."]
        # average DRAM traffic per k_prune launch over ONE evaluation's worth of captured launches
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        ur, uw = scale[rows[1][ri]], scale[rows[1][wi]]
        pr = [r for r in rows[2:] if "k_prune" in r[ki]]
        if pr:
            tot = sum(float(r[ri]) * ur + float(r[wi]) * uw for r in pr)
            json.dump({"kernel": "k_prune", "workload": a.workload, "data": a.data, "launches_captured": len(pr),
                       "dram_bytes_per_launch": tot / len(pr), "dram_bytes_total": tot,
                       "source": f"ncu --set full, {os.path.basename(a.report)}"},
                      open(os.path.join(out, f"{a.tag}_traffic.json"), "w"), indent=1)
            md += [f"k_prune DRAM traffic: {tot / 1e6:.1f} MB over {len(pr)} captured launches = {tot / len(pr) / 1e6:.2f} MB per launch."]
    open(os.path.join(out, f"{a.tag}_kernel_share.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
