#!/bin/bash
# What the host-side changes made after the last GPU run of round 2 still have to show on a B200 (they were measured on
# the host alone: profiles/r02_host_timings.md): the GPU parity suite, then the bench lines whose host share moved —
# config 3 (MCMC rate, end to end), config 2 (MCMC rate, end to end), config 5 end to end (host-bound before).
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_recheck_host.sh TAG'; compare with profiles/r02_bench_g1.json (52.4
# MCMC iterations/s, e2e 6.44 ms), r02_bench_config2.json (470 iterations/s, e2e 0.516 ms), r02_bench_e2e_config5.json (7.33 ms).
tag=${1:-h}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err
timeout 120 python bench.py --workload config2 --steps 50 --warmup 5 --no-cpu-baseline --mcmc-only > gpurun_out/${tag}_bench_config2.json 2> gpurun_out/${tag}_bench_config2.err
timeout 120 python bench.py --workload config5 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_config5.json 2> gpurun_out/${tag}_bench_config5.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/*_bench_config*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    m = d.get("mcmc") or {}
    print(f, "ms/step %.4f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "mcmc it/s", m.get("iterations_per_sec"))
PY
tail -3 gpurun_out/${tag}_tests.log
