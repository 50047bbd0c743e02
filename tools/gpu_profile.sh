#!/bin/bash
# ncu evidence for profiles/: the launch list of a short bench run (durations only) and one `--set full` capture of the
# k_prune / k_walk / k_root_gather launches of ONE evaluation.  Numbers printed under ncu are never bench values.
# Usage: gpurun --timeout 300 -- 'bash tools/gpu_profile.sh TAG'
tag=${1:-p}
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3"
timeout 100 $B > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || exit 1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  $B > gpurun_out/${tag}_ncu1.log 2>&1
# skip the create + warm-up evaluations (4 x (9 prune + 1 gather) + mask kernel), then capture one evaluation
timeout 150 ncu --set full --clock-control none --import-source on -k 'regex:k_prune|k_walk|k_root_gather' -s 60 -c 10 -f \
  -o gpurun_out/${tag}_full $B > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out/${tag}_*
