#!/bin/bash
# One GPU-box pass: the GPU parity suite, the default bench line, and the caterpillar shape of config 5.
# Usage: gpurun --timeout 560 -- 'bash tools/gpu_validate.sh TAG'
tag=${1:-v}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpu.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 120 python bench.py --steps 200 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?" >> gpurun_out/${tag}_tests.log
timeout 120 python bench.py --workload config5 --tree caterpillar --steps 5 --warmup 3 --no-cpu-baseline --no-extras \
  > gpurun_out/${tag}_bench_c5cat.json 2> gpurun_out/${tag}_bench_c5cat.err
echo "c5cat rc=$?" >> gpurun_out/${tag}_tests.log
tail -4 gpurun_out/${tag}_tests.log
