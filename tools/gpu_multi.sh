#!/bin/bash
# multi-GPU evidence: bash tools/gpu_multi.sh N  (under gpurun --gpus N)
N=$1
o=gpurun_out/m${N}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29601 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > ${o}_bench_config3.json 2> ${o}_bench_config3.err
if [ "$N" = "2" ]; then
  $T --master-port 29602 bench.py --gpus $N --steps 10 --warmup 5 --no-cpu-baseline > ${o}_bench_config3_mcmc.json 2> ${o}_bench_config3_mcmc.err
  timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_multi_device_handle.py -x -q > ${o}_tests.log 2>&1
fi
$T --master-port 29603 bench.py --gpus $N --workload config5 --steps 10 --warmup 3 --no-extras > ${o}_bench_config5_random.json 2> ${o}_bench_config5.err
if [ "$N" = "8" ]; then
  $T --master-port 29604 bench.py --gpus $N --workload config5 --tree balanced --steps 10 --warmup 3 --no-extras > ${o}_bench_config5_balanced.json 2>> ${o}_bench_config5.err
  $T --master-port 29605 bench.py --gpus $N --workload config5 --tree caterpillar --steps 5 --warmup 3 --no-extras > ${o}_bench_config5_caterpillar.json 2>> ${o}_bench_config5.err
  $T --master-port 29606 bench.py --gpus $N --workload config2 --weak --steps 50 --warmup 5 --no-extras > ${o}_bench_config2_weak.json 2> ${o}_bench_config2_weak.err
  timeout 300 python -m pytest tests/test_multi_device_handle.py -x -q -k real_devices > ${o}_tests.log 2>&1
fi
ls -la gpurun_out/m${N}_*
