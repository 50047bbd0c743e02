#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun --timeout 1700 -- 'bash tools/gpu_final.sh'): the driver's bench command lines,
# the other shapes, and the ncu captures profiles/ summarises.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
o=gpurun_out/f1
python bench.py --gpus 1 --steps 20 --warmup 5 > ${o}_bench_g1.json 2> ${o}_bench_g1.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > ${o}_bench_ref.json 2> ${o}_bench_ref.err
for shape in random balanced caterpillar; do
  python bench.py --workload config5 --tree $shape --steps 10 --warmup 3 --no-cpu-baseline --no-extras > ${o}_bench_config5_${shape}.json 2>> ${o}_misc.err
done
python bench.py --workload config4 --steps 20 --warmup 5 --no-cpu-baseline --mcmc-only > ${o}_bench_config4.json 2>> ${o}_misc.err
python bench.py --workload config2 --steps 50 --warmup 5 --no-cpu-baseline --mcmc-only > ${o}_bench_config2.json 2>> ${o}_misc.err
python bench.py --data iid --steps 5 --warmup 3 --no-cpu-baseline --no-extras > ${o}_bench_config3_iid.json 2>> ${o}_misc.err
python bench.py --proposal-sweep > ${o}_bench_sweep.json 2>> ${o}_misc.err
# ncu: launch list of a short config-3 run, then per-launch metrics of the pruning + root kernels of one evaluation
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3"
$B > ${o}_plain.json 2> ${o}_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file ${o}_launches.csv $B > ${o}_ncu1.log 2>&1
$B > ${o}_plain2.json 2>> ${o}_plain.err && \
ncu --set full --clock-control none --import-source on -k 'regex:k_prune|k_walk|k_root_gather|k_final' -s 40 -c 9 -f -o ${o}_full $B > ${o}_ncu2.log 2>&1
C2="python bench.py --workload config2 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3"
$C2 > ${o}_plain3.json 2>> ${o}_plain.err && \
ncu --set full --clock-control none -k 'regex:k_prune|k_walk|k_root_gather|k_final' -s 40 -c 8 -f -o ${o}_full_c2 $C2 > ${o}_ncu3.log 2>&1
python bench.py --proposal-sweep > ${o}_plain4.json 2>> ${o}_plain.err && \
ncu --set full --clock-control none -k 'regex:k_prune|k_root_gather|k_final' -s 3 -c 3 -f -o ${o}_full_batch python bench.py --proposal-sweep > ${o}_ncu4.log 2>&1
ls -la gpurun_out/f1_* | head -40
