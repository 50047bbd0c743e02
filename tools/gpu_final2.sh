#!/bin/bash
# second part of the round-2 evidence run: ncu --set full captures aligned on ONE full evaluation
o=gpurun_out/f2
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3"
$B > ${o}_plain.json 2> ${o}_plain.err && \
ncu --set full --clock-control none -k 'regex:k_prune|k_walk|k_root_gather|k_final' -s 58 -c 8 -f -o ${o}_full $B > ${o}_ncu2.log 2>&1
C2="python bench.py --workload config2 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3"
$C2 > ${o}_plain3.json 2>> ${o}_plain.err && \
ncu --set full --clock-control none -k 'regex:k_prune|k_walk|k_root_gather|k_final' -s 23 -c 6 -f -o ${o}_full_c2 $C2 > ${o}_ncu3.log 2>&1
python bench.py --proposal-sweep > ${o}_plain4.json 2>> ${o}_plain.err && \
ncu --set full --clock-control none -k 'regex:k_prune|k_root_gather|k_final' -s 22 -c 10 -f -o ${o}_full_batch python bench.py --proposal-sweep > ${o}_ncu4.log 2>&1
ls -la gpurun_out/f2_*
# the reports are tens of MB each (gpurun brings back at most 64 MiB): keep their raw pages as CSV only
for r in full full_c2 full_batch; do
  ncu -i ${o}_${r}.ncu-rep --page raw --csv > ${o}_${r}_raw.csv 2>/dev/null
  rm -f ${o}_${r}.ncu-rep
done
du -sh gpurun_out
