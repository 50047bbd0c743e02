for g in 24 32 64 128 1000; do
  for w in config2 config3 config3_shard8; do
    DMPC_GROUP_COST=$g timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/w16_g${g}_${w}.json 2>> gpurun_out/w15.err
  done
done
echo done
