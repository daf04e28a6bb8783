#!/bin/bash
# The 8-GPU shard shape on one GPU: launch list of one sampling step at 32 images (64 CFG rows), next to the live step time.
O=gpurun_out; mkdir -p $O
P="python bench.py --global-batch 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --torch-baseline none --no-hbm-kernels --profile"
python bench.py --global-batch 32 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --torch-baseline none --no-hbm-kernels > $O/bench_b32_live.log 2>&1; tail -1 $O/bench_b32_live.log | cut -c1-300
$P > $O/plain_b32.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $O/launches_b32.csv $P > $O/ncu_b32.log 2>&1; echo "rc=$?"
