#!/bin/bash
# same-box A/B of the fused elementwise kernels of the training step (DECO_B200_FUSE_GATE_NORM=0: one kernel per op)
for rep in 1 2; do for f in 0 1; do
  echo -n "train256 fuse_gate_norm=$f: "
  DECO_B200_FUSE_GATE_NORM=$f timeout 300 python bench.py --workload train256 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-parity 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'])"
done; done
