set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backward.py -q -x > $O/tests_r2ae.log 2>&1; echo "backward tests rc=$?"; tail -3 $O/tests_r2ae.log
for c in 4 2 1 4 2 1; do
DECO_ROWS_BLOCK_CTAS=$c python bench.py --workload train256 --no-cpu-baseline --torch-baseline none --steps 40 > $O/bench_train_rb$c.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_train_rb$c.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('train256 rows_block_ctas=$c', d['value'], d['ms_per_step'], d['parity']['grad_rel_l2'], d['full_iteration']['ms'], d['clocks']['sm_mhz'])
else:
    print(open('gpurun_out/bench_train_rb$c.log').read()[-3000:])
PY
done
