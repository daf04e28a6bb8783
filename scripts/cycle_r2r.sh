set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/tests_r2r.log 2>&1; echo "pytest rc=$?"; tail -3 $O/tests_r2r.log; grep -n "Error\|^E " $O/tests_r2r.log | head
B="python bench.py --torch-baseline none --no-cpu-baseline --no-hbm-kernels --no-e2e"
for i in 1 2; do
$B > $O/r_new.log 2>&1; echo "xl256 $(grep -o '"ms_per_step": [0-9.]*' $O/r_new.log) $(grep -o '"sm_mhz": [0-9]*' $O/r_new.log)"
done
python bench.py --workload train256 --no-cpu-baseline > $O/r_train.log 2>&1; echo "train rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r_train.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('train256', d['value'], d['ms_per_step'], d['parity'], d['optimizer']['iteration_with_optimizer_ms'])
else:
    print(open('gpurun_out/r_train.log').read()[-2500:])
PY
