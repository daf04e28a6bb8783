set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/tests_r2o.log 2>&1; echo "pytest rc=$?"; tail -4 $O/tests_r2o.log; grep -n "Error\|^E " $O/tests_r2o.log | head -20
python bench.py --workload train256 --no-cpu-baseline > $O/bench_r2o_train.log 2>&1; echo "train bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_r2o_train.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('train256', d['value'], d['ms_per_step'], d['e2e'], d['optimizer'])
else:
    print(open('gpurun_out/bench_r2o_train.log').read()[-2000:])
PY
python bench.py --workload pixnerd256 --no-cpu-baseline --torch-baseline eager > $O/bench_r2o_pixnerd.log 2>&1; echo "pixnerd bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_r2o_pixnerd.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('pixnerd256', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'], d['roofline']['frac'], d['torch_gpu_baseline'])
else:
    print(open('gpurun_out/bench_r2o_pixnerd.log').read()[-2000:])
PY
