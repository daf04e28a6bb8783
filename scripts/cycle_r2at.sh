set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backward.py -q -x > $O/tests_r2at.log 2>&1; echo "backward tests rc=$?"; tail -2 $O/tests_r2at.log
python bench.py --workload train256 --no-cpu-baseline --torch-baseline none --steps 40 > $O/bench_train_at.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_train_at.log') if x.startswith('{')]
d=json.loads(l[-1]); print('train256', d['value'], d['ms_per_step'], d['parity']['grad_rel_l2'], d['full_iteration']['ms'], d['roofline']['step_frac_of_peak'], d['clocks']['sm_mhz'])
PY
