set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/tests_r2am.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/tests_r2am.log
python bench.py --workload train256 --no-cpu-baseline --torch-baseline none --steps 40 > $O/bench_r2_train256.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_r2_train256.log') if x.startswith('{')]
d=json.loads(l[-1]); print('train256', d['value'], d['ms_per_step'], d['parity']['grad_rel_l2'], d['full_iteration']['ms'], d['roofline']['frac'], d['roofline']['step_frac_of_peak'], d['clocks']['sm_mhz'])
PY
bash scripts/cycle_r2ad.sh 2>&1 | grep -E "total|gemm"
