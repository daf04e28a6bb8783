set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/tests_r2as.log 2>&1; echo "gpu tests rc=$?"; tail -4 $O/tests_r2as.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_backward.py -q -s -k "t2i_training_step_gradients_xxl" 2>&1 | grep "t2i global" | cut -c1-500
python bench.py --workload train256 --no-cpu-baseline --torch-baseline none --steps 40 > $O/bench_train_as.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_train_as.log') if x.startswith('{')]
d=json.loads(l[-1]); print('train256', d['value'], d['ms_per_step'], d['parity']['grad_rel_l2'], d['full_iteration']['ms'], d['roofline']['step_frac_of_peak'], d['clocks']['sm_mhz'])
PY
