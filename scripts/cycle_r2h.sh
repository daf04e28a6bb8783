set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/tests_r2h.log 2>&1; echo "pytest rc=$?"; tail -4 $O/tests_r2h.log
python bench.py --torch-baseline eager --no-cpu-baseline > $O/bench_r2h.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_r2h.log') if x.startswith('{')][-1]
d=json.loads(l)
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','parity','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_frac_of_peak'])
PY
DECO_B200_FUSED_STEP=0 python bench.py --torch-baseline none --no-cpu-baseline --no-hbm-kernels --no-e2e > $O/bench_r2h_unfused.log 2>&1; tail -c 300 $O/bench_r2h_unfused.log | head -c 10; grep -o '"ms_per_step": [0-9.]*' $O/bench_r2h_unfused.log
DECO_B200_DECODER=legacy python bench.py --torch-baseline none --no-cpu-baseline --no-hbm-kernels --no-e2e > $O/bench_r2h_legacy.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $O/bench_r2h_legacy.log
