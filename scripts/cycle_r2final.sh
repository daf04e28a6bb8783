set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_final.log 2>&1; echo "gpu tests rc=$?"; tail -2 $O/tests_final.log
python __graft_entry__.py --smoke > $O/smoke_r2.log 2>&1; echo "smoke rc=$?"; grep "^smoke" $O/smoke_r2.log
python bench.py > $O/bench_r2.log 2>&1; echo "bench rc=$?"
python bench.py --workload train256 --no-cpu-baseline --torch-baseline none > $O/bench_r2_train256.log 2>&1; echo "train rc=$?"
python - <<'PY'
import json
for f in ('bench_r2','bench_r2_train256'):
    l=[x for x in open(f'gpurun_out/{f}.log') if x.startswith('{')]
    d=json.loads(l[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'].get('rel_l2', d['parity'].get('grad_rel_l2')), d['roofline']['frac'], d['roofline']['step_frac_of_peak'], d['clocks']['sm_mhz'], (d.get('full_iteration') or {}).get('ms'), (d.get('torch_gpu_baseline') or {}).get('compiled',{}).get('value') if d.get('torch_gpu_baseline') else None)
PY
