set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_r2af.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/tests_r2af.log
python bench.py --workload train256 --no-cpu-baseline --torch-baseline none > $O/bench_r2_train256.log 2>&1; echo "train rc=$?"
python - <<'PY'
import json
for f in ('bench_r2_train256',):
    l=[x for x in open(f'gpurun_out/{f}.log') if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, d['value'], d['ms_per_step'], d['parity']['grad_rel_l2'], d.get('full_iteration'), d['roofline']['frac'], d['roofline']['step_frac_of_peak'], d['clocks'])
    else:
        print(open(f'gpurun_out/{f}.log').read()[-2000:])
PY
