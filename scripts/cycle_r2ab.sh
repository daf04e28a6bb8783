set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests_r2ab.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/tests_r2ab.log
python bench.py --no-cpu-baseline --torch-baseline none > $O/bench_r2ab.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('bench_r2ab',):
    l=[x for x in open(f'gpurun_out/{f}.log') if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, d['value'], d['ms_per_step'], d['parity']['rel_l2'], d['roofline']['frac'], d['roofline']['step_frac_of_peak'], d['clocks'], d['e2e']['value'])
    else:
        print(open(f'gpurun_out/{f}.log').read()[-2000:])
PY
