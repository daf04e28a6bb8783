set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/tests_r2j.log 2>&1; echo "pytest rc=$?"; tail -4 $O/tests_r2j.log
for wl in xl256 t2i512; do
python bench.py --workload $wl --torch-baseline none --no-cpu-baseline --no-hbm-kernels > $O/bench_r2j_$wl.log 2>&1; echo "bench $wl rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_r2j_$wl.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('$wl', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['rel_l2'], d['roofline']['frac'], d['clocks'])
else:
    print(open('gpurun_out/bench_r2j_$wl.log').read()[-1500:])
PY
done
DECO_B200_QKV_PITCH80=0 python bench.py --torch-baseline none --no-cpu-baseline --no-hbm-kernels --no-e2e > $O/bench_r2j_nopitch.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $O/bench_r2j_nopitch.log
