set -u
O=gpurun_out; mkdir -p $O
python scripts/attn_bwd_bench.py > $O/attn_bwd_bench_r2.txt 2>&1; cat $O/attn_bwd_bench_r2.txt
timeout 900 python -m pytest tests/test_gpu_backward.py -q -x > $O/tests_r2u.log 2>&1; echo "backward tests rc=$?"; tail -3 $O/tests_r2u.log
python bench.py --workload train256 --no-cpu-baseline --torch-baseline none > $O/bench_r2_train256.log 2>&1; echo "train rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_r2_train256.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('train256', d['value'], d['ms_per_step'], d['parity']['grad_rel_l2'], d['full_iteration']['ms'], d['roofline']['step_frac_of_peak'])
else:
    print(open('gpurun_out/bench_r2_train256.log').read()[-2000:])
PY
