set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backward.py -q -x > $O/tests_r2y.log 2>&1; echo "backward tests rc=$?"; tail -5 $O/tests_r2y.log
python scripts/swiglu_fuse_bench.py 2>&1 | tee $O/swiglu_fuse_bench.txt
for fs in 0 1; do
DECO_B200_FUSE_SWIGLU=$fs python bench.py --workload train256 --no-cpu-baseline --torch-baseline none > $O/bench_train_fs$fs.log 2>&1; echo "train fs=$fs rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_train_fs$fs.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('train256 fs=$fs', d['value'], d['ms_per_step'], d['parity']['grad_rel_l2'], d['full_iteration']['ms'], d['roofline']['step_frac_of_peak'], d['clocks'])
else:
    print(open('gpurun_out/bench_train_fs$fs.log').read()[-3000:])
PY
done
