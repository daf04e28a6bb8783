set -u
O=gpurun_out; mkdir -p $O
for fs in 0 fwd 1 0 fwd 1; do
DECO_B200_FUSE_SWIGLU=$fs python bench.py --workload train256 --no-cpu-baseline --torch-baseline none --steps 40 > $O/bench_train_fs$fs.log 2>&1; echo "train fs=$fs rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_train_fs$fs.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('train256 fs=$fs', d['value'], d['ms_per_step'], d['parity']['grad_rel_l2'], d['full_iteration']['ms'], d['roofline']['step_frac_of_peak'], d['clocks'])
else:
    print(open('gpurun_out/bench_train_fs$fs.log').read()[-3000:])
PY
done
