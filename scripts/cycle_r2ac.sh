set -u
O=gpurun_out; mkdir -p $O
for r in 512 64; do for s in 1 2 1 2; do DECO_QKV_EPI_SETS=$s python scripts/qkv_bench.py $r; done; done 2>&1 | tee $O/qkv_bench.txt
for s in 1 2 1 2; do
DECO_QKV_EPI_SETS=$s python bench.py --no-cpu-baseline --torch-baseline none --no-e2e > $O/bench_sets$s.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sets$s.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('sets=$s', d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])
else:
    print(open('gpurun_out/bench_sets$s.log').read()[-2000:])
PY
done
