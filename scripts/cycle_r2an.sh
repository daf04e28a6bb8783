set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/tests_r2an.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/tests_r2an.log
run() { # name, env..., args
  local name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --torch-baseline none --no-e2e --no-hbm-kernels $ARGS > $O/bench_$name.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('$name', d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])
else:
    print(open('gpurun_out/bench_$name.log').read()[-1500:])
PY
}
for rep in 1 2; do
ARGS="" ; run n1_rowmajor DECO_TILE_ORDER=rowmajor; run n1_snake DECO_TILE_ORDER=snake; run n1_snake_proj256 DECO_TILE_ORDER=snake DECO_STREAM_RING_K=1000
ARGS="--global-batch 32 --steps 60"; run b32_rowmajor DECO_TILE_ORDER=rowmajor; run b32_snake DECO_TILE_ORDER=snake; run b32_snake_proj256 DECO_TILE_ORDER=snake DECO_STREAM_RING_K=1000
done
