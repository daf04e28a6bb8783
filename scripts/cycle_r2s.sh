set -u
O=gpurun_out; mkdir -p $O
for m in 0 4 3 2; do
  echo "== poly mod $m"
  DECO_B200_LIB=build/libdeco_b200_poly$m.so python scripts/attn_pitch_bench.py 2>&1 | grep "pitch 80"
  DECO_B200_LIB=build/libdeco_b200_poly$m.so python scripts/attn_bench.py 2>&1 | cut -c1-80
  DECO_B200_LIB=build/libdeco_b200_poly$m.so timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" 2>&1 | tail -1
done
