set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -q -x -k "attention or heun or qkv or t2i or xl16" > $O/tests_r2p.log 2>&1; echo "pytest rc=$?"; tail -3 $O/tests_r2p.log; grep -n "Error\|^E " $O/tests_r2p.log | head
python scripts/attn_pitch_bench.py > $O/attn_pitch_r2b.txt 2>&1; cat $O/attn_pitch_r2b.txt
python scripts/attn_bench.py > $O/attn_bench_r2.txt 2>&1; cat $O/attn_bench_r2.txt
