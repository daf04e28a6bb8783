set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_path.py -x -q -k "dct" > $O/tests_r2l.log 2>&1; echo "dct tests rc=$?"; tail -3 $O/tests_r2l.log
python scripts/dct_bench.py > $O/dct_bench_r2.txt 2>&1; cat $O/dct_bench_r2.txt
