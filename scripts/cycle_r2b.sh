set -u
O=gpurun_out; mkdir -p $O
./build/tmem_bench > $O/tmem_bench_r2.txt 2>&1; echo "tmem_bench rc=$?"; cat $O/tmem_bench_r2.txt
python -m pytest tests -m gpu -x -q -s > $O/tests_r2b.log 2>&1; echo "pytest rc=$?"; tail -3 $O/tests_r2b.log
