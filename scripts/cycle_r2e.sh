set -u
O=gpurun_out; mkdir -p $O
./build/tmem_bench > $O/tmem_bench_r2.txt 2>&1; echo "tmem_bench rc=$?"; tail -24 $O/tmem_bench_r2.txt
timeout 1500 python -m pytest tests -m gpu -q -s > $O/tests_r2e.log 2>&1; echo "pytest rc=$?"; tail -8 $O/tests_r2e.log
