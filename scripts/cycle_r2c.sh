set -u
O=gpurun_out; mkdir -p $O
./build/tmem_bench > $O/tmem_bench_r2.txt 2>&1; echo "tmem_bench rc=$?"; tail -4 $O/tmem_bench_r2.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -s -k "decoder_tc" > $O/tests_r2c_dec.log 2>&1; echo "decoder tests rc=$?"; grep -E "decoder_tc|passed|failed|Error|error" $O/tests_r2c_dec.log | head -30
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/tests_r2c.log 2>&1; echo "pytest rc=$?"; tail -5 $O/tests_r2c.log
