set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -s -k "decoder_tc" > $O/tests_r2g_dec.log 2>&1; echo "decoder tests rc=$?"; grep -E "decoder_tc|passed|failed|Error|error" $O/tests_r2g_dec.log | head -30
python scripts/decoder_bench.py 512 256 > $O/decoder_bench_r2.txt 2>&1; echo "decoder bench rc=$?"; cat $O/decoder_bench_r2.txt | tail -8
DECO_B200_LIB=build/libdeco_b200_trace.so python scripts/dtc_trace.py > $O/dtc_trace.txt 2>&1
