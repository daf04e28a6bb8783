set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -s -k "decoder_tc or gemm_bias or timestep" > $O/tests_r2f_dec.log 2>&1; echo "decoder tests rc=$?"; grep -E "decoder_tc|passed|failed|Error|error" $O/tests_r2f_dec.log | head -30
python scripts/decoder_bench.py 512 256 > $O/decoder_bench_r2.txt 2>&1; echo "decoder bench rc=$?"; cat $O/decoder_bench_r2.txt | tail -8
python scripts/decoder_bench.py 64 256 >> $O/decoder_bench_r2.txt 2>&1; tail -7 $O/decoder_bench_r2.txt
