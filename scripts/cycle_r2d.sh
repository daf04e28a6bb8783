set -u
O=gpurun_out; mkdir -p $O
python scripts/decoder_bench.py 512 256 > $O/decoder_bench_r2.txt 2>&1; echo "decoder bench rc=$?"; cat $O/decoder_bench_r2.txt | tail -8
python scripts/decoder_bench.py 64 256 >> $O/decoder_bench_r2.txt 2>&1; tail -7 $O/decoder_bench_r2.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/tests_r2d.log 2>&1; echo "pytest rc=$?"; tail -5 $O/tests_r2d.log
python bench.py --torch-baseline eager > $O/bench_r2d.log 2>&1; echo "bench rc=$?"; tail -c 600 $O/bench_r2d.log
