set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backward.py -q -x -k "swiglu or gradients_toy or ragged" > $O/tests_r2z.log 2>&1; echo "backward tests rc=$?"; tail -3 $O/tests_r2z.log
python scripts/swiglu_fuse_bench.py 2>&1 | tee $O/swiglu_fuse_bench.txt
