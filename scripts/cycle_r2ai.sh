set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_backward.py -q -x -k "headnorm or qknorm or norm_rope or gradients" > $O/tests_r2ai.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests_r2ai.log
python scripts/headnorm_bench.py 72; python scripts/headnorm_bench.py 64
