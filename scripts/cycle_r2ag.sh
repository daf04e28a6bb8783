set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_r2ag.log 2>&1; echo "gpu tests rc=$?"; tail -4 $O/tests_r2ag.log; grep -h "baseline global grad" $O/tests_r2ag.log
timeout 600 python -m pytest tests/test_gpu_backward.py -q -s -k "baseline_training" 2>&1 | grep "baseline global" 
