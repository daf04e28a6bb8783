set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_backward.py -q -x -s -k "t2i_training" > $O/tests_r2ar.log 2>&1; echo "t2i tests rc=$?"; tail -30 $O/tests_r2ar.log | cut -c1-400
