set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_baseline.py -x -q -s -k "pixnerd" > $O/tests_r2n.log 2>&1; echo "pixnerd tests rc=$?"; grep -E "pixnerd|PixNerd|passed|failed|Error|error|deco:" $O/tests_r2n.log | head -20
