set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_backward.py -q -x -k "attention_backward" > $O/tests_r2t.log 2>&1; echo "attn bwd tests rc=$?"; tail -4 $O/tests_r2t.log; grep -n "Error\|^E \|deco:" $O/tests_r2t.log | head
