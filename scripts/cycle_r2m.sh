set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/tests_r2m.log 2>&1; echo "pytest rc=$?"; tail -4 $O/tests_r2m.log; grep -n "Error\|^E " $O/tests_r2m.log | head -20
