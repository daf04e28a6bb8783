set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests_r2aj.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests_r2aj.log
python scripts/headnorm_bench.py 72; python scripts/headnorm_bench.py 64
