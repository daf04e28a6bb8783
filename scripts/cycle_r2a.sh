set -u
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -s > $O/tests_r2a.log 2>&1; echo "pytest rc=$?"; tail -3 $O/tests_r2a.log
python bench.py > $O/bench_r2a.log 2>&1; echo "bench rc=$?"; tail -c 1500 $O/bench_r2a.log
P="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --torch-baseline none --profile"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $O/traffic_r2a.csv $P > $O/ncu_traffic_r2a.log 2>&1; echo "traffic rc=$?"
SAN_TIMEOUT=600 scripts/sanitize.sh r2a
