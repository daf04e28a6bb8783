set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -q -x -k "gemm_stream or xl16 or fused_and_unfused or golden" > $O/tests_r2q.log 2>&1; echo "pytest rc=$?"; tail -3 $O/tests_r2q.log; grep -n "Error\|^E " $O/tests_r2q.log | head
B="python bench.py --torch-baseline none --no-cpu-baseline --no-hbm-kernels --no-e2e"
for i in 1 2; do
$B > $O/q_new.log 2>&1; echo "narrow-last-tile $(grep -o '"ms_per_step": [0-9.]*' $O/q_new.log) $(grep -o '"sm_mhz": [0-9]*' $O/q_new.log)"
done
P="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --torch-baseline none --no-hbm-kernels --profile"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $O/launches_r2q.csv $P > /dev/null 2>&1
python profiles/agg_launches.py $O/launches_r2q.csv | head -8
