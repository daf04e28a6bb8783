# ncu launch list of one training step on ONE stream (eager launches), aggregated per kernel: run under gpurun
set -u
O=gpurun_out; mkdir -p $O
PT="python bench.py --workload train256 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --torch-baseline none --profile"
export DECO_B200_GRAPH=0 DECO_B200_WGRAD_STREAM=0
$PT > $O/plain_train_r2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv \
    --log-file $O/launches_train_r2.csv $PT > $O/ncu_launches_train_r2.log 2>&1; echo "train launch list rc=$?"
python profiles/agg_launches.py $O/launches_train_r2.csv | head -50
