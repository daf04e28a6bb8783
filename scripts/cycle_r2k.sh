set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --torch-baseline none --no-cpu-baseline --no-hbm-kernels --no-e2e --no-parity"
for i in 1 2; do
$B > $O/k_base.log 2>&1; echo "base   $(grep -o '"ms_per_step": [0-9.]*' $O/k_base.log) $(grep -o '"sm_mhz": [0-9]*' $O/k_base.log)"
DECO_STREAM_RING_BN=192 $B > $O/k_192.log 2>&1; echo "w2 192 $(grep -o '"ms_per_step": [0-9.]*' $O/k_192.log) $(grep -o '"sm_mhz": [0-9]*' $O/k_192.log)"
DECO_STREAM_RING_K=0 $B > $O/k_noring.log 2>&1; echo "noring $(grep -o '"ms_per_step": [0-9.]*' $O/k_noring.log) $(grep -o '"sm_mhz": [0-9]*' $O/k_noring.log)"
done
nvidia-smi --query-gpu=power.draw,power.limit,clocks.sm,temperature.gpu --format=csv
