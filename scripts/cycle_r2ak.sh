set -u
O=gpurun_out; mkdir -p $O
PT="python bench.py --workload train256 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --torch-baseline none --profile"
export DECO_B200_GRAPH=0 DECO_B200_WGRAD_STREAM=0
for spec in "rmsnorm_modulate_bwd_kernel:7:norm_bwd" "gate_residual_norm_kernel:7:gate_norm"; do
  K=${spec%%:*}; rest=${spec#*:}; S=${rest%%:*}; N=${rest#*:}
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:$K" -s $S -c 1 -f -o $O/full_r2_$N $PT > $O/ncu_full_r2_$N.log 2>&1; echo "full capture $N rc=$?"
  ncu -i $O/full_r2_$N.ncu-rep --page details 2>/dev/null | grep -E "^  [a-z_A-Z:]+.*\(|Duration|Throughput|Registers|Theoretical Occ|Achieved Occ|Warp Cycles|Eligible|No Eligible|Hit Rate|Executed Ipc|Block Limit|Grid Size|Block Size|Waves|stalled|Stall" | head -40
done
