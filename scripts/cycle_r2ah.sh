set -u
O=gpurun_out; mkdir -p $O
PT="python bench.py --workload train256 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --torch-baseline none --profile"
export DECO_B200_GRAPH=0 DECO_B200_WGRAD_STREAM=0
for spec in "pixel_decoder_bwd_mma_kernel:0:decoder_bwd" "rmsnorm_modulate_bwd_kernel<true>:5:norm_bwd" "headnorm_rope_bwd_kernel:5:headnorm_bwd"; do
  K=${spec%%:*}; rest=${spec#*:}; S=${rest%%:*}; N=${rest#*:}
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:$K" -s $S -c 1 -f -o $O/full_r2_$N $PT > $O/ncu_full_r2_$N.log 2>&1; echo "full capture $N rc=$?"
  ncu -i $O/full_r2_$N.ncu-rep --page raw --csv > $O/full_r2_${N}_raw.csv 2>/dev/null
  ncu -i $O/full_r2_$N.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Registers|Theoretical Occ|Achieved Occ|Stall|stall|Warp Cycles|Issued|Eligible|No Eligible|L1/TEX Hit|L2 Hit|DRAM Throughput|Executed Ipc|Local" | head -40
done
