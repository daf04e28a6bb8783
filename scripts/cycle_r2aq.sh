set -u
O=gpurun_out; mkdir -p $O
PT="python bench.py --workload train256 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --torch-baseline none --profile"
export DECO_B200_GRAPH=0 DECO_B200_WGRAD_STREAM=0
for spec in "gemm_bf16_tcgen05_kernel<\(int\)256, \(int\)5:5:gemm_swiglu_dual" "gemm_bf16_tcgen05_kernel<\(int\)256, \(int\)4, \(int\)2, \(bool\)1, \(bool\)1:9:gemm_wgrad_tn" "gemm_bf16_tcgen05_kernel<\(int\)256, \(int\)0, \(int\)2:9:gemm_dgrad" "rmsnorm_modulate_bwd_kernel:7:norm_bwd" "headnorm_rope_bwd_kernel:5:headnorm_bwd" "attn_bwd_pipe_kernel:4:attn_bwd"; do
  K=${spec%%:*}; rest=${spec#*:}; S=${rest%%:*}; N=${rest#*:}
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:$K" -s $S -c 1 -f -o $O/full_r2t_$N $PT > $O/ncu_full_r2t_$N.log 2>&1; echo "full capture $N rc=$?"
  ncu -i $O/full_r2t_$N.ncu-rep --page raw --csv > $O/full_r2t_${N}_raw.csv 2>/dev/null
  rm -f $O/full_r2t_$N.ncu-rep
done
ls -la $O/full_r2t_*_raw.csv
