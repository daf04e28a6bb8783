#!/bin/bash
# Round-2 evidence cycle (run under gpurun, 1 GPU): bench lines of every workload, ncu launch list + DRAM traffic of one
# sampling step, ncu --set full captures of the kernels this round changed, micro-benchmarks.  Raw outputs -> gpurun_out/;
# the summaries are copied into profiles/ by scripts/collect_profiles_r2.sh.
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_event_reasons.active --format=csv > $O/smi_r2.txt 2>&1
WHAT=${1:-"bench launches full micro"}
if [[ $WHAT == *bench* ]]; then
  python bench.py > $O/bench_r2.log 2>&1; echo "bench xl256 rc=$?"; tail -c 400 $O/bench_r2.log
  for wl in xl512 l256 t2i512 jit256 pixnerd256 train256; do
    python bench.py --workload $wl --no-cpu-baseline --torch-baseline none > $O/bench_r2_$wl.log 2>&1; echo "bench $wl rc=$?"
  done
  python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_r2_reference_arm.log 2>&1; echo "reference arm rc=$?"
fi
P="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --torch-baseline none --no-hbm-kernels --profile"
if [[ $WHAT == *launches* ]]; then
  $P > $O/plain_r2.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 400 --csv \
      --log-file $O/traffic_r2.csv $P > $O/ncu_traffic_r2.log 2>&1; echo "launch list + traffic rc=$?"
  PT="python bench.py --workload train256 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --torch-baseline none --profile"
  # one stream (DECO_B200_WGRAD_STREAM=0) so that every launch is timed alone
  DECO_B200_GRAPH=0 DECO_B200_WGRAD_STREAM=0 $PT > $O/plain_train_r2.log 2>&1 &&
  DECO_B200_GRAPH=0 DECO_B200_WGRAD_STREAM=0 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv \
      --log-file $O/launches_train_r2.csv $PT > $O/ncu_launches_train_r2.log 2>&1; echo "train launch list rc=$?"
fi
if [[ $WHAT == *full* ]]; then
  for spec in "pixel_decoder_tc_kernel:0:decoder_tc" "attention_tc_kernel:5:attention" "gemm_fused_kernel<\(int\)224:5:gemm_qkv" "gemm_fused_kernel<\(int\)192:5:gemm_proj"; do
    K=${spec%%:*}; rest=${spec#*:}; S=${rest%%:*}; N=${rest#*:}
    ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
        -k "regex:$K" -s $S -c 1 -f -o $O/full_r2_$N $P > $O/ncu_full_r2_$N.log 2>&1; echo "full capture $N rc=$?"
    ncu -i $O/full_r2_$N.ncu-rep --page raw --csv > $O/full_r2_${N}_raw.csv 2>/dev/null
  done
fi
if [[ $WHAT == *micro* ]]; then
  ./build/tmem_bench > $O/tmem_bench_r2.txt 2>&1
  python scripts/decoder_bench.py 512 256 > $O/decoder_bench_r2.txt 2>&1; python scripts/decoder_bench.py 64 256 >> $O/decoder_bench_r2.txt 2>&1
  python scripts/dct_bench.py > $O/dct_bench_r2.txt 2>&1
  python scripts/attn_pitch_bench.py > $O/attn_pitch_r2.txt 2>&1
  python scripts/attn_bench.py > $O/attn_bench_r2.txt 2>&1
fi
echo done
