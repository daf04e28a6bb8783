#!/bin/bash
# One GPU measurement cycle (run under gpurun): parity tests, bench line, ncu launch list, one full capture of the top
# kernel.  Outputs land in gpurun_out/ (scratch); summaries worth judging are copied into profiles/ by hand.
#   scripts/gpu_cycle.sh [tag] [what]     what: any of "tests bench launches full gemm" (default: all)
set -u
TAG=${1:-r1}
WHAT=${2:-"tests bench launches full gemm"}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi_$TAG.txt 2>&1
BENCH_PROF="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --profile"
if [[ $WHAT == *tests* ]]; then
  python -m pytest tests -m gpu -x -q -s > $O/tests_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/tests_$TAG.log
  tail -3 $O/tests_$TAG.log
fi
if [[ $WHAT == *gemm* ]]; then
  python scripts/gemm_bench.py > $O/gemm_bench_$TAG.log 2>&1; echo "gemm_bench rc=$?"
fi
if [[ $WHAT == *bench* ]]; then
  python bench.py > $O/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -1 $O/bench_$TAG.log
fi
if [[ $WHAT == *launches* ]]; then
  $BENCH_PROF > $O/plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
      --log-file $O/launches_$TAG.csv $BENCH_PROF > $O/ncu_launches_$TAG.log 2>&1
  echo "launch list rc=$?"
fi
if [[ $WHAT == *full* ]]; then
  $BENCH_PROF > $O/plain2_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --profile-from-start off \
      --kernel-name-base demangled -k "regex:${NCU_KERNEL:-gemm_bf16_tcgen05_kernel<\(int\)256, \(int\)3}" -s ${NCU_SKIP:-5} -c 1 \
      -f -o $O/top_$TAG $BENCH_PROF > $O/ncu_full_$TAG.log 2>&1
  echo "full capture rc=$?"; tail -3 $O/ncu_full_$TAG.log
fi
