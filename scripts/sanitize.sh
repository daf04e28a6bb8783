#!/bin/bash
# compute-sanitizer passes over __graft_entry__.smoke() (small shapes: denoiser forward, 4-step CFG Euler, DCT loss, one
# training step).  Run under gpurun; summaries land in gpurun_out/sanitizer_<tool>_<tag>.txt (copy into profiles/).
#   scripts/sanitize.sh [tag] [tools]      tools: any of "memcheck racecheck synccheck initcheck" (default: first three)
set -u
TAG=${1:-r2}
TOOLS=${2:-"memcheck racecheck synccheck"}
O=gpurun_out
mkdir -p $O
export DECO_B200_GRAPH=0      # eager launches: the sanitizer instruments kernels, not graph replays
for tool in $TOOLS; do
  log=$O/sanitizer_${tool}_$TAG.log
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 86 \
      python __graft_entry__.py --smoke > $log 2>&1
  rc=$?
  { echo "# compute-sanitizer --tool $tool python __graft_entry__.py --smoke   (exit code $rc; 86 = errors reported, 124 = timeout)";
    grep -E "^smoke:|ERROR SUMMARY|RACECHECK SUMMARY|Error:|Hazard|=========     at |Invalid|LEAK" $log | head -60; } > $O/sanitizer_${tool}_$TAG.txt
  echo "$tool rc=$rc"; tail -2 $O/sanitizer_${tool}_$TAG.txt
done
