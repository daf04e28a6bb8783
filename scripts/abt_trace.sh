#!/bin/bash
# Builds a scratch copy of the library with the attention-backward clock counters compiled in (-DABT_TRACE; run HERE, nvcc
# cross-compiles) -> deco_b200/_C/libdeco_trace.so; then on the GPU: DECO_B200_LIB=deco_b200/_C/libdeco_trace.so python scripts/abt_trace.py
set -e
cd "$(dirname "$0")/.."
python -c "from deco_b200 import build; build.build(verbose=False)"
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -DABT_TRACE -c deco_b200/csrc/attention_bwd_tc.cu -o deco_b200/_C/abt_trace.obj
objs=$(ls deco_b200/_C/*.o | grep -v attention_bwd_tc.o)
nvcc -shared -o deco_b200/_C/libdeco_trace.so $objs deco_b200/_C/abt_trace.obj -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64
echo built deco_b200/_C/libdeco_trace.so
