#!/bin/bash
# Copy the round-2 evidence from gpurun_out/ (scratch) into profiles/ (tracked): bench JSON lines, aggregated ncu launch lists
# and DRAM traffic, summaries of the --set full captures, micro-benchmarks, the SASS mnemonic table.
set -u
cd "$(dirname "$0")/.."
G=gpurun_out; P=profiles
line() { grep '^{' "$1" | tail -1; }
line $G/bench_r2.log > $P/bench_r2.json
for wl in xl512 l256 t2i512 jit256 pixnerd256 train256; do [ -f $G/bench_r2_$wl.log ] && line $G/bench_r2_$wl.log > $P/bench_r2_$wl.json; done
[ -f $G/bench_r2_reference_arm.log ] && line $G/bench_r2_reference_arm.log > $P/bench_r2_reference_arm.json
for n in 2 4 8; do [ -f $G/bench_r2_n$n.log ] && line $G/bench_r2_n$n.log > $P/bench_r2_n$n.json; done
HDR="# B200 (sm_100a), round 2, python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --torch-baseline none --no-hbm-kernels --profile
# (XL/16 256px, 512 CFG rows, one CUDA-graph-replayed sampling step); ncu --clock-control none --profile-from-start off;
# per-launch times are cold-cache and serialised at boost clock: compare SHARES with the live step (121-123 ms under sw_power_cap)"
{ echo "$HDR"; echo "# --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum; aggregated by profiles/agg_traffic.py"; python $P/agg_traffic.py $G/traffic_r2.csv; } > $P/traffic_r2_xl256.txt
{ echo "$HDR"; python $P/agg_launches.py $G/traffic_r2.csv 2>/dev/null || true; } > /dev/null
python - <<'PY' > profiles/launches_r2_xl256.txt
import collections, csv, re
rows = [r for r in csv.DictReader([l for l in open("gpurun_out/traffic_r2.csv") if not l.startswith("==")]) if r["Metric Name"] == "gpu__time_duration.sum"]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for r in rows:
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"]))
    agg[name][0] += 1; agg[name][1] += v; tot += v
print("# ncu launch list of one sampling step (round 2): python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --torch-baseline none --no-hbm-kernels --profile")
print("# B200, XL/16 256px, 512 CFG rows, CUDA-graph replay; cold-cache serialised launches at boost clock: compare SHARES (live step 121-123 ms at ~1.35 GHz under sw_power_cap)")
print(f"total {tot:.3f} ms over {sum(n for n, _ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.3f} ms {100 * t / tot:5.1f}%  n={n:4d} avg={t / n:8.4f} ms  {k[:110]}")
PY
{ echo "# ncu launch list of one training step (round 2, eager launches on ONE stream: DECO_B200_GRAPH=0 DECO_B200_WGRAD_STREAM=0): python bench.py --workload train256 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --torch-baseline none --profile"; python $P/agg_launches.py $G/launches_train_r2.csv; } > $P/launches_r2_train256.txt
for n in decoder_tc attention gemm_qkv gemm_proj; do
  [ -f $G/full_r2_${n}_raw.csv ] && { echo "# ncu --set full --clock-control none --import-source on, one launch of the sampling step (round 2); profiles/summarize_full.py"; python $P/summarize_full.py $G/full_r2_${n}_raw.csv; } > $P/full_r2_$n.txt
done
for f in tmem_bench_r2 decoder_bench_r2 dct_bench_r2 attn_pitch_r2 attn_bench_r2; do [ -f $G/$f.txt ] && cp $G/$f.txt $P/$f.txt; done
[ -f $G/dtc_trace.txt ] && { echo "# clock64 stamps of csrc/decoder_tc.cu (-DDECO_DTC_TRACE; CTA 0, slot 0; scripts/dtc_trace.py), 64 CFG rows, version with one issuer per slot before the 32-bit index / prefetch changes"; echo "# epilogue codes: 1 tile start, 2 x ready, 3 scale' ready, 4 loaded, 5 A stored, 6 released, 10/11/12/13 SiLU stage (MMA done / loaded / stored / released), 20/21/22 gate + next norm stage, 30 final linear done; issuer: 1xx arrival seen, 2xx critical group committed, 3xx extras issued"; head -130 $G/dtc_trace.txt; } > $P/dtc_trace_r2.txt
python $P/sass_summary.py > $P/sass_r2.txt
[ -f $G/smoke_r2.log ] && grep "^smoke" $G/smoke_r2.log > $P/smoke_r2.txt
ls -la $P | tail -40
