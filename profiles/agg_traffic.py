"""Per-kernel DRAM traffic and time from an ncu pass with
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv:  python agg_traffic.py file.csv
The last line (`#json {...}`) is the same table in machine-readable form: bench.py reads `roofline.traffic` from the
committed profiles/traffic_rN_<workload>.txt instead of carrying a constant."""
import collections
import csv
import json
import re
import sys

BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
MS = {"ns": 1e-6, "us": 1e-3, "ms": 1, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1}


def main(path):
    rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for r in rows:
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"]))
        v, u, m = float(r["Metric Value"].replace(",", "")), r["Metric Unit"], r["Metric Name"]
        per[name][m] += v * (BYTES[u] if m.startswith("dram") else MS[u])
        ids[name].add(r["ID"])
    gb = gn = gms = 0.0
    table = {}
    for n, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        k = len(ids[n])
        rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        print(f"{d['gpu__time_duration.sum']:8.3f} ms  n={k:3d}  DRAM per launch {(rd + wr) / k / 1e6:8.1f} MB "
              f"({rd / k / 1e6:6.0f} read / {wr / k / 1e6:6.0f} write)  {n[:80]}")
        table[n] = dict(launches=k, ms=d["gpu__time_duration.sum"], dram_bytes_per_launch=(rd + wr) / k,
                        read=rd / k, write=wr / k)
        if "gemm" in n:
            gb, gn, gms = gb + rd + wr, gn + k, gms + d["gpu__time_duration.sum"]
    if gn:
        print(f"# all GEMM kernels: {int(gn)} launches, {gms:.3f} ms, DRAM {gb / gn / 1e6:.1f} MB per launch on average")
    print("#json " + json.dumps(dict(gemm_launches=int(gn), gemm_ms=gms, gemm_dram_bytes_per_launch=(gb / gn) if gn else None,
                                     kernels=table)))


if __name__ == "__main__":
    main(sys.argv[1])
