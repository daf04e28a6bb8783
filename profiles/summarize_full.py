"""Key metrics of an `ncu --set full` capture: ncu -i x.ncu-rep --page raw --csv > x_raw.csv; python summarize_full.py x_raw.csv"""
import csv
import re
import sys

KEEP = [r"^Kernel Name$", r"^gpu__time_duration\.sum$", r"^gpc__cycles_elapsed\.max\.per_second$", r"^launch__grid_size$",
        r"^launch__block_size$", r"^launch__registers_per_thread$", r"^launch__shared_mem_per_block_dynamic$",
        r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$", r"^dram__bytes_read\.sum\.per_second$",
        r"^dram__bytes_write\.sum\.per_second$", r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$",
        r"^sm__pipe_tensor_subpipe_hmma_cycles_active\.avg\.pct_of_peak_sustained_active$",
        r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__inst_executed\.sum$", r"^sm__inst_executed_pipe_(alu|fma|lsu|tensor.*)\.sum$",
        r"^l1tex__data_bank_conflicts_pipe_lsu.*\.sum$", r"^smsp__warp_issue_stalled_.*_per_warp_active\.pct$",
        r"^lts__t_sector_hit_rate\.pct$", r"^sm__cycles_active\.avg$"]


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        stalls = []
        for h, u, v in zip(hdr, units, vals):
            if any(re.search(k, h) for k in KEEP) and v not in ("", "0"):
                if "warp_issue_stalled" in h:
                    stalls.append((float(v.replace(",", "")), h))
                    continue
                print(f"{h:90s} {v[:110]} {u}")
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f"{h:90s} {v:.2f} %")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
