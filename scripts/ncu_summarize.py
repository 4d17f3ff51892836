"""Condense `ncu --page raw --csv` / `--page source --csv` exports into a small text summary.

    python scripts/ncu_summarize.py gpurun_out/<tag> [more tags...] > profiles/<round>_ncu_summary.txt

For every tag: the headline counters of the captured launch (duration, DRAM bytes, DMMA / FP64
pipe utilisation, occupancy limiters, issue rate, L2 hit rate), the warp-stall breakdown, and the
ten SASS instructions with the most stall samples (needs -lineinfo at compile time).
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_active.avg",
]


def raw_summary(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        out.append("kernel: %s" % vals[hdr.index("Kernel Name")])
        d = dict((h, (units[i], vals[i])) for i, h in enumerate(hdr))
        for k in KEYS:
            if k in d:
                out.append("  %-80s %12s %s" % (k, d[k][1], d[k][0]))
        stalls = []
        for h, (u, v) in d.items():
            if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                try:
                    stalls.append((float(v), h.split("warp_issue_stalled_")[1].replace("_per_warp_active.pct", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        out.append("  warp stalls (% of active warps): " + ", ".join("%s %.1f" % (n, v) for v, n in stalls[:6]))
    return out


def source_summary(path, top=10):
    rows = list(csv.reader(open(path)))
    hi = None
    for i, r in enumerate(rows):
        if r and r[0] == "Address":
            hi = i
            break
    if hi is None:
        return ["  (no source page)"]
    hdr = rows[hi]
    ci, cs, cn = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("# Samples")
    body = [r for r in rows[hi + 1:] if len(r) > cn]
    tot = sum(int(r[cn] or 0) for r in body) or 1
    body.sort(key=lambda r: -int(r[cn] or 0))
    out = ["  top SASS by stall samples (of %d):" % tot]
    for r in body[:top]:
        out.append("    %5.1f%%  %s" % (100.0 * int(r[cn] or 0) / tot, r[ci].strip()[:100]))
    dmma = sum(1 for r in body if "DMMA" in r[ci])
    out.append("  SASS: %d instructions, %d DMMA" % (len(body), dmma))
    return out


def main():
    for tag in sys.argv[1:]:
        print("== %s" % tag)
        try:
            print("\n".join(raw_summary(tag + ".raw.csv")))
        except Exception as e:  # noqa: BLE001
            print("  raw page unreadable: %s" % e)
        try:
            print("\n".join(source_summary(tag + ".source.csv")))
        except Exception as e:  # noqa: BLE001
            print("  source page unreadable: %s" % e)
        print()


if __name__ == "__main__":
    main()
