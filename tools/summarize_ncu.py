#!/usr/bin/env python3
"""Turn ncu exports into the small text summaries kept under profiles/.

    ncu -i X.ncu-rep --page raw --csv    > raw.csv
    ncu -i X.ncu-rep --page source --csv > source.csv
    python tools/summarize_ncu.py raw.csv source.csv > profiles/<name>.md
"""
import collections
import csv
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def main():
    raw, src = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print("## per-launch metrics (ncu --set full, --clock-control none)\n")
    for r in data:
        print("### %s\n" % r[name_i][:110])
        for k in RAW_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("- `%s` = %s %s" % (k, r[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.05:
                    stalls.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        print("- warp stall cycles per issued instruction: " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)))
        print()
    if not src:
        return
    print("## dynamic SASS mix (ncu source page, warp-level instructions executed)\n")
    kern, agg, hdr2 = None, collections.OrderedDict(), None
    for row in csv.reader(open(src)):
        if row and row[0] == "Kernel Name":
            kern = row[1][:110]
            if kern in agg:
                kern = None  # same kernel captured again: keep the first
            else:
                agg[kern] = collections.Counter()
            hdr2 = None
            continue
        if row and row[0] == "Address":
            hdr2 = row
            iex = hdr2.index("Instructions Executed")
            continue
        if kern and hdr2 and len(row) > iex and row[0].startswith("0x"):
            op = row[1].split()
            o = (op[1] if op[0].startswith("@") else op[0]).rstrip(";")
            agg[kern][o] += int(row[iex])
    for k, c in agg.items():
        tot = sum(c.values())
        print("### %s\n\ntotal %d warp instructions\n" % (k, tot))
        for o, v in c.most_common(12):
            print("- %-20s %5.1f %%" % (o, 100.0 * v / tot))
        print()


if __name__ == "__main__":
    main()
