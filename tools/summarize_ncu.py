#!/usr/bin/env python3
"""Turn ncu exports into the small text summaries kept under profiles/.

    ncu -i X.ncu-rep --page raw --csv    > raw.csv
    ncu -i X.ncu-rep --page source --csv > source.csv
    python tools/summarize_ncu.py raw.csv source.csv > profiles/<name>.md
    python tools/summarize_ncu.py raw.csv --traffic profiles/ncu_traffic.json 23 profiles/<name>.md

The second form records the DRAM bytes of the captured gather-level and pair-level launches (one 4096-blob chunk) as
bytes per average batch_add launch for comb width 23; bench.py reports that number as roofline.traffic.
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


def traffic(raw, out, comb_width, source):
    """gather-level launch + the tree levels (the captured first pair level scaled by the additions of all of them)"""
    import json
    import os
    rows = list(csv.reader(open(raw)))
    hdr, data = rows[0], rows[2:]
    name_i, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    u_rd, u_wr = unit[rows[1][rd]], unit[rows[1][wr]]
    by = {}
    for r in data:
        key = "gather" if "GatherPolicy" in r[name_i] else "pair" if "PairPolicy" in r[name_i] else None
        if key and key not in by:
            by[key] = float(r[rd]) * u_rd + float(r[wr]) * u_wr
    groups = -(-4096 // comb_width)
    level_rows, pair_adds = (groups + 1) // 2, []
    while level_rows > 1:
        pair_adds.append(level_rows // 2)
        level_rows = (level_rows + 1) // 2
    tree = by["pair"] * sum(pair_adds) / pair_adds[0]
    launches = 1 + len(pair_adds)
    doc = json.load(open(out)) if os.path.exists(out) else {}
    doc[str(comb_width)] = {"dram_bytes_per_avg_launch": (by["gather"] + tree) / launches, "launches_per_chunk": launches,
                            "gather_launch_bytes": by["gather"], "first_pair_launch_bytes": by["pair"],
                            "dram_bytes_per_blob": (by["gather"] + tree) / 4096, "source": source}
    json.dump(doc, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps(doc[str(comb_width)]))


def main():
    if len(sys.argv) > 2 and sys.argv[2] == "--traffic":
        traffic(sys.argv[1], sys.argv[3], int(sys.argv[4]), sys.argv[5])
        return
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
