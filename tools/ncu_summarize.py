#!/usr/bin/env python
"""Turn the ncu outputs of tools/profile.sh into the small tracked files under profiles/.

    python tools/ncu_summarize.py c2 [c3 c5 ...]

  gpurun_out/launches_<w>.csv  -> profiles/r02_<w>_launches.csv (copy) + r02_<w>_launch_shares.txt
  gpurun_out/prof_<w>_raw.csv  -> profiles/r02_<w>_ncu_full_summary.csv and the per-kernel DRAM
                                  traffic table profiles/r02_ncu_traffic.json that bench.py reads
"""
import csv
import json
import os
import re
import shutil
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def short(name):
    """'void poly_flat_kernel<0, 1, 0>(const double *, ...)' -> 'poly_flat_kernel<0, 1, 0>'"""
    name = re.sub(r"^void\s+", "", name)
    depth, out = 0, []
    for ch in name:
        if ch == "(" and depth == 0:
            break
        depth += ch == "<"
        depth -= ch == ">"
        out.append(ch)
    return "".join(out).replace("dnlp::", "").strip()


def launches(w):
    src = os.path.join(OUT, "launches_%s.csv" % w)
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PROF, "r02_%s_launches.csv" % w))
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        t = float(r[v].replace(",", "")) * TIME.get(r[u], 1e-3)
        a = agg.setdefault(short(r[k]), [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, "r02_%s_launch_shares.txt" % w), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --device-only --steps 2 --warmup 3 (%s)\n" % w)
        f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's roofline.share_of_step\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-52s n=%4d total %9.1f us share %5.1f%% avg %8.1f us\n" % (name, n, t, 100 * t / total, t / n))


def full(w, traffic):
    src = os.path.join(OUT, "prof_%s_raw.csv" % w)
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    idx = {c: hdr.index(c) for c in COLS if c in hdr}
    kcol = hdr.index("Kernel Name")
    with open(os.path.join(PROF, "r02_%s_ncu_full_summary.csv" % w), "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["Kernel Name"] + list(idx))
        wr.writerow([""] + [units[i] for i in idx.values()])
        for r in rows[2:]:
            wr.writerow([r[kcol]] + [r[i] for i in idx.values()])
    per = traffic.setdefault(w, {})
    per.clear()
    for r in rows[2:]:
        def val(col, table):
            i = idx[col]
            return float(r[i].replace(",", "")) * table.get(units[i], 1.0)
        name = short(r[kcol])
        dram = val("dram__bytes_read.sum", UNIT) + val("dram__bytes_write.sum", UNIT)
        e = per.setdefault(name, {"dram_bytes_per_launch_max": 0.0, "duration": 0.0, "duration_unit": "us",
                                  "launches_captured": 0})
        e["launches_captured"] += 1
        if dram >= e["dram_bytes_per_launch_max"]:
            e["dram_bytes_per_launch_max"] = dram
            e["duration"] = val("gpu__time_duration.sum", TIME)


def main():
    tj = os.path.join(PROF, "r02_ncu_traffic.json")
    traffic = json.load(open(tj)) if os.path.exists(tj) else {}
    for w in sys.argv[1:]:
        launches(w)
        full(w, traffic)
    json.dump(traffic, open(tj, "w"), indent=1)


if __name__ == "__main__":
    main()
