"""Summarise `ncu --page raw --csv` exports (made on the GPU box by scripts/gpu_round2_d.sh, because the .ncu-rep files
of ~80 kernels exceed the return limit) into a markdown table: one row per kernel NAME (median over its launches).

    python scripts/ncu_csv_summary.py gpurun_out/r2_ncu_hbm_stages.csv [more.csv ...] > profiles/r2_ncu_hbm_stages.md

HBM roofline figures: achieved DRAM GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration, against the
measured copy peak of MEASURED_PEAKS.json; note that ncu launches are cold-cache, serialised and ~40x replayed."""
import collections
import csv
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
         "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
COLS = [("dur us", "gpu__time_duration.sum"), ("DRAM rd MB", "dram__bytes_read.sum"), ("DRAM wr MB", "dram__bytes_write.sum"),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 hit %", "lts__t_sector_hit_rate.pct"), ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
        ("grid", "launch__grid_size"), ("block", "launch__block_size")]


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        out.append({h: (r[i], units[i]) for i, h in enumerate(hdr)})
    return out


def num(d, key):
    if key not in d:
        return None
    v, u = d[key]
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    return x * SCALE.get(u, 1.0)


def main():
    peak = 6569.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    by = collections.OrderedDict()
    for path in sys.argv[1:]:
        for d in load(path):
            name = d["Kernel Name"][0].split("(")[0].replace("void ", "").replace("vpb::", "")
            by.setdefault(name, []).append(d)
    print("| kernel | launches | " + " | ".join(c for c, _ in COLS) + " | DRAM GB/s | of measured HBM peak |")
    print("|---|---|" + "---|" * (len(COLS) + 2))
    for name, ds in by.items():
        cells = []
        for label, key in COLS:
            vals = [num(d, key) for d in ds]
            vals = [v for v in vals if v is not None]
            if not vals:
                cells.append("-")
                continue
            m = statistics.median(vals)
            if "MB" in label:
                m /= 1e6
            cells.append(f"{m:.2f}" if m < 100 else f"{m:.0f}")
        gbs = []
        for d in ds:
            t, rd, wr = num(d, "gpu__time_duration.sum"), num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
            if t and rd is not None and wr is not None:
                gbs.append((rd + wr) / t / 1e3)
        g = statistics.median(gbs) if gbs else 0.0
        print(f"| {name[:44]} | {len(ds)} | " + " | ".join(cells) + f" | {g:.0f} | {g / peak:.3f} |")


if __name__ == "__main__":
    main()
