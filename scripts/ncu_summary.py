"""Summarise ncu --set full reports (run here, no GPU needed): python scripts/ncu_summary.py gpurun_out/ncu2_*.ncu-rep"""
import csv
import subprocess
import sys

KEYS = [
    ("dur us", "gpu__time_duration.sum"),
    ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("TMA ld MB", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"),
    ("xbar %", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.pct_of_peak_sustained_elapsed"),
    ("DRAM rd MB", "dram__bytes_read.sum"),
    ("DRAM wr MB", "dram__bytes_write.sum"),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("smem pipe %", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
]


def load(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            d[h] = (r[i], units[i])
        out.append(d)
    return out


def val(d, k):
    if k not in d:
        return None
    x, u = d[k]
    try:
        x = float(x.replace(",", ""))
    except ValueError:
        return x
    scale = {"Kbyte": 1e-3, "byte": 1e-6, "Gbyte": 1e3, "Mbyte": 1.0, "ns": 1e-3, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return x * scale


def main():
    print("| report | kernel | " + " | ".join(k for k, _ in KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for p in sys.argv[1:]:
        for d in load(p):
            name = d["Kernel Name"][0].split("(")[0].replace("void ", "").replace("vpb::", "")[:28]
            cells = []
            for _, k in KEYS:
                v = val(d, k)
                cells.append("-" if v is None else (f"{v:.1f}" if isinstance(v, float) else str(v)))
            print(f"| {p.split('/')[-1].replace('.ncu-rep', '')} | {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
