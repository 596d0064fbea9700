#!/bin/bash
# DRAM bytes + duration of every launch of one tensor-core kernel of the 4-task frame (one ncu pass with 3 metrics):
#   bash scripts/ncu_conv_traffic.sh [kernel-name-regex, default upconv_pair] [launches to skip] [launches to capture]
# Writes gpurun_out/conv_traffic.csv and the profiles-ready JSON gpurun_out/conv_traffic.json
mkdir -p gpurun_out
KERNEL=${1:-upconv_pair}; SKIP=${2:-60}; COUNT=${3:-60}
export KERNEL
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   -k regex:"$KERNEL" -s $SKIP -c $COUNT --csv --log-file gpurun_out/conv_traffic.csv \
   python bench.py --steps 6 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/conv_traffic.log 2>&1
python - <<'PY'
import csv, json, collections, os
rows = [r for r in csv.reader(open("gpurun_out/conv_traffic.csv")) if len(r) > 10 and r[0].isdigit()]
by_id = collections.defaultdict(dict)
for r in rows:
    by_id[r[0]][r[-3]] = (float(r[-1].replace(",", "")), r[-2])
def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
n = 0; rd = wr = dur = 0.0
for k, d in by_id.items():
    if "dram__bytes_read.sum" not in d: continue
    n += 1
    rd += to_bytes(*d["dram__bytes_read.sum"]); wr += to_bytes(*d["dram__bytes_write.sum"])
    v, u = d["gpu__time_duration.sum"]; dur += v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1e-3)
out = {"launches": n, "avg_dram_read_bytes": rd / max(n, 1), "avg_dram_write_bytes": wr / max(n, 1),
       "avg_dram_bytes": (rd + wr) / max(n, 1), "avg_duration_us_under_ncu": dur / max(n, 1),
       "kernel": os.environ["KERNEL"] + "_kernel", "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:%s over whole bench.py frames (scripts/ncu_conv_traffic.sh)" % os.environ["KERNEL"]}
json.dump(out, open("gpurun_out/conv_traffic.json", "w"), indent=1)
print(out)
PY
