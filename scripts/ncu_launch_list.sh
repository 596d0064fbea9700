#!/bin/bash
# Per-launch device times of one benchmark run (cold-cache, serialised: compare SHARES, not absolutes).
# NOTE: the skip count passes ~9000 launches (engine construction incl. the load-time weight composition) through ncu: ~8 min of box time.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 700 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, total {tot/1e3:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={v[0]:5d} total={v[1]/1e3:10.1f} us  avg={v[1]/v[0]/1e3:8.2f} us  share={100*v[1]/tot:5.1f}%")
PY
