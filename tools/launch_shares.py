"""Summarise an `ncu --metrics ... --csv` launch list (one row per kernel launch and metric) per kernel name:
launches, mean duration, share of the summed device time, DRAM / L2 bytes per launch, tensor-pipe activity.

    python tools/launch_shares.py gpurun_out/launches.csv > profiles/launch_shares.md
"""
import csv, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 12 and r[0].isdigit()]
per = defaultdict(lambda: defaultdict(dict))
for r in rows:
    per[r[4]][int(r[0])][r[12]] = float(r[14].replace(",", "") or 0)
tot = sum(m.get("gpu__time_duration.sum", 0) for k in per.values() for m in k.values())
print("| kernel | launches | mean us (ncu, serialised, caches flushed) | share of time | DRAM read+write per launch | L2 bytes per launch | tensor pipe active (mean %) |")
print("|---|---|---|---|---|---|---|")
n_front = 0
for name, launches in sorted(per.items(), key=lambda kv: -sum(m.get("gpu__time_duration.sum", 0) for m in kv[1].values())):
    n = len(launches)
    t = sum(m.get("gpu__time_duration.sum", 0) for m in launches.values())
    dram = sum(m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0) for m in launches.values()) / n
    l2 = sum(m.get("lts__t_bytes.sum", 0) for m in launches.values()) / n
    tp = sum(m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) for m in launches.values()) / n
    print(f"| `{name.replace('void ', '').split('(')[0]}` | {n} | {t / n / 1000:.2f} | {t / tot:.3f} | {dram / 1024:.1f} KB | {l2 / 1e6:.2f} MB | {tp:.2f} |")
    if "front_kernel" in name:
        n_front = n
updates = sum(1 for name, l in per.items() if "head_kernel" in name for _ in l) / 1.5   # 1 head per critic-only, 2 per policy update
print(f"\n{sum(len(l) for l in per.values())} launches in the window (~{updates:.0f} updates: 7 launches per critic-only update, 14 per policy update)")
