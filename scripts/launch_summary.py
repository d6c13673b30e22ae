"""Per-kernel totals from an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list.
Usage: python scripts/launch_summary.py X.csv [top]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows:
    if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = r[ki].split("(")[0]
    tot[name] += v; cnt[name] += 1
allv = sum(tot.values())
print(f"total {allv:.2f} ms over {sum(cnt.values())} launches")
print("| kernel | launches | ms | share |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"| {k} | {cnt[k]} | {v:.2f} | {100 * v / allv:.1f}% |")
