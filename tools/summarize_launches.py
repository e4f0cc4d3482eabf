"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> markdown table of per-kernel shares.
Usage: python tools/summarize_launches.py launches.csv "title" > profiles/xxx_summary.md"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    u = r[mu]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "ms": 1.0, "second": 1e3}.get(u, 1e-6)
    tot[r[kn]] += ms; cnt[r[kn]] += 1
T = sum(tot.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else 'ncu launch list'}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, not absolutes).")
print(f"Raw CSV: {sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]}.  {sum(cnt.values())} launches, {T:.3f} ms in total.\n")
print("| share | total ms | launches | avg us | kernel |\n|---:|---:|---:|---:|---|")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"| {100*tot[k]/T:.2f}% | {tot[k]:.3f} | {cnt[k]} | {1e3*tot[k]/cnt[k]:.1f} | `{k[:110]}` |")
