"""Per-kernel shares, tensor-pipe activity and DRAM traffic of an ncu CSV with the four metrics of profiles/r1_equiv_kernels.csv."""
import csv
import sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn = h.index('Kernel Name'); mn = h.index('Metric Name'); mv = h.index('Metric Value'); mu = h.index('Metric Unit'); idc = h.index('ID')
per = defaultdict(dict)
for r in rows[hi + 1:]:
    per[(r[idc], r[kn])][r[mn]] = (float(r[mv].replace(',', '')), r[mu])
agg = defaultdict(lambda: [0, 0, 0, 0, 0])
for (i, k), m in per.items():
    t, u = m['gpu__time_duration.sum']; t *= {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(u, 1e-3)
    a = agg[k.split('(')[0][:40]]
    a[0] += t; a[1] += 1; a[2] += t * m['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'][0]
    for j, key in ((3, 'dram__bytes_read.sum'), (4, 'dram__bytes_write.sum')):
        v, u = m[key]; a[j] += v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1e-6)
T = sum(a[0] for a in agg.values())
print("| share | total us | launches | tensor-pipe active % (time-weighted) | dram MB | dram GB/s | kernel |\n|---:|---:|---:|---:|---:|---:|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"| {100*a[0]/T:.1f}% | {a[0]:.1f} | {a[1]} | {a[2]/a[0]:.1f} | {a[3]+a[4]:.1f} | {(a[3]+a[4])/a[0]*1e3:.0f} | `{k}` |")
