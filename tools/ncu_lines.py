"""Aggregate an ncu source page (`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`) per CUDA source line:
share of executed warp instructions and of stall samples.  Usage: python tools/ncu_lines.py file.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
cur, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and r[2] == "-":       # a CUDA source line (its SASS rows follow with an address)
        d = dict(zip(hdr[4:], r[4:]))
        lines.append((cur, int(r[0]), r[1].strip(), int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0), d))
ti = sum(l[3] for l in lines) or 1
ts = sum(l[4] for l in lines) or 1
print(f"total warp instructions {ti}, samples {ts}")
stall_keys = ["stall_barrier", "stall_long_sb", "stall_wait", "stall_mio", "stall_short_sb", "stall_math", "stall_lg", "stall_not_selected"]
for f, ln, src, ins, smp, d in lines:
    if 100 * ins / ti >= min_pct or 100 * smp / ts >= min_pct:
        top = sorted(((int(d.get(k, 0) or 0), k) for k in stall_keys), reverse=True)[:2]
        tops = " ".join(f"{k[6:]}={v}" for v, k in top if v)
        print(f"{100*ins/ti:5.1f}% inst {100*smp/ts:5.1f}% samp  {f}:{ln:<4d} {src[:90]}  [{tops}]")
