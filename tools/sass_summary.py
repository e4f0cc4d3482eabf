"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA / bulk-async use (B200_PROFILING.md), from
`cuobjdump -sass aur_ppo_b200/libaurppo.so`.  Usage: python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "aur_ppo_b200/libaurppo.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTCBAR", "HMMA", "FFMA2", "DFMA", "RED", "ATOM"]
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        for k in MN:
            if op == k or (k in ("RED", "ATOM") and op.startswith(k)):
                counts[cur][k] += 1
print("# SASS evidence per kernel (round 2)\n")
print(f"`cuobjdump -sass {lib}` (sm_100a), instruction counts per kernel.  UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st")
print("(TMEM), UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, UTCBAR = tcgen05.commit.\n")
print("| kernel | SASS instr | " + " | ".join(MN) + " |")
print("|---|---:|" + "---:|" * len(MN))
tot = collections.Counter()
for k, c in counts.items():
    if not any(c[m] for m in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP")):
        continue
    print(f"| `{k[:90]}` | {c['_total']} | " + " | ".join(str(c[m]) if c[m] else "" for m in MN) + " |")
    tot.update(c)
print(f"| **all kernels with tensor-core / TMA / bulk-copy instructions** | {tot['_total']} | " + " | ".join(str(tot[m]) if tot[m] else "" for m in MN) + " |")
others = [k for k, c in counts.items() if not any(c[m] for m in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP"))]
print(f"\n{len(others)} further kernels are plain SIMT (env reset, shuffle, moments, reductions, Adam, elementwise helpers, the SIMT cross-check kernels).")
