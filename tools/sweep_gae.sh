#!/bin/bash
# GAE tile-shape sweep (variants defined in csrc/gae.cu)
for v in 0 1 2 3 4 5 6; do
  echo "variant $v"; AUR_GAE_VARIANT=$v python tools/microbench.py gae 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if (d['T'],d['N']) in ((128,65536),(256,65536),(128,131072),(2048,131072)): print('  ',d['T'],d['N'],round(d['us_median'],1),'us',round(d['frac_of_measured_peak'],3))
"
done
