#!/bin/bash
# the update-side ncu --set full capture (the first capture's launch cap ran out before the update kernels)
set -u
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
timeout 300 python tools/profile_target.py update > $O/target_plain.log 2>&1; echo "target rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ppo_grad_tc|adv_moments_multi|grad_reduce|adam_kernel|pack_records|shuffle" -c 14 -o $O/upd_full python tools/profile_target.py update > $O/ncu_upd.log 2>&1; echo "upd full rc=$?"
ncu -i $O/upd_full.ncu-rep --page raw --csv > $O/upd_full_raw.csv 2> /dev/null
python tools/ncu_summary.py $O/upd_full.ncu-rep "update kernels (round 2)" > $O/upd_full_summary.md 2> $O/upd_full_summary.err
ncu -i $O/upd_full.ncu-rep --page source --csv --kernel-name regex:ppo_grad_tc --launch-skip 0 --launch-count 1 2> /dev/null | head -c 8000000 > $O/upd_full_source.csv
rm -f $O/upd_full.ncu-rep
ls -la $O
