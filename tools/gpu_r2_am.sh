#!/bin/bash
# round 2, call AM: ncu launch list of one hidden-128 iteration (bench.py --hidden_dim 128)
mkdir -p gpurun_out/r2am
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 5400 -c 1750 --csv --log-file gpurun_out/r2am/launches_ppo_hidden128.csv python bench.py --hidden_dim 128 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2am/ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2am/launches_ppo_hidden128.csv 2>/dev/null | grep "^|" | head -16 | cut -c1-170
