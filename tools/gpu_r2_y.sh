#!/bin/bash
# round 2, call Y: wide update after occupancy tuning + coalesced GEMM epilogue
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_update_wide_gpu.py tests/test_update_generic_gpu.py -x -q -m gpu > gpurun_out/r2y/pytest_wide.log 2>&1; echo "wide tests rc=$?"
tail -3 gpurun_out/r2y/pytest_wide.log
timeout 300 python tools/bench_wide.py > gpurun_out/r2y/bench_wide.jsonl 2> gpurun_out/r2y/bench_wide.err; echo "bench rc=$?"
grep "wide tc\|fused" gpurun_out/r2y/bench_wide.jsonl
WIDE_ONLY=128 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 320 --csv --log-file gpurun_out/r2y/launches_wide128.csv python tools/bench_wide.py > gpurun_out/r2y/ncu_wide.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2y/launches_wide128.csv 2>/dev/null | grep "aur::" | head -12
