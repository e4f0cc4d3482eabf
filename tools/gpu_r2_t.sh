#!/bin/bash
# round 2, call T: layer-wise tensor-core update for hidden 128 / 256
mkdir -p gpurun_out/r2t
timeout 600 python -m pytest tests/test_update_wide_gpu.py tests/test_update_generic_gpu.py -x -q -m gpu -s > gpurun_out/r2t/pytest_wide.log 2>&1; echo "wide tests rc=$?"
grep "rel L2" gpurun_out/r2t/pytest_wide.log; tail -5 gpurun_out/r2t/pytest_wide.log
timeout 300 python tools/bench_wide.py > gpurun_out/r2t/bench_wide.jsonl 2> gpurun_out/r2t/bench_wide.err; echo "bench rc=$?"
cat gpurun_out/r2t/bench_wide.jsonl; tail -3 gpurun_out/r2t/bench_wide.err
WIDE_ONLY=128 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 320 --csv --log-file gpurun_out/r2t/launches_wide128.csv python tools/bench_wide.py > gpurun_out/r2t/ncu_wide.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2t/launches_wide128.csv 2>/dev/null | head -30
