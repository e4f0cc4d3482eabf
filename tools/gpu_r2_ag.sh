#!/bin/bash
# round 2, call AG: wide update after the TN-GEMM grid reorder
mkdir -p gpurun_out/r2ag
timeout 600 python -m pytest tests/test_update_wide_gpu.py tests/test_iteration_gpu.py -x -q -m gpu > gpurun_out/r2ag/pytest.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2ag/pytest.log
timeout 300 python tools/bench_wide.py > gpurun_out/r2ag/bench_wide.jsonl 2> gpurun_out/r2ag/bench_wide.err; echo "bench rc=$?"
grep "wide tc\|fused" gpurun_out/r2ag/bench_wide.jsonl | cut -c1-200
