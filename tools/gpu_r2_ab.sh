#!/bin/bash
# round 2, call AB: layer-wise rollout with the two-plane / four-product actor GEMM
mkdir -p gpurun_out/r2ab
timeout 900 python -m pytest tests/test_rollout_gpu.py -x -q -m gpu -k "runtime_width" > gpurun_out/r2ab/pytest.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2ab/pytest.log
SKIP_SIMT=1 timeout 300 python tools/bench_wide_rollout.py > gpurun_out/r2ab/bench_wide_rollout.jsonl 2> gpurun_out/r2ab/bench_wide_rollout.err; echo "rollout bench rc=$?"
cat gpurun_out/r2ab/bench_wide_rollout.jsonl
