#!/bin/bash
# round 2, call U: hidden-128 rollout / value kernels on tcgen05, wide update after the head-kernel templating
mkdir -p gpurun_out/r2u
timeout 900 python -m pytest tests/test_rollout_gpu.py -x -q -m gpu -k "runtime_width or replay or golden or evaluate" > gpurun_out/r2u/pytest_rollout.log 2>&1; echo "rollout tests rc=$?"
tail -5 gpurun_out/r2u/pytest_rollout.log
timeout 300 python tools/bench_wide_rollout.py > gpurun_out/r2u/bench_wide_rollout.jsonl 2> gpurun_out/r2u/bench_wide_rollout.err; echo "rollout bench rc=$?"
cat gpurun_out/r2u/bench_wide_rollout.jsonl; tail -3 gpurun_out/r2u/bench_wide_rollout.err
timeout 600 python -m pytest tests/test_update_wide_gpu.py -x -q -m gpu > gpurun_out/r2u/pytest_wide.log 2>&1; echo "wide tests rc=$?"
tail -3 gpurun_out/r2u/pytest_wide.log
timeout 300 python tools/bench_wide.py > gpurun_out/r2u/bench_wide.jsonl 2> gpurun_out/r2u/bench_wide.err; echo "bench rc=$?"
grep "wide tc\|fused" gpurun_out/r2u/bench_wide.jsonl
