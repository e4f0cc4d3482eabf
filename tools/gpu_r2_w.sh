#!/bin/bash
# round 2, call W: layer-wise rollout for hidden 256 / deeper 128
mkdir -p gpurun_out/r2w
timeout 900 python -m pytest tests/test_rollout_gpu.py -x -q -m gpu -k "runtime_width" > gpurun_out/r2w/pytest_rollout.log 2>&1; echo "rollout tests rc=$?"
tail -15 gpurun_out/r2w/pytest_rollout.log
SKIP_SIMT=1 timeout 300 python tools/bench_wide_rollout.py > gpurun_out/r2w/bench_wide_rollout.jsonl 2> gpurun_out/r2w/bench_wide_rollout.err; echo "rollout bench rc=$?"
cat gpurun_out/r2w/bench_wide_rollout.jsonl; tail -3 gpurun_out/r2w/bench_wide_rollout.err
