#!/bin/bash
# round 2, call AA: persistent skinny GEMM behind the 256-wide update and the layer-wise rollout
mkdir -p gpurun_out/r2aa
timeout 900 python -m pytest tests/test_update_wide_gpu.py tests/test_rollout_gpu.py -x -q -m gpu -k "wide or runtime_width" > gpurun_out/r2aa/pytest.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2aa/pytest.log
timeout 300 python tools/bench_wide.py > gpurun_out/r2aa/bench_wide.jsonl 2> gpurun_out/r2aa/bench_wide.err; echo "bench rc=$?"
grep "wide tc\|fused" gpurun_out/r2aa/bench_wide.jsonl
SKIP_SIMT=1 timeout 300 python tools/bench_wide_rollout.py > gpurun_out/r2aa/bench_wide_rollout.jsonl 2> gpurun_out/r2aa/bench_wide_rollout.err; echo "rollout bench rc=$?"
cat gpurun_out/r2aa/bench_wide_rollout.jsonl
