#!/bin/bash
mkdir -p gpurun_out/r2aj
timeout 600 python -m pytest tests/test_rollout_gpu.py -x -q -m gpu -k "wide_rollout_matches" > gpurun_out/r2aj/pytest.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/r2aj/pytest.log | cut -c1-200
