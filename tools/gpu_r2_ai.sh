#!/bin/bash
# round 2, call AI: wide update vs the reference-generated golden vectors
mkdir -p gpurun_out/r2ai
timeout 600 python -m pytest tests/test_update_wide_gpu.py -x -q -m gpu -s -k "golden" > gpurun_out/r2ai/pytest.log 2>&1; echo "tests rc=$?"
grep "rel L2" gpurun_out/r2ai/pytest.log; tail -3 gpurun_out/r2ai/pytest.log
