#!/bin/bash
# round 2, call AC: whole-iteration parity test
mkdir -p gpurun_out/r2ac
timeout 900 python -m pytest tests/test_iteration_gpu.py -x -q -m gpu -s > gpurun_out/r2ac/pytest.log 2>&1; echo "tests rc=$?"
grep "iteration [12]:" gpurun_out/r2ac/pytest.log; tail -30 gpurun_out/r2ac/pytest.log | cut -c1-220
