#!/bin/bash
set -u
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 900 python -m pytest tests/test_equiv_split_gpu.py -q -s > $O/pytest_split.log 2>&1; echo "rc=$?" >> $O/pytest_split.log
for p in 1 2; do timeout 300 
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_equiv_split_gpu.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
tail -4 $O/pytest_split.log; tail -4 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/debug_conv0.log
