#!/bin/bash
set -u
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
timeout 1500 python -m pytest tests/test_equiv_split_gpu.py -q -s > $O/pytest_split.log 2>&1; echo "rc=$?" >> $O/pytest_split.log
timeout 900 python -m pytest tests/test_equiv_gpu.py tests/test_plain_cnn_gpu.py tests/test_tc_gpu.py -q > $O/pytest_equiv.log 2>&1; echo "rc=$?" >> $O/pytest_equiv.log
timeout 600 python bench.py --workload equiv --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_equiv_fp32.json 2> $O/bench_equiv_fp32.err
timeout 600 python bench.py --workload cnn --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cnn_fp32.json 2> $O/bench_cnn_fp32.err
grep -c "passed\|failed" $O/pytest_split.log; tail -3 $O/pytest_split.log; tail -3 $O/pytest_equiv.log
