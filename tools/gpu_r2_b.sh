#!/bin/bash
set -u
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 900 python -m pytest tests/test_equiv_split_gpu.py -q -s > $O/pytest_split.log 2>&1; echo "rc=$?" >> $O/pytest_split.log
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_equiv_split_gpu.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --workload equiv --precision split --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_equiv_split.json 2> $O/bench_equiv_split.err
timeout 600 python bench.py --workload equiv --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_equiv_bf16.json 2> $O/bench_equiv_bf16.err
timeout 600 python bench.py --workload cnn --precision split --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cnn_split.json 2> $O/bench_cnn_split.err
tail -5 $O/pytest_split.log; tail -5 $O/pytest_gpu.log
