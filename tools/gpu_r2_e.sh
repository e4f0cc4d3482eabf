#!/bin/bash
set -u
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 1500 python -m pytest tests/test_equiv_split_gpu.py -q -s > $O/pytest_split.log 2>&1; echo "rc=$?" >> $O/pytest_split.log
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_equiv_split_gpu.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_ppo.json 2> $O/bench_ppo.err
timeout 600 python bench.py --workload pendulum --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_pendulum.json 2> $O/bench_pendulum.err
timeout 600 python bench.py --workload equiv --steps 5 --warmup 3 > $O/bench_equiv_split.json 2> $O/bench_equiv_split.err
tail -3 $O/pytest_split.log; tail -4 $O/pytest_gpu.log
