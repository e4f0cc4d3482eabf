#!/bin/bash
# round 2, GPU call A: full GPU test suite, the new bench workloads, reference arm, gym download attempt, sanitizers
set -u
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
( timeout 60 python -m pip download gym==0.26.2 --no-deps -d /tmp/gymdl > $O/pip_gym.log 2>&1; echo "rc=$?" >> $O/pip_gym.log ) 
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_ppo.json 2> $O/bench_ppo.err
timeout 600 python bench.py --workload pendulum --steps 5 --warmup 3 > $O/bench_pendulum.json 2> $O/bench_pendulum.err
timeout 600 python bench.py --workload scale1m --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_scale1m.json 2> $O/bench_scale1m.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 python bench.py --workload equiv --precision bf16 --steps 5 --warmup 3 > $O/bench_equiv_bf16.json 2> $O/bench_equiv_bf16.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py mlp > $O/memcheck_mlp.log 2>&1; echo "rc=$?" >> $O/memcheck_mlp.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py mlp > $O/racecheck_mlp.log 2>&1; echo "rc=$?" >> $O/racecheck_mlp.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py cnn > $O/memcheck_cnn.log 2>&1; echo "rc=$?" >> $O/memcheck_cnn.log
tail -3 $O/pytest_gpu.log; tail -2 $O/memcheck_mlp.log $O/racecheck_mlp.log $O/memcheck_cnn.log; cat $O/pip_gym.log | tail -3
