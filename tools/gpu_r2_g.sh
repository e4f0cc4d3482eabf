#!/bin/bash
# round 2 profiles: ncu launch list of the bench command, full captures of the hot kernels, equivariant split-mode kernel table
set -u
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_plain.json 2> $O/bench_plain.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 300 python tools/profile_target.py > $O/target_plain.log 2>&1; echo "target rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ppo_grad_tc|gae_bulk|rollout_tc|critic_values_tc|adv_moments_multi|grad_reduce|adam_kernel" -c 14 -o $O/mlp_full python tools/profile_target.py > $O/ncu_mlp.log 2>&1; echo "mlp full rc=$?"
EQUIV_B=1024 timeout 300 python tools/profile_equiv.py > $O/equiv_plain.log 2>&1; echo "equiv target rc=$?"
EQUIV_B=1024 EQUIV_PRECISION=fp32 timeout 1200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/equiv_fp32_kernels.csv python tools/profile_equiv.py > $O/ncu_equiv.log 2>&1; echo "equiv list rc=$?"
EQUIV_B=1024 EQUIV_PRECISION=bf16 timeout 1200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/equiv_bf16_kernels.csv python tools/profile_equiv.py > $O/ncu_equiv_bf16.log 2>&1; echo "equiv bf16 list rc=$?"
EQUIV_B=256 EQUIV_PRECISION=fp32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm|wgrad3x3" -c 8 -o $O/equiv_full python tools/profile_equiv.py > $O/ncu_equiv_full.log 2>&1; echo "equiv full rc=$?"
# summarise on the box and leave the (large) reports behind: gpurun_out/ is limited to 64 MiB
for r in mlp_full equiv_full; do
  if [ -f $O/$r.ncu-rep ]; then
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2> /dev/null
    python tools/ncu_summary.py $O/$r.ncu-rep "$r (round 2)" > $O/${r}_summary.md 2> $O/${r}_summary.err
    ncu -i $O/$r.ncu-rep --page source --csv --kernel-name regex:"ppo_grad_tc|conv_igemm" 2> /dev/null | head -c 6000000 > $O/${r}_source.csv
    rm -f $O/$r.ncu-rep
  fi
done
ls -la $O
