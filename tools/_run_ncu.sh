set -x
P="python tools/profile_kernels.py 64 spmv ilu_apply ilu_factor assemble_system"
$P > gpurun_out/r2p_plain64.log 2>&1 || exit 1
for K in spmv_groups_kernel trsv_team_kernel ilu_factor_runs_kernel assemble_cells; do
  SKIP=0; CNT=2
  if [ $K = assemble_cells ]; then SKIP=8; CNT=8; fi      # skip the set-up assembly, take the 8 colours of one timed assembly
  if [ $K = ilu_factor_runs_kernel ]; then SKIP=1; CNT=1; fi
  if [ $K = trsv_team_kernel ]; then SKIP=2; CNT=2; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/r2p_${K}_n64 $P > gpurun_out/r2p_ncu_$K.log 2>&1
done
B="python bench.py --cells 32 --steps 1 --warmup 1 --no-cpu-baseline"
$B > gpurun_out/r2p_plain_bench32.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3000 --csv --log-file gpurun_out/r2p_launches_n32.csv $B > gpurun_out/r2p_ncu_launches.log 2>&1
ls -la gpurun_out/r2p_*
