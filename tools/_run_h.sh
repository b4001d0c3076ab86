#!/bin/bash
# validation + evidence for the shipped kernels (outputs kept small: ncu reports are turned into csv on the box)
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/h_gputests.log 2>&1; echo rc=$? >> $O/h_gputests.log
tail -3 $O/h_gputests.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/h_bench_n1.json 2> $O/h_bench_n1.err; echo rc=$? >> $O/h_bench_n1.err
GLSNS_GMRES_LOOKAHEAD=0 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/h_bench_n1_nolook.json 2> $O/h_bench_n1_nolook.err
B="python bench.py --cells 32 --steps 1 --warmup 3 --no-cpu-baseline"
$B > $O/h_plain_bench32.json 2> $O/h_plain_bench32.err && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3000 --csv --log-file $O/h_launches_n32.csv $B > $O/h_ncu_launches.log 2>&1
P="python tools/profile_kernels.py 64 spmv ilu_apply ilu_factor assemble_system"
$P > $O/h_plain64.log 2>&1 || exit 1
for K in trsv_team_kernel spmv_ ilu_factor_runs_kernel assemble_cells; do
  SKIP=0; CNT=1
  if [ $K = assemble_cells ]; then SKIP=8; CNT=2; fi
  if [ $K = ilu_factor_runs_kernel ]; then SKIP=1; CNT=1; fi
  if [ $K = trsv_team_kernel ]; then SKIP=2; CNT=2; fi
  timeout 600 ncu --set full --clock-control none -k regex:$K -s $SKIP -c $CNT -f -o /tmp/rep_$K $P > $O/h_ncu_$K.log 2>&1
  ncu -i /tmp/rep_$K.ncu-rep --page raw --csv > $O/h_ncu_raw_$K.csv 2>> $O/h_ncu_$K.log
  ls -la /tmp/rep_$K.ncu-rep >> $O/h_ncu_$K.log
done
du -sh $O
