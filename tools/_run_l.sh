#!/bin/bash
# end-of-session validation and evidence (outputs kept small)
mkdir -p gpurun_out; O=gpurun_out
for n in 16 32 64; do timeout 400 python tools/trsv_sweep.py $n $([ $n = 16 ] && echo check) > $O/l_sweep_$n.json 2> $O/l_sweep_$n.err; cat $O/l_sweep_$n.json; done
timeout 1200 python -m pytest tests -m gpu -q -x > $O/l_gputests.log 2>&1; echo rc=$? >> $O/l_gputests.log
tail -3 $O/l_gputests.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/l_bench_n1.json 2> $O/l_bench_n1.err; echo rc=$? >> $O/l_bench_n1.err
tail -c 300 $O/l_bench_n1.json
P="python tools/profile_kernels.py 64 spmv ilu_apply ilu_factor assemble_system"
$P > $O/l_plain64.log 2>&1 || exit 1
cat $O/l_plain64.log
for K in trsv_team_kernel ilu_factor_cta_kernel; do
  SKIP=0; CNT=1
  if [ $K = ilu_factor_cta_kernel ]; then SKIP=1; CNT=1; fi
  if [ $K = trsv_team_kernel ]; then SKIP=2; CNT=2; fi
  timeout 600 ncu --set full --clock-control none -k regex:$K -s $SKIP -c $CNT -f -o /tmp/rep_$K $P > $O/l_ncu_$K.log 2>&1
  ncu -i /tmp/rep_$K.ncu-rep --page raw --csv > $O/l_ncu_raw_$K.csv 2>> $O/l_ncu_$K.log
done
B="python bench.py --cells 32 --steps 1 --warmup 3 --no-cpu-baseline"
$B > $O/l_plain_bench32.json 2> $O/l_plain_bench32.err && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3000 --csv --log-file $O/l_launches_n32.csv $B > $O/l_ncu_launches.log 2>&1
timeout 600 python tools/trsv_trace.py 64 > $O/l_trace_64.json 2> $O/l_trace_64.err
du -sh $O
