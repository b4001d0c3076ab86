#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/d_gputests.log 2>&1; echo rc=$? >> $O/d_gputests.log
tail -3 $O/d_gputests.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/d_bench_n1.json 2> $O/d_bench_n1.err; echo rc=$? >> $O/d_bench_n1.err
GLSNS_GMRES_LOOKAHEAD=0 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/d_bench_n1_nolook.json 2> $O/d_bench_n1_nolook.err
set -x
P="python tools/profile_kernels.py 64 spmv ilu_apply ilu_factor assemble_system"
$P > $O/r2p_plain64.log 2>&1 || exit 1
for K in spmv_ ilu_factor_runs_kernel assemble_cells; do
  SKIP=0; CNT=2
  if [ $K = assemble_cells ]; then SKIP=8; CNT=8; fi
  if [ $K = ilu_factor_runs_kernel ]; then SKIP=1; CNT=1; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o $O/r2p_${K}_n64 $P > $O/r2p_ncu_$K.log 2>&1
done
ls -la $O/r2p_*
