#!/bin/bash
# GPU call A (round 2, session 2): pipelined solver warp of the triangular solves vs the previous kernel
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python tools/trsv_sweep.py 16 check > $O/a_check16.json 2> $O/a_check16.err; echo rc=$? >> $O/a_check16.err
if ! grep -q apply_err $O/a_check16.json; then echo "check failed"; cat $O/a_check16.err | tail -5; exit 1; fi
cat $O/a_check16.json
for n in 32 64; do
  timeout 400 python tools/trsv_sweep.py $n > $O/a_sweep_new_$n.json 2> $O/a_sweep_new_$n.err
  GLSNS_LIB=$PWD/softx_2020_200_b200/libglsns_old.so timeout 400 python tools/trsv_sweep.py $n > $O/a_sweep_old_$n.json 2> $O/a_sweep_old_$n.err
done
GLSNS_TRSV_HELPERS=6 timeout 400 python tools/trsv_sweep.py 64 > $O/a_sweep_new_64_h6.json 2>&1
cat $O/a_sweep_*.json
timeout 600 python tools/trsv_trace.py 64 > $O/a_trace_new_64.json 2> $O/a_trace_new_64.err
timeout 1200 python -m pytest tests -m gpu -q -x > $O/a_gputests.log 2>&1; echo rc=$? >> $O/a_gputests.log
tail -3 $O/a_gputests.log
timeout 900 python bench.py --steps 2 --warmup 3 > $O/a_bench_n1.json 2> $O/a_bench_n1.err; echo rc=$? >> $O/a_bench_n1.err
