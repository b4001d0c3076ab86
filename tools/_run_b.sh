#!/bin/bash
# quick A/B of the triangular solves: check at 16^3, sweeps at 32^3 / 64^3 (+ optional trace)
mkdir -p gpurun_out; O=gpurun_out; T=${1:-b}
timeout 300 python tools/trsv_sweep.py 16 check > $O/${T}_check16.json 2> $O/${T}_check16.err; echo rc=$? >> $O/${T}_check16.err
if ! grep -q apply_err $O/${T}_check16.json; then echo "check failed"; tail -5 $O/${T}_check16.err; exit 1; fi
cat $O/${T}_check16.json
for n in 32 64; do
  timeout 400 python tools/trsv_sweep.py $n > $O/${T}_sweep_$n.json 2> $O/${T}_sweep_$n.err; cat $O/${T}_sweep_$n.json
done
if [ "$2" = trace ]; then timeout 600 python tools/trsv_trace.py 64 > $O/${T}_trace_64.json 2> $O/${T}_trace_64.err; fi
