#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for n in ${@:-16 32 64}; do
  timeout 400 python tools/ilu_factor_check.py $n > $O/ilu_cta_$n.json 2> $O/ilu_cta_$n.err; cat $O/ilu_cta_$n.json; tail -2 $O/ilu_cta_$n.err
done
