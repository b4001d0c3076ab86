#!/bin/bash
# young/old split of the helpers' entries: correctness at 16^3, then timings at 64^3 for several ages / helper splits
mkdir -p gpurun_out; O=gpurun_out
GLSNS_TRSV_DEBUG=1 timeout 300 python tools/trsv_sweep.py 16 check > $O/e_check16.json 2> $O/e_check16.err; echo rc=$? >> $O/e_check16.err
if ! grep -q apply_err $O/e_check16.json; then echo "check failed"; tail -5 $O/e_check16.err; exit 1; fi
cat $O/e_check16.json; grep young $O/e_check16.err
run() { tag=$1; n=$2; shift 2; env "$@" timeout 400 python tools/trsv_sweep.py $n > $O/e_$tag.json 2> $O/e_$tag.err || tail -3 $O/e_$tag.err; echo "$tag $(cat $O/e_$tag.json)"; grep young $O/e_$tag.err | head -2; }
run y0_64 64 GLSNS_TRSV_YOUNG=0
run y2_64 64 GLSNS_TRSV_YOUNG=2 GLSNS_TRSV_DEBUG=1
run y1_64 64 GLSNS_TRSV_YOUNG=1 GLSNS_TRSV_DEBUG=1
run y4_64 64 GLSNS_TRSV_YOUNG=4 GLSNS_TRSV_DEBUG=1
run y2k1_64 64 GLSNS_TRSV_YOUNG=2 GLSNS_TRSV_YOUNG_HELPERS=1
run y2k3_64 64 GLSNS_TRSV_YOUNG=2 GLSNS_TRSV_YOUNG_HELPERS=3
run y2_32 32 GLSNS_TRSV_YOUNG=2
# ILU(0) factorisation with per-row descriptors: same factors (digest), time
for n in 32 64; do
  timeout 400 python tools/ilu_factor_check.py $n > $O/e_ilu_runs_$n.json 2> $O/e_ilu_runs_$n.err
  GLSNS_ILU_BY_PIVOT_ROWS=1 timeout 400 python tools/ilu_factor_check.py $n > $O/e_ilu_rows_$n.json 2> $O/e_ilu_rows_$n.err
  cat $O/e_ilu_runs_$n.json $O/e_ilu_rows_$n.json
done
