#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; L=$PWD/softx_2020_200_b200
timeout 300 python tools/trsv_sweep.py 16 check > $O/f_check16.json 2> $O/f_check16.err; echo rc=$? >> $O/f_check16.err
if ! grep -q apply_err $O/f_check16.json; then echo "check failed"; tail -5 $O/f_check16.err; exit 1; fi
cat $O/f_check16.json
run() { tag=$1; n=$2; shift 2; env "$@" timeout 400 python tools/trsv_sweep.py $n > $O/f_$tag.json 2> $O/f_$tag.err || tail -3 $O/f_$tag.err; echo "$tag $(cat $O/f_$tag.json)"; }
run d3_64 64 A=1
run d2_64 64 GLSNS_LIB=$L/libglsns_d2.so
run d4_64 64 GLSNS_LIB=$L/libglsns_d4.so
run d3_32 32 A=1
run d4_32 32 GLSNS_LIB=$L/libglsns_d4.so
GLSNS_LIB=$L/libglsns_d4.so timeout 300 python tools/trsv_sweep.py 16 check
