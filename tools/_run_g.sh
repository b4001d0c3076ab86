#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; L=$PWD/softx_2020_200_b200
run() { tag=$1; n=$2; shift 2; env "$@" timeout 400 python tools/trsv_sweep.py $n > $O/g_$tag.json 2> $O/g_$tag.err || tail -3 $O/g_$tag.err; echo "$tag $(cat $O/g_$tag.json)"; }
run nodeps_64 64 GLSNS_TRSV_NODEPS=1
run nodeps_d2_64 64 GLSNS_TRSV_NODEPS=1 GLSNS_LIB=$L/libglsns_d2.so
run deps_64 64 A=1
run nodeps_32 32 GLSNS_TRSV_NODEPS=1
run nodeps_h5_64 64 GLSNS_TRSV_NODEPS=1 GLSNS_TRSV_HELPERS=5
