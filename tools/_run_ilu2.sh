#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
GLSNS_LIB=$PWD/softx_2020_200_b200/libglsns_ilu64.so timeout 300 python tools/ilu_factor_check.py 64 > $O/ilu_t64_64.json 2> $O/ilu_t64_64.err; cat $O/ilu_t64_64.json; tail -1 $O/ilu_t64_64.err
