#!/bin/bash
# tools/build_variant.sh <name> <nvcc -D flags...>: libglsns_<name>.so beside the product library (experiments; select with GLSNS_LIB)
name=$1; shift
D=softx_2020_200_b200/csrc; O=/tmp/glsns_var_$name; mkdir -p $O
for f in api assembly sparse trsv krylov comm host_mesh; do
  ext=cu; [ $f = host_mesh ] && ext=cpp
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fopenmp,-O3 -ccbin /usr/bin/g++ "$@" -x cu -c $D/$f.$ext -o $O/$f.o 2>&1 | grep -v deprecat &
done; wait
/usr/local/cuda/bin/nvcc -shared -o softx_2020_200_b200/libglsns_$name.so $O/*.o -Xcompiler -fopenmp -ccbin /usr/bin/g++ -lcudart -ldl -lgomp 2>&1 | grep -v deprecat
ls -la softx_2020_200_b200/libglsns_$name.so
