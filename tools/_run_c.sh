#!/bin/bash
# variants of the triangular solves at 64^3: "name:lib:env" triples
mkdir -p gpurun_out; O=gpurun_out
L=$PWD/softx_2020_200_b200
run() { # tag lib envs...
  tag=$1; lib=$2; shift 2
  env GLSNS_LIB=$L/$lib "$@" timeout 400 python tools/trsv_sweep.py ${N:-64} > $O/c_$tag.json 2> $O/c_$tag.err || tail -3 $O/c_$tag.err
  echo "$tag $(cat $O/c_$tag.json)"
}
for spec in "$@"; do
  IFS=: read tag lib envs <<< "$spec"
  run $tag $lib $(echo $envs | tr ',' ' ')
done
