#!/bin/bash
# usage: profiles/build_variant.sh NAME "-DFLAG ..."   -> profiles/ab/libcfear_NAME.so (experiment build: Huber instantiations only)
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/cfear_radarodometry_code_public_b200/csrc
obj=/tmp/cfear_ab_$name; mkdir -p $obj $root/profiles/ab
common="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -DCFEAR_K5_MINIMAL $flags"
for f in cfear_b200 k5_cost0 k5_cost1 k5_cost2; do
  ( cd $src && nvcc $common -Xptxas -v -c -o $obj/$f.o $f.cu > $obj/$f.log 2>&1 ) &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/profiles/ab/libcfear_$name.so $obj/cfear_b200.o $obj/k5_cost0.o $obj/k5_cost1.o $obj/k5_cost2.o -lcudart
grep -h -A2 "k5_registerILi2ELi1ELb0\|k3_surface_points\|k1_kstrongestILb1ELb0" $obj/*.log | grep -E "Compiling|registers|spill" | sed 's/ptxas info    : //' | paste - - - | cut -c1-330
