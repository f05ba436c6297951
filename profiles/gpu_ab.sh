#!/bin/bash
# usage: gpu_ab.sh TAG name...   -- gpu_ab2.sh over profiles/ab/libcfear_<name>.so, first one = reference poses
tag=$1; shift
rm -f /tmp/ab2_ref.npz
libs=""; for n in "$@"; do libs="$libs profiles/ab/libcfear_$n.so"; done
bash profiles/gpu_ab2.sh $tag $libs
