#!/bin/bash
# tuning helper: build libpdelab_b200 with one object recompiled under extra flags, for same-box A/B runs through
# PDB200_LIB (not part of the product).   tools/build_variant.sh NAME FILE.cu "-DFLAG=.. -maxrregcount .."
set -e
name=$1; src=$2; extra=$3
root=$(cd "$(dirname "$0")/.." && pwd)
c=$root/dune-pdelab_b200/csrc; b=$root/dune-pdelab_b200/build; out=$root/dune-pdelab_b200/lib/variants
mkdir -p $out $b/variants
obj=$b/variants/${name}_$(basename $src .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $extra -c $c/$src -o $obj
objs=""
for f in $b/*.o; do
  if [ "$(basename $f .o)" == "$(basename $src .cu)" ]; then objs="$objs $obj"; else objs="$objs $f"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libpdelab_b200_$name.so $objs -lcudart
echo $out/libpdelab_b200_$name.so
