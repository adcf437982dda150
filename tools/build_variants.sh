#!/bin/bash
# Build timing-ablation variants of libfa_b200.so in the dev container (nvcc cross-compiles; build/ travels with gpurun).
#   tools/build_variants.sh name1 "-DFLAG=1 ..." [name2 "flags" ...]   ->  build/var/libfa_<name>.so
# Only the translation units a flag can touch are recompiled per variant: FA_FUSED_* live in fa_api.cu; pass TU="a.cu b.cu"
# in the environment to recompile others.
set -e
cd "$(dirname "$0")/.."
CS=flash_attention_dlrs_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
mkdir -p build/var/obj/base
for f in $CS/*.cu; do
  o=build/var/obj/base/$(basename $f .cu).o
  if [ ! -f $o ] || [ -n "$(find $CS include -newer $o -type f | head -1)" ]; then
    nvcc $FLAGS -c -o $o $f &
  fi
done
wait
TU=${TU:-fa_api.cu}
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  mkdir -p build/var/obj/$name
  objs=""
  for f in $CS/*.cu; do
    b=$(basename $f .cu)
    if echo " $TU " | grep -q " $b.cu "; then
      nvcc $FLAGS $flags -c -o build/var/obj/$name/$b.o $f &
      objs="$objs build/var/obj/$name/$b.o"
    else
      objs="$objs build/var/obj/base/$b.o"
    fi
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/var/libfa_$name.so $objs
  echo "built build/var/libfa_$name.so ($flags)"
done
