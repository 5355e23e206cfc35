#!/bin/bash
# Build variants/libvda_<tag>.so: the in-tree objects with ONE source recompiled under extra nvcc flags.
#   tools/build_variant.sh <tag> <source.cu> [nvcc flags...]      (select it at run time with VDA_LIB=variants/libvda_<tag>.so)
set -e
cd "$(dirname "$0")/.."
tag=$1; src=$2; shift 2
pkg=video_depth_anything_b200
mkdir -p variants $pkg/build/variants
python -m $pkg.build > /dev/null
obj=$pkg/build/variants/${src%.cu}_$tag.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math -Xcompiler -fPIC -Xcompiler -O3 "$@" \
     -c $pkg/csrc/$src -o $obj
objs="$obj"
for o in $pkg/build/*.o; do
  [[ $(basename $o .o) == ${src%.cu} ]] || objs="$objs $o"
done
nvcc -shared -o variants/libvda_$tag.so $objs -gencode arch=compute_100a,code=sm_100a
echo variants/libvda_$tag.so
