#!/bin/bash
# Build an A/B variant of libsionna_b200.so: same objects as the in-tree build except ONE source recompiled with extra
# flags.   usage: tools/build_variant.sh out.so source.cu -DFLAG=...      (select it with SIONNA_B200_LIB=$PWD/out.so)
set -e
out=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
python -m sionna_b200.csrc.build > /dev/null
obj=/tmp/variant_$$.o
fmad=-fmad=false
case $src in ofdm_mimo.cu|channel.cu|frontend.cu) fmad=-fmad=true;; esac
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $fmad "$@" -c -o $obj $root/sionna_b200/csrc/$src
objs=$(ls $root/build/obj/*.o | grep -v "/${src%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a --shared -o $out $objs $obj
rm -f $obj
echo built $out
