#!/bin/bash
# usage: [FMAD=true] scripts/exp_build.sh <name> <extra nvcc flags...>  -> exp/<name>/libmxadv.so
# A/B builds of the fused kernels (mx_fused.cu) linked against the in-tree objects of the rest.
# Select at run time with MXADV_LIB=exp/<name>/libmxadv.so.  exp/ is git-ignored but travels to
# the GPU box.
set -e
name=$1; shift
src=/root/repo/maestroex_b200/csrc
out=/root/repo/exp/$name
mkdir -p $out
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 -lineinfo -fmad=${FMAD:-false} $ARCH -Xcompiler -fPIC -Xptxas -v "$@" -c -o $out/mx_fused.o $src/mx_fused.cu 2> $out/mx_fused.ptxas.log
nvcc $ARCH -shared -o $out/libmxadv.so $out/mx_fused.o $(ls $src/*.o | grep -v mx_fused.o)
rm -f $out/mx_fused.o
echo built $out/libmxadv.so
