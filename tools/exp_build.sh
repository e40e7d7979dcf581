#!/bin/bash
# Builds an experimental variant of the library next to the shipped one:
#   tools/exp_build.sh NAME "-DLCB_EXP_FOO -DLCB_EXP_BAR=3"  ->  lattice_cryptography_b200/liblcb200_exp_NAME.so
# Select it at run time with LCB200_LIB=<path> (lattice_cryptography_b200/_ffi.py).  The variants are git-ignored (*.so)
# but travel to the GPU box with gpurun.
set -e
NAME=$1; FLAGS=$2
cd "$(dirname "$0")/../lattice_cryptography_b200/csrc"
mkdir -p exp_$NAME
for f in api sampler ring ring_generic wire; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $FLAGS -c $f.cu -o exp_$NAME/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../liblcb200_exp_$NAME.so exp_$NAME/*.o
echo built ../liblcb200_exp_$NAME.so
