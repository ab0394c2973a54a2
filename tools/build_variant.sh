#!/bin/bash
# usage: tools/build_variant.sh <name> <source.cu> "<extra nvcc flags>"   ->  variants/lib<name>.so
# Recompiles ONE source with extra flags (e.g. -DCM3P_ATTN_PROF) and links it with the objects of the regular build.
set -e
cd "$(dirname "$0")/.."
python -c "from cm3p_b200 import build; build.build()"
mkdir -p variants
OBJ=variants/$1_$(basename $2 .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 128 $3 -c cm3p_b200/csrc/$2 -o $OBJ
OTHERS=$(ls cm3p_b200/csrc/build/*.o | grep -v "/$(basename $2 .cu).o")
nvcc -shared -o variants/lib$1.so $OBJ $OTHERS -gencode arch=compute_100a,code=sm_100a
echo variants/lib$1.so
