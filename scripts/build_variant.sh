#!/bin/bash
# Build an experimental variant of the library (WAM kernels only recompiled) into variants/lib_<name>.so
# usage: scripts/build_variant.sh name "-DFOO=1 -DBAR=2"      (select at run time with GPMP2B_LIB=variants/lib_<name>.so)
set -e
cd "$(dirname "$0")/../gpmp2_b200/csrc"
mkdir -p ../../variants
name=$1; shift
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $* -DINST_IS_LIE=0 -DINST_D=7 -c -o ../../variants/inst_vec_7_$name.o kernels_inst.cu
objs=$(ls c_abi.o inst_vec_[1-6].o inst_lie_*.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_$name.so $objs ../../variants/inst_vec_7_$name.o -lcudart
cuobjdump -res-usage ../../variants/inst_vec_7_$name.o 2>/dev/null | grep -A1 "pk_.*VecOptILi7ELi3" | grep -E "Function|REG" | sed 's/Function \(_Z[0-9]*[a-z_]*\).*/\1/' | paste - - | cut -c1-60
rm ../../variants/inst_vec_7_$name.o
