#!/bin/bash
# Like build_variant.sh but with the clock64 phase instrumentation (GPMP2B_PHASE_TIMING) in the kernel and c_abi.
set -e
cd "$(dirname "$0")/../gpmp2_b200/csrc"
mkdir -p ../../variants
name=$1; shift
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DGPMP2B_PHASE_TIMING $*"
LISTS=("-DGPMP2B_DOF_LIST(X)=X(7)" "-DGPMP2B_LIE_DOF_LIST(X)=")
nvcc $F -DINST_IS_LIE=0 -DINST_D=7 -c -o ../../variants/iv7_$name.o kernels_inst.cu &
nvcc $F "${LISTS[@]}" -c -o ../../variants/cabi_$name.o c_abi.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_$name.so ../../variants/cabi_$name.o ../../variants/iv7_$name.o -lcudart
rm ../../variants/iv7_$name.o ../../variants/cabi_$name.o
