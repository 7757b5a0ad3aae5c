#!/bin/bash
# Phase-instrumented build of the Pose2Vector kernel for dof 5 only (config 4): variants/lib_<name>.so
set -e
cd "$(dirname "$0")/../gpmp2_b200/csrc"
mkdir -p ../../variants
name=$1; shift
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DGPMP2B_PHASE_TIMING $*"
LISTS=("-DGPMP2B_DOF_LIST(X)=" "-DGPMP2B_LIE_DOF_LIST(X)=X(5)")
nvcc $F -DINST_IS_LIE=1 -DINST_D=5 -c -o ../../variants/il5_$name.o kernels_inst.cu &
nvcc $F "${LISTS[@]}" -c -o ../../variants/cabi_$name.o c_abi.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_$name.so ../../variants/cabi_$name.o ../../variants/il5_$name.o -lcudart
rm ../../variants/il5_$name.o ../../variants/cabi_$name.o
cuobjdump --dump-resource-usage ../../variants/lib_$name.so 2>/dev/null | grep -A1 "LieOptILi5ELi2EELi1EE" | grep -o "REG:[0-9]*\|STACK:[0-9]*\|SHARED:[0-9]*" | paste - - -
