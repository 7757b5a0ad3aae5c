#!/bin/bash
# Build an experimental variant of the library that holds only the Pose2MobileArm D = 5 kernels (and the planar D = 2
# ones) into variants/lib_<name>.so      usage: scripts/build_variant_lie.sh name "-DFOO=1 ..."   (GPMP2B_LIB selects it)
set -e
cd "$(dirname "$0")/../gpmp2_b200/csrc"
mkdir -p ../../variants
name=$1; shift
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
if [ ! -f ../../variants/c_abi_lie5.o ] || [ c_abi.cu -nt ../../variants/c_abi_lie5.o ]; then
  nvcc $F '-DGPMP2B_DOF_LIST(X)=X(2)' '-DGPMP2B_LIE_DOF_LIST(X)=X(5)' -c -o ../../variants/c_abi_lie5.o c_abi.cu
fi
nvcc $F $* -DINST_IS_LIE=1 -DINST_D=5 -c -o ../../variants/inst_lie_5_$name.o kernels_inst.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -pthread -o ../../variants/lib_$name.so ../../variants/c_abi_lie5.o inst_vec_2.o ../../variants/inst_lie_5_$name.o -lcudart
cuobjdump -res-usage ../../variants/inst_lie_5_$name.o 2>/dev/null | grep -A1 "pk_.*LieOptILi5ELi2" | grep -E "Function|REG" | sed 's/Function \(_Z[0-9]*[a-z_]*\).*/\1/' | paste - - | cut -c1-70
rm ../../variants/inst_lie_5_$name.o
