#!/bin/bash
# dev helper: compile the library with extra -D flags into build/variants/lib<NAME>.so (select it with SMPC_LIB=<path>)
set -e
NAME=$1; shift
cd "$(dirname "$0")/../safe_mpc_b200/csrc"
OUT=../../build/variants; mkdir -p $OUT/$NAME
FL="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC"
for f in api kernels qp qp_f32 mlp_tc mlp_tc2 peaks; do [ -f $f.cu ] && nvcc $FL "$@" -c $f.cu -o $OUT/$NAME/$f.o & done; wait
nvcc -shared -o $OUT/lib$NAME.so $OUT/$NAME/*.o -lcudart
echo built $OUT/lib$NAME.so
