#!/bin/bash
# tools/build_variant.sh <name> "<extra nvcc -D flags>"  -> abr_control_b200/variants/libabrb_<name>.so (N=6 only)
set -e
cd "$(dirname "$0")/../abr_control_b200/csrc"
mkdir -p ../variants build/var_$1
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden $2"
nvcc $F -DABRB_N=6 -c kernels.cu -o build/var_$1/k6.o &
nvcc $F -Xcompiler -fvisibility=default "-DABRB_N_LIST(X)=X(6)" -c api.cu -o build/var_$1/api.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libabrb_$1.so build/var_$1/k6.o build/var_$1/api.o -lcudart
echo built $1
