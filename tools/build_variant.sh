#!/bin/sh
# Build an alternative libswiftover_cuda (kernel tuning experiments): tools/build_variant.sh NAME "-DFOO=1 ..."
# -> build/lib_NAME.so, used through SWO_LIB=build/lib_NAME.so (swiftover_b200/_lib.py).
set -e
cd "$(dirname "$0")/../swiftover_b200/csrc"
mkdir -p ../../build
make -s chain_table.o genome.o synth.o
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-pthread -I../../include -I. $2 -c swo_api.cu -o ../../build/swo_api_$1.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/lib_$1.so ../../build/swo_api_$1.o synth.o chain_table.o genome.o -lcudart
echo built build/lib_$1.so
