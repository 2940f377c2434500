#!/bin/sh
# Builds the measurement helpers under tools/ (not product): the decompression-engine yardstick.
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -Xcompiler -fPIC -shared -cudart static -o libde_yardstick.so de_yardstick.cu
