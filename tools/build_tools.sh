#!/bin/sh
set -e
cd "$(dirname "$0")"
mkdir -p bin
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo microbench.cu -o bin/microbench
