#!/bin/sh
# builds the helper programs under tools/bin (git-ignored; they travel to the GPU box with the snapshot)
set -e
cd "$(dirname "$0")"
mkdir -p bin
SRC=../hexray_b200/csrc/host
g++ -std=c++17 -O2 -w objgen.cpp $SRC/scene.cpp $SRC/mesh.cpp $SRC/flatten.cpp $SRC/bitmap.cpp $SRC/kdtree.cpp $SRC/cache.cpp -o bin/hxr_objgen -lz -lpthread
if [ "$1" = "all" ]; then
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo microbench.cu -o bin/microbench
fi
