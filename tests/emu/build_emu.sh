#!/bin/sh
# tests/emu/build_emu.sh — TEST INFRASTRUCTURE ONLY. Builds tests/emu/libhxr_emu.so: the host
# front-end + frame driver of the product linked against launch_emu.cpp instead of the CUDA
# kernels, for the CPU test tier. Not a product artefact and never loaded by hexray_b200.
set -e
cd "$(dirname "$0")"
SRC=../../hexray_b200/csrc
g++ -std=c++17 -O2 -fPIC -shared -DHXR_EMU -Wall -Wno-unused-function \
    $SRC/abi.cpp $SRC/renderer.cpp $SRC/multi.cpp $SRC/host/scene.cpp $SRC/host/mesh.cpp $SRC/host/flatten.cpp \
    $SRC/host/bitmap.cpp $SRC/host/kdtree.cpp $SRC/host/cache.cpp launch_emu.cpp \
    -o libhxr_emu.so -lz -lpthread
