#!/bin/sh
# tests/emu/build_emu.sh — TEST INFRASTRUCTURE ONLY. Builds tests/emu/libhxr_emu.so: the host
# front-end + frame driver of the product linked against launch_emu.cpp instead of the CUDA
# kernels, for the CPU test tier. Not a product artefact and never loaded by hexray_b200.
#   build_emu.sh        -> libhxr_emu.so
#   build_emu.sh asan   -> libhxr_emu_asan.so with -fsanitize=address,undefined (run the CPU tier on it with
#                          HXR_EMU_LIB=tests/emu/libhxr_emu_asan.so LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0)
set -e
cd "$(dirname "$0")"
SRC=../../hexray_b200/csrc
OUT=libhxr_emu.so
OPT="-O2"
if [ "$1" = "asan" ]; then OUT=libhxr_emu_asan.so; OPT="-O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer"; fi
g++ -std=c++17 $OPT -fPIC -shared -DHXR_EMU -Wall -Wno-unused-function \
    $SRC/abi.cpp $SRC/renderer.cpp $SRC/multi.cpp $SRC/kd_device_build.cpp $SRC/host/scene.cpp $SRC/host/mesh.cpp $SRC/host/flatten.cpp \
    $SRC/host/bitmap.cpp $SRC/host/kdtree.cpp $SRC/host/cache.cpp launch_emu.cpp \
    -o $OUT -lz -lpthread
