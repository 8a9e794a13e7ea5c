#!/bin/sh
# builds the C++ host tests against a given shared library (default: the product): tests/cpp/build.sh [lib.so] [outdir]
set -e
cd "$(dirname "$0")"
LIB=${1:-../../hexray_b200/libhexray_b200.so}
OUT=${2:-bin}
mkdir -p "$OUT"
DIR="$(dirname "$(realpath "$LIB")")"
# (-l: keeps the bare file name in DT_NEEDED, so the binary runs from any directory through its rpath)
g++ -std=c++17 -O2 -Wall test_multi_gpu.cpp -o "$OUT/test_multi_gpu" -L"$DIR" -l:"$(basename "$LIB")" -Wl,-rpath,"$DIR"
