#!/bin/bash
# quick parity subset, then A/B one environment knob on the bench: gpu_ab2.sh TAG VAR val1 val2 ...
TAG=$1; VAR=$2; shift 2
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "walk_equals_brute_force or terrain_against_reference or many_mesh or visible" 2>&1 | tail -3
bash tools/gpu_ab.sh $TAG $VAR "$@"
