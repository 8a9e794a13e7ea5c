#!/bin/bash
# full ncu capture of kernels matching REGEX on a secondary workload
TAG=$1; REGEX=$2; SKIP=$3; COUNT=$4; shift 4
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline $@"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c $COUNT -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
