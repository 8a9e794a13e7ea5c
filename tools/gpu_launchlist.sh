#!/bin/bash
# launch list only (cheap): per-kernel device time of one bench frame. usage: gpu_launchlist.sh TAG SPP [extra bench args]
TAG=${1:-ll}; SPP=${2:-16}; shift 2
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --spp $SPP --no-cpu-baseline $@"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "launchlist rc=$?"
