#!/bin/bash
# one gpurun job: plain run, then launch list, then a full ncu capture of kernels matching REGEX (same command line)
TAG=${1:-prof}; REGEX=${2:-k_walk}; SKIP=${3:-2}; COUNT=${4:-2}; SPP=${5:-4}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --spp $SPP --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "launchlist rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c $COUNT -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | grep ${TAG}
