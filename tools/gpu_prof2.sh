#!/bin/bash
# one gpurun job: plain run, launch list, then a full ncu capture (with source) of kernels matching REGEX
# usage: gpu_prof2.sh TAG REGEX SKIP COUNT SPP [extra bench args]
TAG=${1:-prof}; REGEX=${2:-k_shade}; SKIP=${3:-1}; COUNT=${4:-2}; SPP=${5:-16}; shift 5
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --spp $SPP --no-cpu-baseline $@"
echo "$CMD" > gpurun_out/${TAG}_cmd.txt
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "launchlist rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c $COUNT -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | grep ${TAG}
