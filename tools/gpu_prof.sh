#!/bin/bash
# one gpurun job: plain run, launch list, DRAM traffic of every k_walk launch, then a full ncu capture of kernels
# matching REGEX — all with the same command line (the bench's defaults unless SPP is given)
TAG=${1:-prof}; REGEX=${2:-k_walk}; SKIP=${3:-2}; COUNT=${4:-2}; SPP=${5:-32}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --spp $SPP --no-cpu-baseline"
echo "$CMD" > gpurun_out/${TAG}_cmd.txt
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "launchlist rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_walk -c 400 --csv --log-file gpurun_out/${TAG}_walk_dram.csv $CMD > gpurun_out/${TAG}_ncu_dram.log 2>&1; echo "dram rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c $COUNT -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | grep ${TAG}
