#!/bin/bash
# one gpurun job: plain run, then launch list, then a full ncu capture of the KD walk (same command line)
TAG=${1:-prof}; SKIP=${2:-2}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --spp 1 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "launchlist rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_walk -s $SKIP -c 2 -o gpurun_out/${TAG}_walk $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | tail -8
