#!/bin/bash
# launch list (time + DRAM bytes per launch) of the SECOND frame of a bundled scene at 1080p-class size, then a full capture
# of its longest k_walk launches. gpu_scene_ncu.sh TAG scene W H LAUNCHES_PER_FRAME
TAG=$1; S=$2; W=$3; H=$4; N=${5:-194}
mkdir -p gpurun_out
python tools/one_frame.py $S $W $H 2 > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain failed; tail -3 gpurun_out/${TAG}_plain.log; exit 1; }
cat gpurun_out/${TAG}_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size --clock-control none -s $N -c $N --csv --log-file gpurun_out/${TAG}_launches.csv python tools/one_frame.py $S $W $H 2 > gpurun_out/${TAG}_ncu.log 2>&1; echo "launchlist rc=$?"
