#!/bin/bash
# Evidence run for profiles/: plain bench (32 spp per GPU step), launch list with DRAM bytes of every kernel, full ncu captures
# (with source) of one early and one deep bounce of every kernel - summarised ON THE BOX (the reports themselves exceed what
# gpurun brings back).   usage: gpu_evidence.sh TAG
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --spp 32 --no-cpu-baseline --no-extra"
echo "$CMD" > gpurun_out/${TAG}_cmd.txt
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
cut -c1-300 gpurun_out/${TAG}_plain.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "launchlist rc=$?"
for part in "early 10 9" "deep 46 8"; do
  set -- $part
  ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_shade|k_setup|k_resolve|k_finish|k_gen" -s $2 -c $3 -o /tmp/${TAG}_$1 $CMD > gpurun_out/${TAG}_ncu_$1.log 2>&1; echo "$1 rc=$?"
  python tools/ncu_summary.py /tmp/${TAG}_$1.ncu-rep > gpurun_out/${TAG}_$1_summary.txt 2>&1
  for k in "k_walk<(bool)0" "k_walk<(bool)1" "k_shade" "k_setup<(bool)0" "k_setup<(bool)1" "k_resolve_shadow" "k_finish_warp<(bool)0"; do
    n=$(echo "$k" | tr -c 'a-z0-9_' '_')
    python tools/ncu_sass.py /tmp/${TAG}_$1.ncu-rep "--name=$k" > gpurun_out/${TAG}_$1_sass_$n.txt 2>/dev/null
  done
done
ls -la gpurun_out | grep ${TAG}
