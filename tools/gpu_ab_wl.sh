#!/bin/bash
# A/B an environment knob on a secondary workload: gpu_ab_wl.sh TAG WORKLOAD SPP VAR val...
TAG=$1; WL=$2; SPP=$3; VAR=$4; shift 4
mkdir -p gpurun_out; : > gpurun_out/${TAG}.jsonl
for v in "$@"; do
  if [ "$v" = "unset" ]; then E=""; else E="$VAR=$v"; fi
  env $E timeout 600 python bench.py --workload $WL --steps 2 --warmup 2 --spp $SPP --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'$VAR':'$v','value':round(j['value'],1),'ms':{k:round(x,1) for k,x in j['kernel_ms_per_step'].items()}}))" | tee -a gpurun_out/${TAG}.jsonl
done
