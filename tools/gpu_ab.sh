#!/bin/bash
# A/B one environment knob on the bench: gpu_ab.sh TAG VAR val1 val2 ...
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out; : > gpurun_out/${TAG}.jsonl
for v in "$@"; do
  if [ "$v" = "unset" ]; then E=""; else E="$VAR=$v"; fi
  env $E timeout 600 python bench.py --steps 2 --warmup 2 --spp 16 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'$VAR':'$v','value':round(j['value'],1),'ms':{k:round(x,1) for k,x in j['kernel_ms_per_step'].items()}}))" | tee -a gpurun_out/${TAG}.jsonl
done
