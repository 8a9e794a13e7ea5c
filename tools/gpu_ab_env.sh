#!/bin/bash
# A/B environment settings on the bench: gpu_ab_env.sh TAG SPP "VAR=val VAR2=val" "..." (each argument is one configuration; "-" = defaults)
TAG=$1; SPP=${2:-32}; shift 2
mkdir -p gpurun_out; : > gpurun_out/${TAG}.jsonl
for v in "$@"; do
  if [ "$v" = "-" ]; then E=""; else E="$v"; fi
  env $E timeout 600 python bench.py --steps 2 --warmup 2 --spp $SPP --no-cpu-baseline ${HXR_AB_ARGS} 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'env':'$v','value':round(j['value'],1),'ms':{k:round(x,1) for k,x in j['kernel_ms_per_step'].items()}}))" | tee -a gpurun_out/${TAG}.jsonl
done
