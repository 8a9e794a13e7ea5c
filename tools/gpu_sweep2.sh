#!/bin/bash
# sweep SAH build parameters (intersect cost / max leaf size) in one gpurun job
TAG=${1:-sweep2}; shift
mkdir -p gpurun_out
: > gpurun_out/${TAG}.jsonl
for cfg in "$@"; do
  set -- $cfg
  HXR_KD_INTERSECT_COST=$1 HXR_KD_MAX_LEAF=$2 timeout 600 python bench.py --steps 2 --warmup 2 --spp 16 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
kd=j['config']['kd']
print(json.dumps({'ci':$1,'maxleaf':$2,'value':round(j['value'],1),'ms':{k:round(v,1) for k,v in j['kernel_ms_per_step'].items()},'per_ray':{k:round(v,2) for k,v in j['roofline']['per_ray'].items()},'blocks':kd['nodes'],'refs':kd['tri_refs'],'depth':kd['max_depth'],'build_ms':round(kd['build_ms'])}))" | tee -a gpurun_out/${TAG}.jsonl
done
