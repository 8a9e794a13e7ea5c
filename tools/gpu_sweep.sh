#!/bin/bash
# sweep the walk-loop tunables in one gpurun job: "steps refill trigger" triples as arguments after the tag
TAG=${1:-sweep}; shift
mkdir -p gpurun_out
: > gpurun_out/${TAG}.jsonl
for cfg in "$@"; do
  set -- $cfg
  HXR_WALK_STEPS=$1 HXR_REFILL_MIN=$2 HXR_LEAF_TRIGGER=$3 timeout 600 python bench.py --steps 2 --warmup 2 --spp 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'steps':$1,'refill':$2,'trigger':$3,'value':round(j['value'],1),'ms':{k:round(v,1) for k,v in j['kernel_ms_per_step'].items()}}))" | tee -a gpurun_out/${TAG}.jsonl
done
