#!/bin/bash
# parity subset, then A/B several KEY=VALUE settings on the bench: gpu_ab3.sh TAG "K=V" "K2=V2" ... ("none" = defaults)
TAG=$1; shift
mkdir -p gpurun_out; : > gpurun_out/${TAG}.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "${HXR_AB_TESTS:-walk_equals_brute_force or terrain_against_reference or layout or primary_hits}" 2>&1 | tail -3
for v in "$@"; do
  if [ "$v" = "none" ]; then E=""; else E="$v"; fi
  env $E timeout 600 python bench.py --steps 2 --warmup 2 --spp ${HXR_AB_SPP:-16} --no-cpu-baseline ${HXR_AB_ARGS} 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'env':'$v','value':round(j['value'],1),'ms':{k:round(x,1) for k,x in j['kernel_ms_per_step'].items()}}))" | tee -a gpurun_out/${TAG}.jsonl
done
