#!/bin/bash
# A/B builds of the library on the bench: gpu_ab_lib.sh TAG SPP lib1.so lib2.so ... (paths; "tree" = the tree's own build)
TAG=$1; SPP=${2:-16}; shift 2
mkdir -p gpurun_out; : > gpurun_out/${TAG}.jsonl
LIB=hexray_b200/libhexray_b200.so
cp $LIB /tmp/tree.so
for v in "$@"; do
  if [ "$v" = "tree" ]; then cp /tmp/tree.so $LIB; else cp $v $LIB; fi
  timeout 600 python bench.py --steps 2 --warmup 2 --spp $SPP --no-cpu-baseline ${HXR_AB_ARGS} 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'lib':'$v','value':round(j['value'],1),'ms':{k:round(x,1) for k,x in j['kernel_ms_per_step'].items()}}))" | tee -a gpurun_out/${TAG}.jsonl
done
cp /tmp/tree.so $LIB
