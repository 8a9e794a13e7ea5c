#!/bin/bash
# A/B two builds of the library on the bench: gpu_ab_lib.sh TAG [SPP] — tools/bin/libhexray_b200_prev.so against the tree's build
TAG=$1; SPP=${2:-16}
mkdir -p gpurun_out; : > gpurun_out/${TAG}.jsonl
LIB=hexray_b200/libhexray_b200.so
cp $LIB /tmp/new.so
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "walk_equals_brute_force or terrain_against_reference or many_mesh" 2>&1 | tail -3
for v in prev new prev new; do
  if [ "$v" = "prev" ]; then cp tools/bin/libhexray_b200_prev.so $LIB; else cp /tmp/new.so $LIB; fi
  timeout 600 python bench.py --steps 2 --warmup 2 --spp $SPP --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'lib':'$v','value':round(j['value'],1),'setup_s':round(j['config']['setup_s'],1),'ms':{k:round(x,1) for k,x in j['kernel_ms_per_step'].items()}}))" | tee -a gpurun_out/${TAG}.jsonl
done
cp /tmp/new.so $LIB
