#!/bin/bash
# A/B builds of the library on the bench incl. its extra scenes: gpu_ab_scenes.sh TAG lib1.so lib2.so ... ("tree" = the tree's own build)
TAG=$1; shift 1
mkdir -p gpurun_out; : > gpurun_out/${TAG}.txt
LIB=hexray_b200/libhexray_b200.so
cp $LIB /tmp/tree.so
for v in "$@"; do
  if [ "$v" = "tree" ]; then cp /tmp/tree.so $LIB; else cp $v $LIB; fi
  timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline ${HXR_AB_ARGS} 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(j['value'],1), round(j['roofline']['frac'],3), {k:round(x,1) for k,x in j['kernel_ms_per_step'].items()})
for v in j.get('extra',{}).get('scenes',[]): print('  ',v['scene'], {k:(round(x['ms_per_frame'],3), round(x['mrays_per_s'],1)) for k,x in v.items() if isinstance(x,dict) and 'ms_per_frame' in x})
" | tee -a gpurun_out/${TAG}.txt
done
cp /tmp/tree.so $LIB
