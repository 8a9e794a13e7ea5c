#!/bin/bash
# One call: the parity tests that reach the overflow-finish kernels, then A/B of library builds x environment settings on the
# terrain bench.  gpu_ab5.sh TAG SPP "lib|ENV=.. ENV2=.." ...   (lib: "tree" or a path; the part after | may be empty)
TAG=$1; SPP=${2:-32}; shift 2
mkdir -p gpurun_out; : > gpurun_out/${TAG}.jsonl
LIB=hexray_b200/libhexray_b200.so
cp $LIB /tmp/tree.so
if [ -n "$HXR_AB_PYTEST" ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "$HXR_AB_PYTEST" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -3 gpurun_out/${TAG}_pytest.log
fi
for spec in "$@"; do
  v=${spec%%|*}; E=${spec#*|}
  if [ "$v" = "tree" ]; then cp /tmp/tree.so $LIB; else cp $v $LIB; fi
  env $E timeout 600 python bench.py --steps 2 --warmup 2 --spp $SPP --no-cpu-baseline --no-extra ${HXR_AB_ARGS} 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'spec':'$spec','value':round(j['value'],1),'ms':{k:round(x,1) for k,x in j['kernel_ms_per_step'].items()}}))" | tee -a gpurun_out/${TAG}.jsonl
done
cp /tmp/tree.so $LIB
