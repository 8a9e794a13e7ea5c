#!/bin/bash
# two-lane drain on/off: the bench with its extra scenes for both settings
mkdir -p gpurun_out
for E in "HXR_NO_OVERLAP=1" "HXR_X=0"; do
  env $E python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$E', round(j['value'],1), round(j['roofline']['frac'],3), {k:round(x,1) for k,x in j['kernel_ms_per_step'].items()})
for v in j.get('extra',{}).get('scenes',[]): print('  ',v['scene'], {k:(round(x['ms_per_frame'],3), round(x['mrays_per_s'],1)) for k,x in v.items() if isinstance(x,dict) and 'ms_per_frame' in x})
" | tee -a gpurun_out/ov_bench.txt
done
