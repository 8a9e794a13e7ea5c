#!/bin/bash
# like gpu_scene_prof.sh, scenes given: gpu_scene_prof2.sh TAG "scene W H;scene W H" "ENV=.." ...  ("-" = defaults)
TAG=$1; SC=$2; shift 2
mkdir -p gpurun_out; : > gpurun_out/${TAG}.txt
for E in "$@"; do
  if [ "$E" = "-" ]; then E=""; fi
  echo "## env: $E" | tee -a gpurun_out/${TAG}.txt
  IFS=';' read -ra LIST <<< "$SC"
  for s in "${LIST[@]}"; do
    env $E python tools/scene_prof.py $s 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['scene'], 'prof' if j['profiling'] else 'plain', 'Mrays/s=%.0f'%j['mrays_s'], 'ms=%.2f'%j['render_ms'], 'walk=%.2f'%j['walk_ms'], 'finish=%.2f'%j['finish_ms'], 'shade=%.2f'%j['shade_ms'], 'setup=%.2f'%j['setup_ms'])" | tee -a gpurun_out/${TAG}.txt
  done
done
