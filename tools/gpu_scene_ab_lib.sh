#!/bin/bash
# A/B of library builds on the Whitted scenes (per-kernel-class times): gpu_scene_ab_lib.sh TAG "scene W H;scene W H" lib1 lib2 ...  ("tree" = the tree's build)
TAG=$1; SC=$2; shift 2
mkdir -p gpurun_out; : > gpurun_out/${TAG}.txt
LIB=hexray_b200/libhexray_b200.so
cp $LIB /tmp/tree.so
for v in "$@"; do
  if [ "$v" = "tree" ]; then cp /tmp/tree.so $LIB; else cp $v $LIB; fi
  echo "## lib: $v" | tee -a gpurun_out/${TAG}.txt
  IFS=';' read -ra LIST <<< "$SC"
  for s in "${LIST[@]}"; do
    python tools/scene_prof.py $s 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['scene'], 'prof' if j['profiling'] else 'plain', 'Mrays/s=%.0f'%j['mrays_s'], 'ms=%.2f'%j['render_ms'], 'walk=%.2f'%j['walk_ms'], 'finish=%.2f'%j['finish_ms'], 'shade=%.2f'%j['shade_ms'], 'setup=%.2f'%j['setup_ms'], 'gen=%.2f'%j['gen_ms'])" | tee -a gpurun_out/${TAG}.txt
  done
done
cp /tmp/tree.so $LIB
