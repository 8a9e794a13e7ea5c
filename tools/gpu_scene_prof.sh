#!/bin/bash
# Whitted scenes at 1080p-class size: per-kernel-class device time, per environment setting. gpu_scene_prof.sh TAG "ENV=.." ...  ("-" = defaults)
TAG=$1; shift 1
mkdir -p gpurun_out; : > gpurun_out/${TAG}.txt
for E in "$@"; do
  if [ "$E" = "-" ]; then E=""; fi
  echo "## env: $E" | tee -a gpurun_out/${TAG}.txt
  for s in "meshes 1920 1440" "beer 1920 1440" "kdtree_test 1920 1440" "heightfield 1920 1440" "simple 1920 1080"; do
    env $E python tools/scene_prof.py $s 2>/dev/null | tee -a gpurun_out/${TAG}.txt
  done
done
