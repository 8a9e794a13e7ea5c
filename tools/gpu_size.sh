#!/bin/bash
# walk cost against mesh size (is the 10 M case memory-bound?): same scene, grid sides given as arguments
TAG=$1; shift
mkdir -p gpurun_out; : > gpurun_out/${TAG}.jsonl
for g in "$@"; do
  timeout 600 python bench.py --steps 2 --warmup 2 --spp 16 --no-cpu-baseline --grid-side $g 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=j['config']['rays_per_step']; k=j['kernel_ms_per_step']; p=j['roofline']['per_ray']
print(json.dumps({'grid':$g,'tris':j['config']['kd']['n_triangles'],'value':round(j['value'],1),'rays':r,'walk_ns_per_ray':round(k['walk_ms']*1e6/r,4),'nonwalk_ns_per_ray':round((k['render_ms']-k['walk_ms'])*1e6/r,4),'per_ray':{a:round(b,2) for a,b in p.items()},'frac':round(j['roofline']['frac'],3)}))" | tee -a gpurun_out/${TAG}.jsonl
done
