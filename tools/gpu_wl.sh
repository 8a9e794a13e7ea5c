#!/bin/bash
# tests + one secondary workload bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload cornell_box --spp 256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/cornell_v11.json 2> gpurun_out/cornell_v11.err; echo rc=$?
python -c "
import json
j=json.loads(open('gpurun_out/cornell_v11.json').read().strip().splitlines()[-1]); print('cornell', round(j['value'],1), round(j['ms_per_step'],1), j['kernel_ms_per_step'])"
