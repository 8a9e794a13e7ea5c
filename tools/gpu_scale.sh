#!/bin/bash
# multi-GPU check: our arm under torchrun on N GPUs of one box (as the driver launches it)
N=${1:-2}; TAG=${2:-scale}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_smi.txt
python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err; echo "n$N rc=$?"
tail -3 gpurun_out/${TAG}_n$N.err
python -c "
import json
for n in (1,$N):
    j=json.loads(open('gpurun_out/${TAG}_n%d.json'%n).read().strip().splitlines()[-1])
    print(n, round(j['value'],1), round(j['ms_per_step'],1), round(j['wall_ms_per_step'],1), round(j['e2e']['value'],1), j['config']['spp_total'])
"
