#!/bin/bash
# scaling evidence on N GPUs of one box: gpu_scale.sh N [tests]
#   torchrun bench (one process per GPU, NCCL reduce) and the library-owned multi-GPU context (one process, peer-memory reduce)
N=$1
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then
  ( timeout 1200 python -m pytest tests/test_multi_device.py -m gpu -x -q ) > gpurun_out/scale_n${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/scale_n${N}_pytest.log
  ( tests/cpp/build.sh && cd assets && ../tests/cpp/bin/test_multi_gpu data/cornell_box.hexray $(seq -s, 0 $((N-1))) 64 256 ) > gpurun_out/scale_n${N}_cpp.log 2>&1; echo "cpp rc=$?"; tail -5 gpurun_out/scale_n${N}_cpp.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/scale_r2_torchrun_n${N}.json 2> gpurun_out/scale_r2_torchrun_n${N}.err; echo "torchrun rc=$?"
python bench.py --gpus $N --no-cpu-baseline > gpurun_out/scale_r2_library_n${N}.json 2> gpurun_out/scale_r2_library_n${N}.err; echo "library rc=$?"
HXR_REDUCE=nccl python bench.py --gpus $N --no-cpu-baseline --steps 2 > gpurun_out/scale_r2_library_nccl_n${N}.json 2> gpurun_out/scale_r2_library_nccl_n${N}.err; echo "library nccl rc=$?"
python - <<PY
import json
for f in ("torchrun", "library", "library_nccl"):
    try:
        j = json.loads(open("gpurun_out/scale_r2_%s_n$N.json" % f).read().strip().splitlines()[-1])
        print(f, j["n_gpus"], round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "ms/step", round(j["ms_per_step"], 1), "reduce_ms", j["kernel_ms_per_step"].get("reduce_ms"), "setup_s", round(j["config"]["setup_s"], 1), j["config"]["sharding"])
    except Exception as e:
        print(f, "failed", e)
PY
# tile-sharded Whitted frame (16-row bands + one-row halo) at 1080p on 1 and N GPUs of this box
HEXRAY_DATA=assets/data python - <<PY | tee gpurun_out/scale_r2_whitted_n${N}.txt
import numpy as np, hexray_b200 as hx, os
sf = hx.SceneFile(os.path.join(hx.data_root(), "kdtree_test.hexray"))
out = {}
for devs in ([0], list(range($N))):
    r = hx.Renderer(devices=devs).load(sf)
    ms = []
    for i in range(6):
        img, st = r.render(width=1920, height=1080)
        if i >= 2: ms.append(st["render_ms"])
    out[len(devs)] = (img, float(np.median(ms)), st["rays_closest"] + st["rays_shadow"], st.get("reduce_ms", 0.0))
    r.close()
a, b = out[1], out[$N]
print("kdtree_test 1920x1080 Whitted+AA: 1 GPU %.3f ms (%d rays) | $N GPUs %.3f ms (%d rays incl. halo rows, reduce %.3f ms) | speed-up %.2f | max |diff| %.2e" % (a[1], a[2], b[1], b[2], b[3], a[1] / b[1], float(np.abs(a[0] - b[0]).max())))
PY
