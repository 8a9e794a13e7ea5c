"""CPU tier, world_size 2 over gloo: the multi-rank frame logic of hexray_b200.distributed (sample sharding for
Monte-Carlo frames, row-band sharding with halo for Whitted frames, one reduce, resolve on rank 0) — run on the host
emulation library, because this container has no GPU. The product path is the same code with NCCL and CUDA tensors
(bench.py, tests/test_gpu_parity.py::test_shard_invariance_*)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emu_so, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hexray_b200 as hx
    from hexray_b200 import capi, distributed
    import hxr_testlib as T
    api = capi.Api(emu_so)
    res = {}
    for scene, W, H, spp, mode in (("cornell_box", 64, 64, 8, hx.MODE_MONTECARLO), ("kdtree_test", 96, 72, 0, hx.MODE_WHITTED)):
        sf = hx.SceneFile(T.scene_path(scene), api_=api)
        r = hx.Renderer(api_=api, queue_capacity=1 << 18).load(sf)
        acc = torch.zeros(H * W * 3, dtype=torch.float32)
        st = distributed.render_frame(r, acc, W, H, spp_total=spp, seed=7, rank=rank, world=world, mode=mode, dist=dist)
        if rank == 0:
            full, _ = r.render(width=W, height=H, spp=spp, seed=7, mode=mode)
            res[scene] = (acc.numpy().reshape(H, W, 3).copy(), full)
        res[scene + "_rays"] = st["rays_closest"]
        r.close()
        sf.close()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([res["cornell_box_rays"], res["kdtree_test_rays"]]))
    if rank == 0:
        for scene in ("cornell_box", "kdtree_test"):
            a, b = res[scene]
            np.savez(os.path.join(out_dir, scene + ".npz"), sharded=a, full=b)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(emu_api, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), emu_api.path, str(tmp_path)), nprocs=world, join=True)
    for scene in ("cornell_box", "kdtree_test"):
        z = np.load(tmp_path / (scene + ".npz"))
        assert np.abs(z["sharded"] - z["full"]).max() < 2e-4 * max(1.0, float(np.abs(z["full"]).max())), scene
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert r0[0] > 0 and r1[0] > 0 and abs(int(r0[0]) - int(r1[0])) < 0.2 * r0[0]  # both ranks did about half the samples
