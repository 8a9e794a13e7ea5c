"""The binary caches in front of the first ray (host/cache.h): a parsed OBJ and a built KD-tree come back from disk, give
bit-identical results, and a tree that another process is building is waited for, not built twice. Pure host code + the
host emulation: CPU tier."""
import os
import subprocess
import time

import numpy as np

import hexray_b200 as hx
import hxr_testlib as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_obj_and_tree_caches_round_trip(emu_api, tmp_path, monkeypatch):
    monkeypatch.setenv("HXR_CACHE_DIR", str(tmp_path / "cache"))
    # a 320-side terrain (203 522 triangles) as an OBJ file, written by the host-only generator
    gen = os.path.join(ROOT, "tools", "bin", "hxr_objgen")
    if not os.path.exists(gen):
        subprocess.run([os.path.join(ROOT, "tools", "build_tools.sh")], check=True)
    obj = str(tmp_path / "terrain.obj")
    subprocess.run([gen, "terrain", "320", "0x5EED", obj], check=True)
    import argparse
    import bench
    scene = str(tmp_path / "t.hexray")
    with open(scene, "w") as f:
        f.write(bench.scene_text(argparse.Namespace(grid_side=320), "terrain.obj", 64, 36, 4))  # (asset paths are relative to the scene file)
    rays = np.zeros((256, 8))
    rays[:, 0:3] = (0.0, 150.0, -600.0)
    rng = np.random.default_rng(1)
    d = np.stack([rng.uniform(-0.7, 0.7, 256), rng.uniform(-0.6, 0.05, 256), np.ones(256)], 1)
    rays[:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    out, times = [], []
    for run in range(2):
        t0 = time.time()
        sf = hx.SceneFile(scene, api_=emu_api)
        t1 = time.time()
        r = hx.Renderer(api_=emu_api, queue_capacity=1 << 16).load(sf)
        info = r.accel_info(0)
        hits = r.trace_closest(rays).copy()
        r.close()
        sf.close()
        out.append((hits, info))
        times.append(t1 - t0)
    files = sorted(os.listdir(tmp_path / "cache"))
    assert any(f.startswith("obj_") for f in files) and any(f.startswith("kd_") for f in files), files
    (h0, i0), (h1, i1) = out
    assert i0["from_cache"] == 0 and i1["from_cache"] == 1
    assert i0["nodes"] == i1["nodes"] and i0["tri_refs"] == i1["tri_refs"]
    for k in ("status", "node", "dist", "ip", "norm", "u", "v"):
        assert np.array_equal(h0[k], h1[k]), k
    assert (h0["node"] == 0).sum() > 32
    assert times[1] < times[0]  # the binary blob reads faster than the text parses


def test_a_tree_being_built_elsewhere_is_waited_for(emu_api, tmp_path, monkeypatch):
    # two processes load the same big mesh at once (what the ranks of a torchrun job do): exactly one of them builds
    monkeypatch.setenv("HXR_CACHE_DIR", str(tmp_path / "cache"))
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import hexray_b200 as hx, hxr_testlib as T\n"
            "from hexray_b200 import capi\n"
            "api = capi.Api(%r)\n"
            "sf = T.terrain_scene_file(api, 500)\n"
            "r = hx.Renderer(api_=api, queue_capacity=1 << 16).load(sf)\n"
            "print(json.dumps(r.accel_info(0)))\n") % (ROOT, os.path.join(ROOT, "tests"), emu_api.path)
    env = dict(os.environ, HXR_CACHE_DIR=str(tmp_path / "cache"))
    ps = [subprocess.Popen(["python", "-c", code], stdout=subprocess.PIPE, text=True, env=env) for _ in range(3)]
    infos = []
    for p in ps:
        o, _ = p.communicate(timeout=600)
        assert p.returncode == 0
        import json
        infos.append(json.loads(o.strip().splitlines()[-1]))
    assert sum(1 for i in infos if i["from_cache"] == 0) == 1, infos
    assert len({(i["nodes"], i["tri_refs"]) for i in infos}) == 1


def _load_terrain(emu_api, side):
    sf = T.terrain_scene_file(emu_api, side)
    r = hx.Renderer(api_=emu_api, queue_capacity=1 << 16).load(sf)
    info = r.accel_info(0)
    r.close()
    sf.close()
    return info


def test_a_damaged_tree_file_is_rebuilt(emu_api, tmp_path, monkeypatch):
    # the files carry a hash of their payload: a flipped byte in the middle of the tree must not reach the walk
    monkeypatch.setenv("HXR_CACHE_DIR", str(tmp_path / "cache"))
    first = _load_terrain(emu_api, 330)
    assert first["from_cache"] == 0
    kd = [f for f in os.listdir(tmp_path / "cache") if f.startswith("kd_") and f.endswith(".bin")]
    assert len(kd) == 1
    assert _load_terrain(emu_api, 330)["from_cache"] == 1
    path = tmp_path / "cache" / kd[0]
    blob = bytearray(path.read_bytes())
    blob[len(blob) // 2] ^= 0x40
    path.write_bytes(bytes(blob))
    again = _load_terrain(emu_api, 330)
    assert again["from_cache"] == 0 and again["nodes"] == first["nodes"] and again["tri_refs"] == first["tri_refs"]
    assert _load_terrain(emu_api, 330)["from_cache"] == 1  # (the rebuilt tree replaced the damaged file)


def test_a_stale_lock_is_taken_over(emu_api, tmp_path, monkeypatch):
    # a run killed in the middle of its build leaves its lock file behind: the next run must not wait for it
    monkeypatch.setenv("HXR_CACHE_DIR", str(tmp_path / "cache"))
    first = _load_terrain(emu_api, 335)
    kd = [f for f in os.listdir(tmp_path / "cache") if f.startswith("kd_") and f.endswith(".bin")]
    assert len(kd) == 1
    path = tmp_path / "cache" / kd[0]
    p = subprocess.Popen(["true"])
    p.wait()  # a pid that no longer exists
    os.remove(path)
    (tmp_path / "cache" / (kd[0] + ".lock")).write_text("%d\n" % p.pid)
    t0 = time.time()
    again = _load_terrain(emu_api, 335)
    assert time.time() - t0 < 60
    assert again["from_cache"] == 0 and again["nodes"] == first["nodes"]
    assert not os.path.exists(tmp_path / "cache" / (kd[0] + ".lock")) and os.path.exists(path)
