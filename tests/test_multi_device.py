"""One context over several devices (hxr_config.devices): the library shards the frame, sums the partial frames and resolves.
CPU tier: the host emulation's two fake devices through ctypes and through the C++ host program; GPU tier: the product."""
import os
import subprocess

import numpy as np
import pytest

import hexray_b200 as hx
import hxr_testlib as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_cpp(lib, out):
    subprocess.run([os.path.join(ROOT, "tests", "cpp", "build.sh"), lib, out], check=True)
    return os.path.join(ROOT, "tests", "cpp", out, "test_multi_gpu")


def check_multi(api, devices, spp, size):
    out = {}
    for scene, kw in (("cornell_box", dict(spp=spp, seed=5)), ("kdtree_test", {})):
        sf = hx.SceneFile(T.scene_path(scene), api_=api)
        imgs = []
        for devs in (None, devices):
            r = hx.Renderer(api_=api, queue_capacity=1 << 20, devices=devs).load(sf)
            img, st = r.render(width=size, height=size, **kw)
            imgs.append((img.copy(), st))
            if devs:
                assert st["n_devices"] == len(devs) and r.reduce_backend() in ("peer", "nccl", "host")
            r.close()
        sf.close()
        (a, sa), (b, sb) = imgs
        if scene == "cornell_box":
            assert sa["rays_closest"] == sb["rays_closest"]  # sample passes: exactly the same rays, dealt over the devices
        else:
            assert sb["rays_closest"] >= sa["rays_closest"]  # row bands: plus one halo row per band for the AA detector
        assert float(a.mean()) > 0.01
        assert np.abs(a - b).max() < 2e-4 * max(1.0, float(a.max())), scene
        out[scene] = sb
    return out


def test_two_emulated_devices(emu_api):
    check_multi(emu_api, [0, 1], 8, 64)


def test_cpp_host_on_the_emulation(emu_api):
    exe = build_cpp(os.path.join(ROOT, "tests", "emu", "libhxr_emu.so"), "bin_emu")
    p = subprocess.run([exe, T.scene_path("cornell_box"), "0,1", "8", "64"], capture_output=True, text=True, cwd=os.path.dirname(hx.data_root()))
    assert p.returncode == 0 and "OK" in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
def test_multi_context_on_the_gpu(gpu_api):
    # two contexts' worth of GPU state even on a one-GPU box (the same ordinal twice): the sharded render, the peer reduce
    # kernel and the resolve are all exercised; on a multi-GPU box the second device is a different GPU
    n = gpu_api.lib.hxr_device_count()
    assert n >= 1
    check_multi(gpu_api, [0, 1 % n], 64, 256)


@pytest.mark.gpu
def test_cpp_host_renders_on_several_gpus(gpu_api):
    exe = build_cpp(gpu_api.path, "bin")
    n = gpu_api.lib.hxr_device_count()
    devs = ",".join(str(i % n) for i in range(max(2, min(n, 8))))
    p = subprocess.run([exe, T.scene_path("cornell_box"), devs, "64", "256"], capture_output=True, text=True, cwd=os.path.dirname(hx.data_root()))
    assert p.returncode == 0 and "OK" in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
def test_nccl_reduce_when_several_gpus(gpu_api):
    if gpu_api.lib.hxr_device_count() < 2:
        pytest.skip("one GPU: NCCL refuses the same device twice")
    os.environ["HXR_REDUCE"] = "nccl"
    try:
        st = check_multi(gpu_api, [0, 1], 64, 256)
    finally:
        del os.environ["HXR_REDUCE"]
    assert st["cornell_box"]["n_devices"] == 2


def check_progressive(api, devices):
    """Pass-by-pass refinement (hxr_progressive_*): the estimate after the last pass is the frame hxr_render returns, the
    estimates in between are proper (unbiased, noisier) frames, one reduce per pass."""
    sf = hx.SceneFile(T.scene_path("cornell_box"), api_=api)
    r = hx.Renderer(api_=api, queue_capacity=1 << 20, devices=devices).load(sf)
    full, st = r.render(width=64, height=64, spp=16, seed=9)
    errs, spps = [], []
    for est, pst in r.progressive(4, width=64, height=64, spp=16, seed=9):
        errs.append(float(np.abs(est - full).mean()))
        spps.append(pst["spp_done"])
        assert float(est.mean()) > 0.01
    assert spps == [4, 8, 12, 16]
    assert errs[-1] < 2e-5 * max(1.0, float(full.max())) and errs[0] > errs[-1]
    r.close()
    sf.close()


def check_progressive_resume(api, devices):
    """A progressive frame interrupted after two of four passes goes on from its checkpoint (hxr_progressive_state ->
    hxr_progressive_resume) in a NEW context and ends in the same estimates; a checkpoint of another plan is refused."""
    sf = hx.SceneFile(T.scene_path("cornell_box"), api_=api)
    kw = dict(width=64, height=64, spp=16, seed=9)
    r = hx.Renderer(api_=api, queue_capacity=1 << 20, devices=devices).load(sf)
    straight = [est.copy() for est, _ in r.progressive(4, **kw)]
    gen = r.progressive(4, **kw)
    next(gen)
    next(gen)
    ck = r.progressive_state(64, 64)
    assert ck[1] == 2 and ck[2] == 8 and float(ck[0].mean()) > 0.01
    r.close()
    r2 = hx.Renderer(api_=api, queue_capacity=1 << 20, devices=devices).load(sf)
    rest = [(est.copy(), st["spp_done"]) for est, st in r2.progressive(4, checkpoint=ck, **kw)]
    assert [s for _, s in rest] == [12, 16]
    for (est, _), ref in zip(rest, straight[2:]):
        assert np.allclose(est, ref, rtol=2e-5, atol=1e-6), float(np.abs(est - ref).max())
    for bad in (dict(n=8, ck=ck), dict(n=4, ck=(ck[0], 3, ck[2])), dict(n=4, ck=(ck[0], 2, 9))):
        with pytest.raises(hx.HxrError):
            list(r2.progressive(bad["n"], checkpoint=bad["ck"], **kw))
    r2.close()
    sf.close()


def test_progressive_resume_on_the_emulation(emu_api):
    check_progressive_resume(emu_api, None)
    check_progressive_resume(emu_api, [0, 1])


@pytest.mark.gpu
def test_progressive_resume_on_the_gpu(gpu_api):
    check_progressive_resume(gpu_api, None)
    check_progressive_resume(gpu_api, [0, 1 % gpu_api.lib.hxr_device_count()])


def test_progressive_on_the_emulation(emu_api):
    check_progressive(emu_api, None)
    check_progressive(emu_api, [0, 1])


@pytest.mark.gpu
def test_progressive_on_the_gpu(gpu_api):
    check_progressive(gpu_api, None)
    check_progressive(gpu_api, [0, 1 % gpu_api.lib.hxr_device_count()])
