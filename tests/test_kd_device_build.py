"""The device KD-tree build (csrc/device/kdbuild.h, csrc/kd_device_build.cpp): CPU tier on the emulation's loops,
GPU tier on the kernels up to the bench's 10 M triangles."""
import pytest

import hexray_b200 as hx
import hxr_testlib as T


@pytest.mark.parametrize("kind,size,rays", [("terrain", 48, 512), ("terrain", 130, 256), ("soup", 3000, 512)])
def test_device_build_on_the_emulation(emu_api, kind, size, rays):
    T.check_device_kd_build(emu_api, kind, size, rays)


def test_bundled_mesh_scene_on_the_device_built_tree(emu_api):
    # a whole frame (several meshes, shading) through device-built trees equals the frame through host-built ones
    import numpy as np
    sf = hx.SceneFile(T.scene_path("meshes"), api_=emu_api)
    img = {}
    for flags in (0, hx.CFG_DEVICE_KD_BUILD):
        r = hx.Renderer(api_=emu_api, flags=flags).load(sf)
        img[flags] = r.render(width=96, height=72, want_aa=0)[0]
        r.close()
    sf.close()
    assert np.abs(img[0] - img[hx.CFG_DEVICE_KD_BUILD]).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("kind,size,rays", [("terrain", 320, 2048), ("soup", 200000, 2048), ("terrain", 2237, 192)])
def test_device_build_on_the_gpu(gpu_api, kind, size, rays):
    info = T.check_device_kd_build(gpu_api, kind, size, rays)
    print("device build %s %d: %.1f ms (device passes %.1f ms) against host %.1f ms%s" % (
        kind, size, info["device"]["build_ms"], info["device"]["device_ms"], info["host"]["build_ms"], " (host tree from the cache)" if info["host"]["from_cache"] else ""))


@pytest.mark.gpu
def test_whitted_frame_on_device_built_trees(gpu_api):
    import numpy as np
    sf = hx.SceneFile(T.scene_path("kdtree_test"), api_=gpu_api)
    img = {}
    for flags in (0, hx.CFG_DEVICE_KD_BUILD):
        r = hx.Renderer(api_=gpu_api, flags=flags).load(sf)
        img[flags] = r.render()[0]
        r.close()
    sf.close()
    assert np.abs(img[0] - img[hx.CFG_DEVICE_KD_BUILD]).max() < 1e-4
