"""The reference-side binding (oracle/gpu_bridge.cpp): the REFERENCE's own parser and scene graph, after its own
beginRender / beginFrame, filling include/hxr.h from the live `Scene` and rendering through the library. Proves that the POD
tables carry everything the reference's classes hold:
  * the bridge's frame equals the frame of this repo's own front-end (same tables => same rays, same Philox streams), and
  * deterministic scenes are within the parity tolerance of the reference's CPU render of the same process.
CPU tier: the bridge linked with the host emulation (oracle/_ref/hexray_ref_emu); GPU tier: with the product, all 13 scenes."""
import json
import os
import subprocess

import numpy as np
import pytest

import hexray_b200 as hx
import hxr_testlib as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALL_SCENES = ["simple", "meshes", "kdtree_test", "heightfield", "bumpmap", "Lecture8", "beer",
              "cornell_box", "smallpt", "hw12/sphtri", "zaphod", "hw10/bokeh", "boxed"]
STOCHASTIC = {"cornell_box", "smallpt", "hw12/sphtri", "zaphod", "hw10/bokeh", "boxed"}


def bridge_render(binary, scene, W, H, spp):
    exe = os.path.join(ROOT, "oracle", "_ref", binary)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/%s not built (make -C oracle, needs /root/reference)" % binary)
    out = "/tmp/hxr_bridge_%d.f32" % os.getpid()
    args = [exe, "gpurender", T.scene_path(scene), "--width", str(W), "--height", str(H), "--out", out]
    if spp:
        args += ["--spp", str(spp)]
    p = subprocess.run(args, cwd=os.path.dirname(hx.data_root()), capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-500:]
    return np.fromfile(out, dtype=np.float32).reshape(H, W, 3)


def check_bridge(api, binary, scene, W, H, spp):
    img = bridge_render(binary, scene, W, H, spp)
    sf = hx.SceneFile(T.scene_path(scene), api_=api)
    r = hx.Renderer(api_=api, queue_capacity=1 << 20).load(sf)
    # the bridge passes the scene's own settings and seed 17 (oracle/gpu_bridge.cpp); boxed: Whitted with a jittered area light
    ours, st = r.render(width=W, height=H, spp=spp if scene != "boxed" else 0, seed=17)
    r.close()
    sf.close()
    assert float(img.mean()) > 1e-3
    scale = max(1.0, float(np.abs(ours).max()))
    d = np.abs(img - ours)
    # the same tables give the same rays; what is left is the order of the float atomics. A handful of pixels may differ
    # where the two parsers' doubles differ in the last bit (strtod against sscanf) and an edge falls on the other side
    assert np.median(d) < 1e-6 and (d.max(axis=2) > 2e-3 * scale).mean() < 2e-3, "%s: bridge and front-end frames differ (max %.3g)" % (scene, d.max())
    if scene not in STOCHASTIC and T.have_oracle():
        ref, _ = T.oracle_render(scene, W, H)
        frac, mx = T.pixel_match_fraction(img, ref)
        assert frac >= T.PIXEL_FRACTION, "%s through the bridge: %.4f%% within 1/255 of the reference (max %.4f)" % (scene, frac * 100, mx)


@pytest.mark.parametrize("scene,spp", [("kdtree_test", 0), ("heightfield", 0), ("simple", 0), ("cornell_box", 4)])
def test_bridge_on_the_emulation(emu_api, scene, spp):
    check_bridge(emu_api, "hexray_ref_emu", scene, 96, 72, spp)


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ALL_SCENES)
def test_bridge_renders_every_scene(gpu_api, scene):
    check_bridge(gpu_api, "hexray_ref_gpu", scene, 320, 240, 16 if scene in STOCHASTIC else 0)
