"""CPU tier: host-side pieces of the product library (no GPU compute): image codecs, the `.hexray`
language (accepted forms and error behaviour), the ABI surface, the loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import hexray_b200 as hx
from hexray_b200 import capi
import hxr_testlib as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(gpu_api):
    """Every function include/hxr.h declares is exported by libhexray_b200.so (and bound in capi.SYMBOLS)."""
    header = open(os.path.join(ROOT, "include", "hxr.h")).read()
    declared = set(re.findall(r"\b(hxr_[a-z_0-9]+)\s*\(", header)) - {"hxr_scene"}
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for name in declared:
        assert hasattr(gpu_api.lib, name)


def test_struct_sizes_match_header(gpu_api):
    # sizes the C compiler gives the ABI structs (pinned so the ctypes mirror cannot drift)
    assert C.sizeof(capi.Transform) == 240 and C.sizeof(capi.Geometry) == 64 and C.sizeof(capi.Triangle) == 184
    assert C.sizeof(capi.Node) == 256 and C.sizeof(capi.Shader) == 56 and C.sizeof(capi.Layer) == 20
    assert C.sizeof(capi.Texture) == 56 and C.sizeof(capi.Light) == 304 and C.sizeof(capi.Camera) == 208
    assert C.sizeof(capi.RenderParams) == 48 and C.sizeof(capi.Stats) == 192 and C.sizeof(capi.Config) == 32


def test_no_cpu_fallback(gpu_api):
    """Without a CUDA device the product refuses to create a context (this test only asserts on GPU-less hosts)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(hx.HxrError) as e:
        hx.Renderer(api_=gpu_api)
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)


def test_exr_codec_matches_opencv(gpu_api):
    """Our PIZ/EXR reader against OpenCV's independent OpenEXR build, bit for bit, on every bundled cubemap face."""
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    n = 0
    for env in ("forest", "ocean"):
        for face in ("negx", "negy", "negz", "posx", "posy", "posz"):
            p = os.path.join(hx.data_root(), "env", env, face + ".exr")
            ours = hx.load_image(p, api_=gpu_api)
            ref = cv2.imread(p, cv2.IMREAD_UNCHANGED)
            if ref is None:
                pytest.skip("OpenCV was built without OpenEXR")
            assert np.array_equal(ours, ref[..., [2, 1, 0]].astype(np.float32)), p
            n += 1
    assert n == 12


def test_exr_and_bmp_round_trip(gpu_api, tmp_path):
    rng = np.random.default_rng(0)
    img = (rng.random((37, 53, 3)) * 4).astype(np.float32)
    p = str(tmp_path / "a.exr")
    hx.save_image(p, img, api_=gpu_api)
    back = hx.load_image(p, api_=gpu_api)
    assert np.array_equal(back, img.astype(np.float16).astype(np.float32))  # HALF storage, as Bitmap::saveEXR
    q = str(tmp_path / "a.bmp")
    hx.save_image(q, img, api_=gpu_api)
    b8 = hx.load_image(q, api_=gpu_api)
    # BMP goes through the reference's sRGB LUT: index int(x*4096), toe multiplier 12.02 (src/color.h:36-47)
    def lut(x):
        x = np.clip(x, 0, 1)
        xi = np.floor(x * 4096.0).astype(np.int64) / 4096.0
        y = np.where(xi <= 0.0031308, xi * 12.02, 1.055 * np.power(xi, 1 / 2.4) - 0.055)
        out = np.floor(y.astype(np.float32) * 255.0 + 0.5) / 255.0
        return np.where(x >= 1, 1.0, np.where(x <= 0, 0.0, out))
    assert np.abs(b8 - lut(img)).max() <= 1 / 255 + 1e-6


def _scene(tmp_path, body):
    p = tmp_path / "s.hexray"
    p.write_text(body)
    return str(p)


CAM = "Camera c {\n pos (0, 1, -5)\n}\n"


def test_language_forms(gpu_api, tmp_path):
    # comments, quoted values, '(a b c)' vectors without commas, booleans, transform order, layered syntax
    body = """// line comment
# hash comment
/* block
   comment
*/
GlobalSettings {
  frameWidth 123   // trailing comment
  frameHeight 45
  wantAA off
  gi false
}
""" + CAM + """
Plane p { y 2 }
""".replace("Plane p { y 2 }", "Plane p {\n y 2\n}") + """
CheckerTexture ch {
  color1 (1 0.5 0.25)
  scaling 3
}
Fresnel fr {
  ior 1.5
}
Lambert l {
  texture ch
}
Reflection r {
  multiplier 0.5
}
Layered lay {
  layer l (1, 1, 1)
  layer r (0.2, 0.3, 0.4) fr
  layer r (0.1, 0.1, 0.1), NULL
}
Node n {
  geometry p
  shader lay
  scale (2, 2, 2)
  rotate (90, 0, 0)
  translate (1, 2, 3)
}
Node helper {
  geometry p
}
PointLight pl {
  pos (1, 2, 3)
  power 10
}
"""
    sf = hx.SceneFile(_scene(tmp_path, body), api_=gpu_api)
    s = sf.pod.contents
    assert (s.settings.frame_width, s.settings.frame_height, s.settings.want_aa, s.settings.gi) == (123, 45, 0, 0)
    assert s.n_nodes == 1  # the shader-less node is not a scene object
    assert s.n_layers == 3 and s.layers[1].tex == 1 and s.layers[2].tex == -1
    assert abs(s.layers[1].blend[1] - 0.3) < 1e-7
    assert tuple(s.textures[0].color1) == (1.0, 0.5, 0.25)
    assert tuple(s.shaders[1].color) == (0.5, 0.5, 0.5)
    n = s.nodes[0]
    assert tuple(n.T.offset) == (1.0, 2.0, 3.0)
    m = np.array(n.T.m).reshape(3, 3)
    inv = np.array(n.T.inv).reshape(3, 3)
    assert np.allclose(m @ inv, np.eye(3), atol=1e-12) and abs(abs(np.linalg.det(m)) - 8) < 1e-9
    cam = sf.camera()
    assert tuple(cam.pos) == (0.0, 1.0, -5.0) and abs(cam.aperture_size - 1.25) < 1e-12
    sf.close()


@pytest.mark.parametrize("body,needle", [
    ("Bogus b {\n}\n" + CAM, "Unknown object class"),
    ("Camera c {\n fov 60\n}\n", "Required property `pos' not defined"),
    (CAM + "Sphere s {\n R -1\n}\n", "outside the allowed bounds"),
    (CAM + "Node n {\n geometry nothere\n}\n", "Geometry not defined"),
    (CAM + "Mesh m {\n file \"missing.obj\"\n}\n", "Required file not found"),
    (CAM + "Plane p {\n y 1\n", "Unfinished object definition"),
    (CAM + "Plane p {\n y\n}\n", "Unexpected token in object definition"),
])
def test_language_errors(gpu_api, tmp_path, body, needle):
    with pytest.raises(hx.HxrError) as e:
        hx.SceneFile(_scene(tmp_path, body), api_=gpu_api)
    assert e.value.status == -4 and needle in str(e.value)


def test_all_bundled_scenes_parse(gpu_api):
    for scene in T.WHITTED_SCENES + T.MC_SCENES + ["boxed"]:
        sf = hx.SceneFile(T.scene_path(scene), api_=gpu_api)
        assert sf.pod.contents.n_nodes > 0
        sf.close()


def test_obj_round_trip(gpu_api, tmp_path):
    # synthetic terrain -> OBJ -> loader: identical triangles (the path that hands procedural meshes to the reference)
    a = _scene(tmp_path, CAM + "Mesh m {\n file \"synthetic:terrain:9:0x5EED\"\n}\nLambert l {\n}\nNode n {\n geometry m\n shader l\n}\n")
    sf = hx.SceneFile(a, api_=gpu_api)
    sf.write_obj(0, str(tmp_path / "t.obj"))
    (tmp_path / "b.hexray").write_text(CAM + "Mesh m {\n file \"t.obj\"\n}\nLambert l {\n}\nNode n {\n geometry m\n shader l\n}\n")
    sg = hx.SceneFile(str(tmp_path / "b.hexray"), api_=gpu_api)
    ma, mb = sf.pod.contents.meshes[0], sg.pod.contents.meshes[0]
    assert ma.n_triangles == mb.n_triangles == 128 and ma.n_vertices == mb.n_vertices == 82
    va = np.ctypeslib.as_array(ma.vertices, shape=(ma.n_vertices * 3,))
    vb = np.ctypeslib.as_array(mb.vertices, shape=(mb.n_vertices * 3,))
    assert np.array_equal(va, vb)
    sf.close()
    sg.close()


def T_payload(data, shape, ext):
    """Every byte of the file that the reference DEFINES. BMP: Bitmap::saveBMP (src/bitmap.cpp:231-238) fwrite()s rowsz
    bytes of a stack buffer of which it has filled only width * 3, so the 0-3 padding bytes of each row are uninitialised
    stack in the reference's file (zero in ours, as its own comment and the format ask): they are left out. EXR: all of it."""
    if ext != ".bmp":
        return data
    H, W = shape[:2]
    rowsz = (W * 3 + 3) // 4 * 4
    body = data[54:].reshape(H, rowsz)[:, :W * 3]
    return np.concatenate([data[:54], body.reshape(-1)])


def test_saved_files_equal_the_reference_bytes(emu_api, tmp_path):
    """a27, byte for byte: the reference's own screenshot of a frame (takeScreenshot -> Bitmap::saveBMP / saveEXR,
    src/sdl.cpp:103-116, src/bitmap.cpp:202-288; fixture written by the compiled reference, tests/golden/make_golden.py
    output) against hxr_save_image of the SAME float frame - pure host code, so this runs on the CPU tier."""
    import hexray_b200 as hx
    g = np.load(os.path.join(ROOT, "tests", "golden", "output_simple.npz"))
    vfb = g["vfb"]
    for ext, key in ((".bmp", "bmp"), (".exr", "exr")):
        p = str(tmp_path / ("ours" + ext))
        hx.save_image(p, vfb, api_=emu_api)
        ours = np.frombuffer(open(p, "rb").read(), dtype=np.uint8)
        ref = g[key]
        assert len(ours) == len(ref)
        assert np.array_equal(T_payload(ours, vfb.shape, ext), T_payload(ref, vfb.shape, ext)), "%s differs from the reference's file" % ext
    # the EXR pixels against an independent float -> half (numpy, IEEE round to nearest even): exact
    H, W = vfb.shape[:2]
    exr = g["exr"]
    body = exr[len(exr) - H * (8 + W * 8):].reshape(H, 8 + W * 8)[:, 8:].copy().view(np.uint16).reshape(H, 4, W)
    want = vfb.astype(np.float16).view(np.uint16)
    assert np.array_equal(body[:, 3, :], want[:, :, 0]) and np.array_equal(body[:, 2, :], want[:, :, 1]) and np.array_equal(body[:, 1, :], want[:, :, 2])
    assert (body[:, 0, :] == 0x3C00).all()  # alpha 1
    back = hx.load_image(str(tmp_path / "ours.exr"), api_=emu_api)
    assert np.array_equal(back, vfb.astype(np.float16).astype(np.float32))
