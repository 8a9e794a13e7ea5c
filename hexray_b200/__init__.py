"""hexray_b200 — B200-native render hot path of heX-Ray behind the reference's scene language.

Python here is plumbing only (the reference's host language is C++; the product is
libhexray_b200.so = C++ host front-end + hand-written sm_100a kernels, see include/hxr.h).
This module mirrors the reference's top-level flow for tests and benchmarks:

    scene.parseScene(file)            -> SceneFile(path)                 (src/scene.cpp:735)
    scene.beginRender(); beginFrame() -> done inside SceneFile           (src/main.cpp:506-511)
    render(false)                     -> Renderer.render(...)            (src/main.cpp:416-426)
    vfb                               -> the returned float32 [H, W, 3]  (src/main.cpp:50)
    takeScreenshot / Bitmap::saveImage-> save_image(path, rgb)           (src/sdl.cpp:103-116)

There is no CPU fallback: creating a Renderer without the CUDA library or without a GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import capi
from .capi import (HxrError, MODE_AUTO, MODE_MONTECARLO, MODE_WHITTED, RENDER_COUNT_TRAVERSAL, RENDER_ONE_LANE,  # noqa: F401
                   CFG_BRUTE_FORCE_MESHES, CFG_DEVICE_KD_BUILD)

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhexray_b200.so")
_api = None


def api():
    """The product library (CUDA). Raises if it has not been built."""
    global _api
    if _api is None:
        _api = capi.Api(_LIB_PATH)
    return _api


def data_root():
    """Directory holding the scene assets (`data/` of the reference: scenes, meshes, textures, cubemaps). HEXRAY_DATA
    overrides; otherwise the copy staged at <repo>/assets/data (by __graft_entry__.build()), else the reference checkout."""
    env = os.environ.get("HEXRAY_DATA")
    if env:
        return env
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.path.join(here, "assets", "data"), "/root/reference/data"):
        if os.path.isdir(cand):
            return cand
    raise FileNotFoundError("scene assets not found: set HEXRAY_DATA to the reference's data/ directory")


class SceneFile:
    """A parsed + flattened `.hexray` scene (host memory)."""

    def __init__(self, path, api_=None):
        self.api = api_ or api()
        self.handle = C.c_void_p()
        self.path = path
        st = self.api.lib.hxr_scene_load(path.encode(), C.byref(self.handle))
        if st != capi.HXR_OK:
            raise HxrError(st, self.api.last_error(None))

    @property
    def pod(self):
        return self.api.lib.hxr_scene_file_scene(self.handle)

    @property
    def settings(self):
        return self.pod.contents.settings

    def camera(self):
        cam = capi.Camera()
        self.api.check(self.api.lib.hxr_scene_file_camera(self.handle, C.byref(cam)))
        return cam

    def set_synthetic_mesh(self, mesh_index, kind, n, seed):
        self.api.check(self.api.lib.hxr_scene_file_set_synthetic_mesh(self.handle, mesh_index, kind.encode(), int(n), int(seed)))

    def write_obj(self, mesh_index, path):
        self.api.check(self.api.lib.hxr_scene_file_write_obj(self.handle, mesh_index, path.encode()))

    def close(self):
        if self.handle:
            self.api.lib.hxr_scene_file_free(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    """One GPU context (one process per GPU)."""

    def __init__(self, device=0, queue_capacity=0, api_=None, flags=0, devices=None):
        """devices: a list of CUDA ordinals for ONE context over several GPUs (the library shards every frame over them and
        sums the partial frames itself); default: the single GPU `device`."""
        self.api = api_ or api()
        self.ctx = C.c_void_p()
        if devices:
            self._devs = (C.c_int32 * len(devices))(*devices)
            cfg = capi.Config(device, flags, queue_capacity, len(devices), 0, self._devs)
        else:
            cfg = capi.Config(device, flags, queue_capacity, 0, 0, None)
        st = self.api.lib.hxr_create(C.byref(cfg), C.byref(self.ctx))
        if st != capi.HXR_OK:
            raise HxrError(st, self.api.last_error(None))
        self.scene = None

    def close(self):
        if self.ctx:
            self.api.lib.hxr_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        self.api.check(st, self.ctx)

    def load(self, scene_file):
        """Upload a SceneFile (and its camera)."""
        self.scene = scene_file
        self._check(self.api.lib.hxr_upload_scene(self.ctx, scene_file.pod))
        cam = scene_file.camera()
        self._check(self.api.lib.hxr_set_camera(self.ctx, C.byref(cam)))
        return self

    def set_camera(self, cam):
        self._check(self.api.lib.hxr_set_camera(self.ctx, C.byref(cam)))

    def _params(self, width, height, mode, spp, want_aa, max_depth, seed, shard, flags):
        return capi.RenderParams(width, height, mode, spp, want_aa, max_depth, seed, shard[0], shard[1], flags, 0)

    def frame_size(self, width=0, height=0):
        s = self.scene.settings
        return (width or s.frame_width, height or s.frame_height)

    def render(self, width=0, height=0, mode=MODE_AUTO, spp=0, want_aa=-1, max_depth=-1, seed=0, shard=(0, 1), flags=0, out=None):
        """Render into a host float32 [H, W, 3] array (un-clamped linear radiance). Returns (image, stats)."""
        W, H = self.frame_size(width, height)
        if out is None:
            out = np.empty((H, W, 3), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == W * H * 3
        p = self._params(width, height, mode, spp, want_aa, max_depth, seed, shard, flags)
        st = capi.Stats()
        self._check(self.api.lib.hxr_render(self.ctx, C.byref(p), out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st)))
        return out, st.as_dict()

    def render_device(self, dptr, width=0, height=0, mode=MODE_AUTO, spp=0, want_aa=-1, max_depth=-1, seed=0, shard=(0, 1), flags=0):
        """Render into DEVICE memory at `dptr` (W*H*3 floats on this context's GPU), e.g. a torch tensor's data_ptr()."""
        p = self._params(width, height, mode, spp, want_aa, max_depth, seed, shard, flags)
        st = capi.Stats()
        self._check(self.api.lib.hxr_render_device(self.ctx, C.byref(p), C.c_void_p(dptr), C.byref(st)))
        return st.as_dict()

    def resolve_device(self, dptr, width, height, spp):
        self._check(self.api.lib.hxr_resolve_device(self.ctx, C.c_void_p(dptr), width, height, spp))

    def trace_closest(self, rays):
        """rays: structured array (capi.RAY_DTYPE) or float [N, 6..8] (start, dir[, depth, flags]). Returns HIT_DTYPE array."""
        rays = _as_rays(rays)
        hits = np.zeros(len(rays), dtype=capi.HIT_DTYPE)
        self._check(self.api.lib.hxr_trace_closest(self.ctx, rays.ctypes.data, len(rays), hits.ctypes.data))
        return hits

    def trace_visible(self, segments):
        seg = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(len(seg), dtype=np.uint8)
        self._check(self.api.lib.hxr_trace_visible(self.ctx, seg.ctypes.data_as(C.POINTER(C.c_double)), len(seg),
                                                   out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out.astype(bool)

    def trace_color(self, rays):
        rays = _as_rays(rays)
        out = np.zeros((len(rays), 3), dtype=np.float32)
        self._check(self.api.lib.hxr_trace_color(self.ctx, rays.ctypes.data, len(rays), out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def save_frame_bmp(self, path, dptr=None, width=0, height=0):
        """takeScreenshot: the last rendered frame (or the device frame at `dptr`) as a BMP, converted to 8 bits on the GPU."""
        self._check(self.api.lib.hxr_save_frame_bmp(self.ctx, C.c_void_p(dptr) if dptr else None, width, height, path.encode()))

    def save_frame_exr(self, path, dptr=None, width=0, height=0):
        """Bitmap::saveEXR of the last rendered frame (or the device frame at `dptr`): float -> half on the GPU."""
        self._check(self.api.lib.hxr_save_frame_exr(self.ctx, C.c_void_p(dptr) if dptr else None, width, height, path.encode()))

    def progressive(self, n_passes, width=0, height=0, spp=0, seed=0, max_depth=-1, checkpoint=None):
        """Refine one Monte-Carlo frame pass by pass (hxr_progressive_*): yields (estimate [H, W, 3], stats) after every pass.
        checkpoint = (sum [H, W, 3] float32, passes_done, spp_done) from progressive_state() of an interrupted run of the same
        frame: the passes already done are not rendered again (hxr_progressive_resume)."""
        W, H = self.frame_size(width, height)
        p = self._params(width, height, MODE_MONTECARLO, spp, -1, max_depth, seed, (0, 0), 0)
        first = 0
        if checkpoint is None:
            self._check(self.api.lib.hxr_progressive_begin(self.ctx, C.byref(p), n_passes))
        else:
            s, first, spp_done = checkpoint
            s = np.ascontiguousarray(s, dtype=np.float32)
            assert s.size == W * H * 3
            self._check(self.api.lib.hxr_progressive_resume(self.ctx, C.byref(p), n_passes, s.ctypes.data_as(C.POINTER(C.c_float)), int(first), int(spp_done)))
        out = np.empty((H, W, 3), dtype=np.float32)
        for _ in range(first, n_passes):
            st = capi.Stats()
            self._check(self.api.lib.hxr_progressive_pass(self.ctx, out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st)))
            yield out, st.as_dict()

    def progressive_state(self, width=0, height=0):
        """The checkpoint of the progressive frame in flight: (un-normalised sum [H, W, 3], passes done, samples per pixel so far)."""
        W, H = self.frame_size(width, height)
        s = np.empty((H, W, 3), dtype=np.float32)
        passes, spp = C.c_int32(0), C.c_int32(0)
        self._check(self.api.lib.hxr_progressive_state(self.ctx, s.ctypes.data_as(C.POINTER(C.c_float)), C.byref(passes), C.byref(spp)))
        return s, passes.value, spp.value

    def reduce_backend(self):
        return self.api.lib.hxr_reduce_backend(self.ctx).decode()

    def set_profiling(self, on):
        self._check(self.api.lib.hxr_set_profiling(self.ctx, 1 if on else 0))

    def accel_info(self, mesh):
        info = capi.AccelInfo()
        self._check(self.api.lib.hxr_get_accel_info(self.ctx, mesh, C.byref(info)))
        return info.as_dict()


def _as_rays(rays):
    if isinstance(rays, np.ndarray) and rays.dtype == capi.RAY_DTYPE:
        return np.ascontiguousarray(rays)
    a = np.asarray(rays, dtype=np.float64)
    out = np.zeros(len(a), dtype=capi.RAY_DTYPE)
    out["start"] = a[:, 0:3]
    out["dir"] = a[:, 3:6]
    if a.shape[1] > 6:
        out["depth"] = a[:, 6].astype(np.int32)
    if a.shape[1] > 7:
        out["flags"] = a[:, 7].astype(np.uint32)
    return out


def save_image(path, rgb, api_=None):
    """Bitmap::saveImage: '.bmp' (8-bit via the reference's sRGB LUT) or '.exr' (half RGBA)."""
    a = api_ or api()
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    H, W = rgb.shape[:2]
    st = a.lib.hxr_save_image(path.encode(), rgb.ctypes.data_as(C.POINTER(C.c_float)), W, H)
    if st != capi.HXR_OK:
        raise HxrError(st, a.last_error(None))


def load_image(path, api_=None):
    """Bitmap::loadImage: returns float32 [H, W, 3]."""
    a = api_ or api()
    w, h = C.c_int32(), C.c_int32()
    st = a.lib.hxr_load_image(path.encode(), C.byref(w), C.byref(h), None, 0)
    if st != capi.HXR_OK:
        raise HxrError(st, a.last_error(None))
    out = np.empty((h.value, w.value, 3), dtype=np.float32)
    st = a.lib.hxr_load_image(path.encode(), C.byref(w), C.byref(h), out.ctypes.data_as(C.POINTER(C.c_float)), out.size)
    if st != capi.HXR_OK:
        raise HxrError(st, a.last_error(None))
    return out


def render_file(path, device=0, **kw):
    """parse + upload + render one frame; returns (image, stats)."""
    sf = SceneFile(path)
    r = Renderer(device=device)
    try:
        r.load(sf)
        return r.render(**kw)
    finally:
        r.close()
        sf.close()
