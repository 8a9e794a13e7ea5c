"""ctypes binding of include/hxr.h.

`Api(path)` binds one shared library. The package-level entry point (`hexray_b200.api()`) binds
ONLY hexray_b200/libhexray_b200.so — the CUDA product — and raises if it is missing; there is no
CPU fallback. tests/ additionally bind tests/emu/libhxr_emu.so (a host build of the same per-ray
functions) through this same class to check logic on the CPU tier.
"""
import ctypes as C
import os

import numpy as np

c_i32, c_u32, c_u64, c_f32, c_f64 = C.c_int32, C.c_uint32, C.c_uint64, C.c_float, C.c_double

HXR_OK = 0
STATUS_NAMES = {0: "HXR_OK", -1: "HXR_ERR_INVALID", -2: "HXR_ERR_NO_DEVICE", -3: "HXR_ERR_CUDA",
                -4: "HXR_ERR_PARSE", -5: "HXR_ERR_IO", -6: "HXR_ERR_OVERFLOW"}
MODE_AUTO, MODE_WHITTED, MODE_MONTECARLO = 0, 1, 2
RENDER_COUNT_TRAVERSAL = 1
RENDER_ONE_LANE = 2
CFG_BRUTE_FORCE_MESHES = 1
CFG_DEVICE_KD_BUILD = 2


class HxrError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), message))
        self.status = status


class Transform(C.Structure):
    _fields_ = [("offset", c_f64 * 3), ("m", c_f64 * 9), ("inv", c_f64 * 9), ("inv_t", c_f64 * 9)]


class Geometry(C.Structure):
    _fields_ = [("type", c_i32), ("a", c_i32), ("b", c_i32), ("c", c_i32), ("p", c_f64 * 6)]


class Triangle(C.Structure):
    _fields_ = [("v", c_i32 * 3), ("n", c_i32 * 3), ("t", c_i32 * 3), ("pad", c_i32), ("gnormal", c_f64 * 3),
                ("ab", c_f64 * 3), ("ac", c_f64 * 3), ("ab_cross_ac", c_f64 * 3), ("dndx", c_f64 * 3), ("dndy", c_f64 * 3)]


class Mesh(C.Structure):
    _fields_ = [("n_vertices", c_i32), ("n_normals", c_i32), ("n_uvs", c_i32), ("n_triangles", c_i32),
                ("vertices", C.POINTER(c_f64)), ("normals", C.POINTER(c_f64)), ("uvs", C.POINTER(c_f64)),
                ("triangles", C.POINTER(Triangle)), ("faceted", c_i32), ("backface_culling", c_i32),
                ("bbox_min", c_f64 * 3), ("bbox_max", c_f64 * 3)]


class Heightfield(C.Structure):
    _fields_ = [("width", c_i32), ("height", c_i32), ("use_optimization", c_i32), ("max_k", c_i32),
                ("heights", C.POINTER(c_f32)), ("max_h", C.POINTER(c_f32)), ("normals", C.POINTER(c_f64)),
                ("high_map", C.POINTER(c_f32)), ("bbox_min", c_f64 * 3), ("bbox_max", c_f64 * 3)]


class Node(C.Structure):
    _fields_ = [("geom", c_i32), ("shader", c_i32), ("bump_tex", c_i32), ("pad", c_i32), ("T", Transform)]


class Shader(C.Structure):
    _fields_ = [("type", c_i32), ("tex", c_i32), ("first_layer", c_i32), ("n_layers", c_i32), ("i0", c_i32),
                ("f0", c_f32), ("color", c_f32 * 3), ("color2", c_f32 * 3), ("ior", c_f64)]


class Layer(C.Structure):
    _fields_ = [("shader", c_i32), ("tex", c_i32), ("blend", c_f32 * 3)]


class Texture(C.Structure):
    _fields_ = [("type", c_i32), ("image", c_i32), ("color1", c_f32 * 3), ("color2", c_f32 * 3),
                ("scaling", c_f64), ("strength", c_f64), ("ior", c_f64)]


class Image(C.Structure):
    _fields_ = [("width", c_i32), ("height", c_i32), ("rgb", C.POINTER(c_f32))]


class Light(C.Structure):
    _fields_ = [("type", c_i32), ("xsubd", c_i32), ("ysubd", c_i32), ("power", c_f32), ("color", c_f32 * 3),
                ("scale_factor", c_f32), ("area", c_f64), ("pos", c_f64 * 3), ("T", Transform)]


class Settings(C.Structure):
    _fields_ = [("frame_width", c_i32), ("frame_height", c_i32), ("max_trace_depth", c_i32), ("want_aa", c_i32),
                ("gi", c_i32), ("num_paths", c_i32), ("ambient", c_f32 * 3), ("background", c_f32 * 3)]


class Scene(C.Structure):
    _fields_ = [("abi_version", c_i32), ("n_nodes", c_i32), ("n_geometries", c_i32), ("n_meshes", c_i32),
                ("n_heightfields", c_i32), ("n_shaders", c_i32), ("n_layers", c_i32), ("n_textures", c_i32),
                ("n_images", c_i32), ("n_lights", c_i32), ("has_environment", c_i32), ("env_images", c_i32 * 6),
                ("nodes", C.POINTER(Node)), ("geometries", C.POINTER(Geometry)), ("meshes", C.POINTER(Mesh)),
                ("heightfields", C.POINTER(Heightfield)), ("shaders", C.POINTER(Shader)), ("layers", C.POINTER(Layer)),
                ("textures", C.POINTER(Texture)), ("images", C.POINTER(Image)), ("lights", C.POINTER(Light)),
                ("settings", Settings)]


class Camera(C.Structure):
    _fields_ = [("pos", c_f64 * 3), ("top_left", c_f64 * 3), ("top_right", c_f64 * 3), ("bottom_left", c_f64 * 3),
                ("up", c_f64 * 3), ("right", c_f64 * 3), ("front", c_f64 * 3), ("aperture_size", c_f64),
                ("focal_plane_dist", c_f64), ("stereo_separation", c_f64), ("dof", c_i32), ("auto_focus", c_i32),
                ("num_samples", c_i32), ("pad", c_i32)]


class Config(C.Structure):
    _fields_ = [("device", c_i32), ("flags", c_i32), ("queue_capacity", c_u64), ("n_devices", c_i32), ("reserved", c_i32),
                ("devices", C.POINTER(c_i32))]


class RenderParams(C.Structure):
    _fields_ = [("width", c_i32), ("height", c_i32), ("mode", c_i32), ("spp", c_i32), ("want_aa", c_i32),
                ("max_depth", c_i32), ("seed", c_u64), ("shard_index", c_i32), ("shard_count", c_i32),
                ("flags", c_i32), ("reserved", c_i32)]


class Stats(C.Structure):
    _fields_ = [("rays_closest", c_u64), ("rays_shadow", c_u64), ("kd_inner", c_u64), ("kd_leaves", c_u64),
                ("tri_tests", c_u64), ("mesh_queries", c_u64), ("kernel_launches", c_u64), ("render_ms", c_f64),
                ("trace_closest_ms", c_f64), ("trace_shadow_ms", c_f64), ("shade_ms", c_f64), ("other_ms", c_f64),
                ("trace_closest_launches", c_u64), ("trace_shadow_launches", c_u64), ("spp_done", c_u32),
                ("aa_pixels", c_u32), ("walk_ms", c_f64), ("walk_launches", c_u64), ("cand_overflow", c_u64),
                ("shadow_resolve_ms", c_f64), ("gen_ms", c_f64), ("setup_ms", c_f64), ("finish_ms", c_f64), ("reduce_ms", c_f64), ("n_devices", c_u32), ("reserved", c_u32)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class AccelInfo(C.Structure):
    _fields_ = [("nodes", c_u64), ("leaves", c_u64), ("tri_refs", c_u64), ("bytes_nodes", c_u64), ("bytes_tris", c_u64),
                ("max_depth", c_u32), ("n_triangles", c_u32), ("build_ms", c_f64), ("from_cache", c_u32), ("device_build", c_u32), ("device_ms", c_f64)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


RAY_DTYPE = np.dtype([("start", "<f8", 3), ("dir", "<f8", 3), ("depth", "<i4"), ("flags", "<u4")])
HIT_DTYPE = np.dtype([("status", "<i4"), ("node", "<i4"), ("dist", "<f8"), ("ip", "<f8", 3), ("norm", "<f8", 3),
                      ("u", "<f8"), ("v", "<f8"), ("dndx", "<f8", 3), ("dndy", "<f8", 3), ("color", "<f4", 3), ("pad", "<f4")])
assert RAY_DTYPE.itemsize == 56 and HIT_DTYPE.itemsize == 144

# every symbol include/hxr.h declares (tests check that the built library exports them all)
SYMBOLS = ["hxr_create", "hxr_destroy", "hxr_last_error", "hxr_upload_scene", "hxr_set_camera", "hxr_render",
           "hxr_render_device", "hxr_resolve_device", "hxr_trace_closest", "hxr_trace_visible", "hxr_trace_color",
           "hxr_get_accel_info", "hxr_set_profiling", "hxr_scene_load", "hxr_scene_file_scene", "hxr_scene_file_camera",
           "hxr_scene_file_set_synthetic_mesh", "hxr_scene_file_write_obj", "hxr_scene_file_free", "hxr_save_image", "hxr_load_image",
           "hxr_test_tri_filter", "hxr_test_tri_filter_packed", "hxr_save_frame_bmp", "hxr_save_frame_exr", "hxr_device_count", "hxr_reduce_backend",
           "hxr_progressive_begin", "hxr_progressive_pass", "hxr_progressive_state", "hxr_progressive_resume"]


class Api:
    """One loaded library with typed entry points."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s not found. Build it with `python -m hexray_b200.build` (needs nvcc); hexray_b200 has no CPU fallback." % path)
        self.path = path
        self.lib = L = C.CDLL(path)
        vp = C.c_void_p
        L.hxr_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.hxr_destroy.argtypes = [vp]
        L.hxr_destroy.restype = None
        L.hxr_last_error.argtypes = [vp]
        L.hxr_last_error.restype = C.c_char_p
        L.hxr_upload_scene.argtypes = [vp, C.POINTER(Scene)]
        L.hxr_set_camera.argtypes = [vp, C.POINTER(Camera)]
        L.hxr_render.argtypes = [vp, C.POINTER(RenderParams), C.POINTER(c_f32), C.POINTER(Stats)]
        L.hxr_render_device.argtypes = [vp, C.POINTER(RenderParams), vp, C.POINTER(Stats)]
        L.hxr_resolve_device.argtypes = [vp, vp, c_i32, c_i32, c_i32]
        L.hxr_trace_closest.argtypes = [vp, vp, C.c_size_t, vp]
        L.hxr_trace_visible.argtypes = [vp, C.POINTER(c_f64), C.c_size_t, C.POINTER(C.c_uint8)]
        L.hxr_trace_color.argtypes = [vp, vp, C.c_size_t, C.POINTER(c_f32)]
        L.hxr_set_profiling.argtypes = [vp, c_i32]
        L.hxr_get_accel_info.argtypes = [vp, c_i32, C.POINTER(AccelInfo)]
        L.hxr_scene_load.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.hxr_scene_file_scene.argtypes = [vp]
        L.hxr_scene_file_scene.restype = C.POINTER(Scene)
        L.hxr_scene_file_camera.argtypes = [vp, C.POINTER(Camera)]
        L.hxr_scene_file_set_synthetic_mesh.argtypes = [vp, c_i32, C.c_char_p, C.c_int64, c_u64]
        L.hxr_scene_file_write_obj.argtypes = [vp, c_i32, C.c_char_p]
        L.hxr_scene_file_free.argtypes = [vp]
        L.hxr_scene_file_free.restype = None
        L.hxr_save_image.argtypes = [C.c_char_p, C.POINTER(c_f32), c_i32, c_i32]
        L.hxr_save_frame_bmp.argtypes = [vp, vp, c_i32, c_i32, C.c_char_p]
        L.hxr_save_frame_exr.argtypes = [vp, vp, c_i32, c_i32, C.c_char_p]
        L.hxr_progressive_begin.argtypes = [vp, C.POINTER(RenderParams), c_i32]
        L.hxr_progressive_pass.argtypes = [vp, C.POINTER(c_f32), C.POINTER(Stats)]
        L.hxr_progressive_state.argtypes = [vp, C.POINTER(c_f32), C.POINTER(c_i32), C.POINTER(c_i32)]
        L.hxr_progressive_resume.argtypes = [vp, C.POINTER(RenderParams), c_i32, C.POINTER(c_f32), c_i32, c_i32]
        L.hxr_device_count.argtypes = []
        L.hxr_reduce_backend.argtypes = [vp]
        L.hxr_reduce_backend.restype = C.c_char_p
        L.hxr_test_tri_filter.argtypes = [C.c_size_t, vp, vp, vp, c_i32, vp, vp, vp, vp]
        L.hxr_test_tri_filter_packed.argtypes = [C.c_size_t, vp, vp, vp, c_i32, vp, vp, vp, vp]
        L.hxr_load_image.argtypes = [C.c_char_p, C.POINTER(c_i32), C.POINTER(c_i32), C.POINTER(c_f32), C.c_size_t]
        for name in SYMBOLS:
            if getattr(L, name).restype is C.c_int:
                getattr(L, name).restype = C.c_int

    def last_error(self, ctx=None):
        msg = self.lib.hxr_last_error(ctx)
        return msg.decode("utf-8", "replace") if msg else ""

    def check(self, status, ctx=None):
        if status != HXR_OK:
            raise HxrError(status, self.last_error(ctx))
