/* hxr.h — C ABI of the B200 render hot path (libhexray_b200.so).
 *
 * The reference (anrieff/hexray) has no plugin/FFI boundary: its integrators read a
 * global `Scene scene` (reference src/scene.cpp:811, scene.h:271-294) and write a
 * global `Color vfb[1920][1920]` (src/main.cpp:50). This header is the seam a
 * maintainer would bind instead:
 *
 *   scene graph after parseScene()/beginRender()/beginFrame()   ->  hxr_scene (POD)
 *   Camera::beginFrame() results (src/camera.cpp:30-63)         ->  hxr_camera
 *   render(false) (src/main.cpp:416-426)                        ->  hxr_render()
 *   TraceContext::raycast (src/main.cpp:63-103)                 ->  hxr_trace_closest()
 *   visible() (src/main.cpp:171-190)                            ->  hxr_trace_visible()
 *   takeScreenshot -> Bitmap::saveImage (src/sdl.cpp:103-116,
 *                     src/bitmap.cpp:202-288)                   ->  hxr_save_image()
 *   scene.parseScene() (src/scene.cpp:735)                      ->  hxr_scene_load()
 *
 * Conventions: plain C types only; caller owns every input buffer and the library
 * copies what it needs before returning; every call returns HXR_OK (0) or a negative
 * hxr_status; no C++ exception or abort crosses this boundary; a context is driven by
 * one host thread. Vectors are double[3] (reference `Vector`, src/vector.h:30), colours
 * float[3] (reference `Color`, src/color.h:62). Matrices are row-major 3x3 used with
 * ROW vectors, v' = v * M (reference src/matrix.h:53-60).
 */
#ifndef HXR_H
#define HXR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HXR_ABI_VERSION 2

typedef enum hxr_status {
    HXR_OK = 0,
    HXR_ERR_INVALID = -1,    /* bad argument / inconsistent scene description */
    HXR_ERR_NO_DEVICE = -2,  /* no usable CUDA device: there is NO CPU fallback */
    HXR_ERR_CUDA = -3,       /* a CUDA call failed; see hxr_last_error */
    HXR_ERR_PARSE = -4,      /* .hexray syntax error / missing asset (reference SyntaxError / FileNotFoundError) */
    HXR_ERR_IO = -5,
    HXR_ERR_OVERFLOW = -6    /* a ray queue overflowed even at the smallest batch */
} hxr_status;

/* ---------------------------------------------------------------- scene POD */

/* reference Transform (src/matrix.h:62-99): offset + m + invM + transpose(invM) */
typedef struct hxr_transform {
    double offset[3];
    double m[9];
    double inv[9];
    double inv_t[9];
} hxr_transform;

typedef enum hxr_geom_type {
    HXR_GEOM_PLANE = 0,       /* p[0]=y, p[1]=limit                       (src/geometry.h:57-66)  */
    HXR_GEOM_SPHERE = 1,      /* p[0..2]=O, p[3]=R, p[4]=uvscaling        (src/geometry.h:68-80)  */
    HXR_GEOM_CUBE = 2,        /* p[0..2]=O, p[3]=half side                (src/geometry.h:82-106) */
    HXR_GEOM_CSG = 3,         /* a=hxr_csg_op, b=left geom, c=right geom  (src/geometry.h:108-136)*/
    HXR_GEOM_MESH = 4,        /* a=index into hxr_scene.meshes            (src/mesh.h:72-124)     */
    HXR_GEOM_HEIGHTFIELD = 5  /* a=index into hxr_scene.heightfields      (src/heightfield.h)     */
} hxr_geom_type;

typedef enum hxr_csg_op { HXR_CSG_UNION = 0, HXR_CSG_INTER = 1, HXR_CSG_DIFF = 2 } hxr_csg_op;

typedef struct hxr_geometry {
    int32_t type;
    int32_t a, b, c;
    double p[6];
} hxr_geometry;

/* reference Triangle (src/bbox.h:36-43) after Mesh::prepareTriangles (src/mesh.cpp:370-396) */
typedef struct hxr_triangle {
    int32_t v[3], n[3], t[3];
    int32_t pad;
    double gnormal[3];
    double ab[3], ac[3], ab_cross_ac[3];
    double dndx[3], dndy[3];
} hxr_triangle;

/* reference Mesh after beginRender (src/mesh.cpp:49-87): slot 0 of vertices/normals/uvs
 * is the OBJ sentinel (src/mesh.cpp:307-310). The KD-tree is NOT part of the ABI: the
 * library builds its own during hxr_upload_scene - on the host cores, or on the GPU
 * with HXR_CFG_DEVICE_KD_BUILD. */
typedef struct hxr_mesh {
    int32_t n_vertices, n_normals, n_uvs, n_triangles;
    const double* vertices; /* 3 per entry */
    const double* normals;  /* 3 per entry */
    const double* uvs;      /* 3 per entry (third is 0) */
    const hxr_triangle* triangles;
    int32_t faceted;
    int32_t backface_culling;
    double bbox_min[3], bbox_max[3];
} hxr_mesh;

/* reference Heightfield after fillProperties (src/heightfield.cpp:173-261) */
typedef struct hxr_heightfield {
    int32_t width, height;
    int32_t use_optimization;
    int32_t max_k;
    const float* heights;   /* W*H                                    */
    const float* max_h;     /* W*H: max over the texel's 2x2 corners  */
    const double* normals;  /* W*H*3                                  */
    const float* high_map;  /* W*H*16 (only [0, max_k) used) or NULL  */
    double bbox_min[3], bbox_max[3];
} hxr_heightfield;

typedef struct hxr_node { /* reference Node (src/node.h:31-47) */
    int32_t geom;
    int32_t shader;
    int32_t bump_tex; /* -1: none */
    int32_t pad;
    hxr_transform T;
} hxr_node;

typedef enum hxr_shader_type {
    HXR_SHADER_LAMBERT = 0,    /* color, tex                               (src/shading.h:97-111)  */
    HXR_SHADER_PHONG = 1,      /* color, tex, color2=specular, f0=exponent (src/shading.h:113-124) */
    HXR_SHADER_REFLECTION = 2, /* color=reflColor, f0=glossiness, i0=numSamples (shading.h:127-146)*/
    HXR_SHADER_REFRACTION = 3, /* color=refrColor, ior                     (src/shading.h:149-165) */
    HXR_SHADER_LAYERED = 4,    /* layers [first_layer, first_layer+n_layers) (shading.h:167-179)   */
    HXR_SHADER_CONST = 5       /* color                                    (src/shading.h:206-214) */
} hxr_shader_type;

typedef struct hxr_shader {
    int32_t type;
    int32_t tex; /* -1: none */
    int32_t first_layer, n_layers;
    int32_t i0;
    float f0;
    float color[3];
    float color2[3];
    double ior;
} hxr_shader;

typedef struct hxr_layer {
    int32_t shader;
    int32_t tex; /* blend texture, -1: none */
    float blend[3];
} hxr_layer;

typedef enum hxr_texture_type {
    HXR_TEX_CHECKER = 0, /* color1, color2, scaling                     (src/shading.h:58-69)   */
    HXR_TEX_BITMAP = 1,  /* image, scaling (gamma already applied)       (src/shading.h:72-90)   */
    HXR_TEX_FRESNEL = 2, /* ior                                          (src/shading.h:181-190) */
    HXR_TEX_BUMP = 3,    /* image (already differentiated), strength, scaling (shading.h:192-204)*/
    HXR_TEX_BUMPS = 4    /* strength                                     (src/shading.h:217-227) */
} hxr_texture_type;

typedef struct hxr_texture {
    int32_t type;
    int32_t image; /* -1: none */
    float color1[3];
    float color2[3];
    double scaling;
    double strength;
    double ior;
} hxr_texture;

typedef struct hxr_image { /* reference Bitmap (src/bitmap.h): float RGB, row-major, top-down */
    int32_t width, height;
    const float* rgb;
} hxr_image;

typedef enum hxr_light_type { HXR_LIGHT_POINT = 0, HXR_LIGHT_RECT = 1 } hxr_light_type;

typedef struct hxr_light { /* reference PointLight / RectLight (src/lights.h:79-118) */
    int32_t type;
    int32_t xsubd, ysubd;
    float power;
    float color[3];
    float scale_factor; /* RectLight::beginFrame: 1/area (src/lights.cpp:75-88); 1 for point */
    double area;
    double pos[3];
    hxr_transform T;
} hxr_light;

typedef struct hxr_settings { /* reference GlobalSettings (src/scene.h:244-269) */
    int32_t frame_width, frame_height;
    int32_t max_trace_depth;
    int32_t want_aa;
    int32_t gi;
    int32_t num_paths;
    float ambient[3];
    float background[3];
} hxr_settings;

typedef struct hxr_scene {
    int32_t abi_version; /* HXR_ABI_VERSION */
    int32_t n_nodes, n_geometries, n_meshes, n_heightfields, n_shaders, n_layers;
    int32_t n_textures, n_images, n_lights;
    int32_t has_environment;  /* CubemapEnvironment present AND loaded (src/environment.cpp:64) */
    int32_t env_images[6];    /* NEGX NEGY NEGZ POSX POSY POSZ (src/environment.h:32-39) */
    const hxr_node* nodes;    /* scene.nodes order (shader-less nodes already removed) */
    const hxr_geometry* geometries;
    const hxr_mesh* meshes;
    const hxr_heightfield* heightfields;
    const hxr_shader* shaders;
    const hxr_layer* layers;
    const hxr_texture* textures;
    const hxr_image* images;
    const hxr_light* lights;
    hxr_settings settings;
} hxr_scene;

typedef struct hxr_camera { /* results of Camera::beginFrame (src/camera.cpp:30-63) */
    double pos[3];
    double top_left[3], top_right[3], bottom_left[3];
    double up[3], right[3], front[3];
    double aperture_size;
    double focal_plane_dist;
    double stereo_separation;
    int32_t dof;
    int32_t auto_focus;
    int32_t num_samples;
    int32_t pad;
} hxr_camera;

/* ---------------------------------------------------------------- rendering */

/* test hook: test every triangle of every mesh in index order instead of walking the KD-tree — the reference's
 * `useKDTree false` path (src/mesh.cpp:255-262); results are identical, only (much) slower */
#define HXR_CFG_BRUTE_FORCE_MESHES 1
#define HXR_CFG_DEVICE_KD_BUILD 2 /* build the meshes' KD-trees on the GPU (level-synchronous SAH build) instead of on the host cores;
                                    * the environment variable HXR_KD_BUILD=device|host overrides the flag */

/* One context = one or several GPUs of the box, all owned by the library (reference: the ThreadPool of src/threading.cpp:54-97
 * that main() creates at src/main.cpp:546). With n_devices > 1 hxr_render shards every frame over the GPUs (Monte-Carlo:
 * sample passes; Whitted: row bands), sums the partial frames on the first GPU - one kernel reading the peers' buffers over
 * NVLink, or an NCCL reduce on a communicator the library creates with ncclCommInitAll (HXR_REDUCE=nccl) - resolves and
 * returns the finished frame; the scene is uploaded to every GPU, its KD-trees are built once. */
typedef struct hxr_config {
    int32_t device;           /* CUDA device ordinal (used when n_devices == 0) */
    int32_t flags;            /* HXR_CFG_* */
    uint64_t queue_capacity;  /* ray-queue capacity in rays per GPU; 0 = default */
    int32_t n_devices;        /* 0 / 1: one GPU (`device`); > 1: that many GPUs */
    int32_t reserved;
    const int32_t* devices;   /* n_devices ordinals, or NULL for 0 .. n_devices - 1 */
} hxr_config;

typedef enum hxr_render_mode {
    HXR_MODE_AUTO = 0,    /* what reference render() would pick (src/main.cpp:416-426) */
    HXR_MODE_WHITTED = 1, /* renderWithoutMonteCarlo */
    HXR_MODE_MONTECARLO = 2
} hxr_render_mode;

typedef struct hxr_render_params {
    int32_t width, height;      /* 0 = scene settings */
    int32_t mode;               /* hxr_render_mode */
    int32_t spp;                /* Monte-Carlo samples per pixel for the WHOLE job; 0 = scene default */
    int32_t want_aa;            /* -1 = scene setting */
    int32_t max_depth;          /* -1 = scene setting */
    uint64_t seed;
    /* sharding across ranks (one process per GPU):
     *   Monte-Carlo: this rank renders samples {s : s % shard_count == shard_index} of every pixel
     *   Whitted:     this rank renders rows    {y : (y / 16) % shard_count == shard_index}
     * and hxr_render returns the rank's PARTIAL SUM (zeros elsewhere), un-normalised by spp,
     * when shard_count > 1; the caller reduces and calls hxr_resolve. shard_count 0/1 = whole frame. */
    int32_t shard_index, shard_count;
    int32_t flags;              /* HXR_RENDER_* */
    int32_t reserved;
} hxr_render_params;

#define HXR_RENDER_COUNT_TRAVERSAL 1 /* also count KD inner-node visits / triangle tests (slower) */
#define HXR_RENDER_ONE_LANE 2        /* queue every kernel of the frame on ONE stream (default: the shadow chain of a bounce runs on a
                                      * second stream beside the next bounce's closest-hit chain); the per-kernel times of hxr_stats
                                      * are exclusive only in this mode */

typedef struct hxr_stats {
    uint64_t rays_closest;  /* closest-hit queries past the depth guard (src/main.cpp:65) */
    uint64_t rays_shadow;   /* visible() queries (src/main.cpp:171) */
    uint64_t kd_inner, kd_leaves, tri_tests, mesh_queries; /* only with HXR_RENDER_COUNT_TRAVERSAL */
    uint64_t kernel_launches;
    double render_ms;       /* CUDA-event time of the whole render on the context's stream */
    /* with profiling on: closest-hit walk / shadow walk + shadow resolve / shade (exact tests + shading + emission) / the rest */
    double trace_closest_ms, trace_shadow_ms, shade_ms, other_ms;
    uint64_t trace_closest_launches, trace_shadow_launches;
    uint32_t spp_done;
    uint32_t aa_pixels;
    double walk_ms;         /* device time of the KD-walk kernel launches (closest-hit + shadow) */
    uint64_t walk_launches;
    uint64_t cand_overflow; /* rays whose candidate record filled up during the walk (finished by a second, exact-on-the-spot walk) */
    double shadow_resolve_ms, gen_ms, setup_ms, finish_ms; /* with profiling on: the shadow-resolve, primary-ray, inline-setup and overflow-finish kernels */
    double reduce_ms;       /* multi-GPU contexts: summing + resolving the partial frames on the first GPU (inside render_ms) */
    uint32_t n_devices;     /* GPUs that rendered this frame */
    uint32_t reserved;
} hxr_stats;

typedef struct hxr_ray {
    double start[3];
    double dir[3];
    int32_t depth;
    uint32_t flags; /* reference RayFlags (src/vector.h:174-177) */
} hxr_ray;

/* one raycast() result: reference IntersectionInfo (src/geometry.h:33-40) + closest node */
typedef struct hxr_hit {
    int32_t status; /* 0 = surface hit, 1 = early colour (depth guard, light, miss) */
    int32_t node;   /* index into scene.nodes, -1 if none */
    double dist;
    double ip[3];
    double norm[3]; /* after bump modifyNormal, as raycast returns it */
    double u, v;
    double dndx[3], dndy[3];
    float color[3]; /* early colour when status == 1 */
    float pad;
} hxr_hit;

typedef struct hxr_ctx hxr_ctx;

int hxr_create(const hxr_config* cfg, hxr_ctx** out);
void hxr_destroy(hxr_ctx* ctx);
/* message of the last failed call on this context (ctx == NULL: last failed hxr_create / loader call) */
const char* hxr_last_error(const hxr_ctx* ctx);

/* number of usable CUDA devices in this process (0: none) */
int hxr_device_count(void);
/* how a multi-GPU context sums its partial frames: "peer", "nccl", "host" ("none" for one GPU) */
const char* hxr_reduce_backend(const hxr_ctx* ctx);

int hxr_upload_scene(hxr_ctx* ctx, const hxr_scene* scene);
int hxr_set_camera(hxr_ctx* ctx, const hxr_camera* cam);

/* rgb_out: caller-allocated HOST buffer, width*height*3 floats, row-major top-down,
 * un-clamped linear radiance — the same content as the reference's vfb. Blocking. */
int hxr_render(hxr_ctx* ctx, const hxr_render_params* p, float* rgb_out, hxr_stats* stats);
/* same, but the result stays in DEVICE memory (d_rgb: width*height*3 floats on the context's
 * device) so that a caller can reduce it across ranks (NCCL) without a host round trip. */
int hxr_render_device(hxr_ctx* ctx, const hxr_render_params* p, void* d_rgb, hxr_stats* stats);
/* scale a reduced device sum buffer by 1/spp in place (Monte-Carlo resolve, src/main.cpp:376). */
int hxr_resolve_device(hxr_ctx* ctx, void* d_rgb, int32_t width, int32_t height, int32_t spp);

/* ---- progressive rendering (reference: the interactive loop of src/main.cpp:387-414, 443-504 - gameloop / coarseRender -
 * that shows a coarse frame first and refines it; here refinement = more sample passes of the same Monte-Carlo frame).
 * hxr_progressive_begin fixes the frame (size, total spp, seed) and the number of passes; every hxr_progressive_pass renders
 * the next pass (the samples s with (s / n_gpus) % n_passes == pass, dealt over the context's GPUs as usual), reduces it
 * across the GPUs (ONE reduce per pass), adds it to the running per-pixel sum on the first GPU and returns the current
 * estimate sum / samples_so_far in rgb_out (host, width*height*3 floats; may be NULL to skip the read-back). After the last
 * pass the estimate equals what hxr_render returns for the same parameters (up to FP32 summation order). A camera change
 * between frames is hxr_set_camera + hxr_progressive_begin. hxr_progressive_state exports the checkpoint (sum, samples):
 * sum_out may be NULL; returns the number of passes done through *passes_done and the samples per pixel so far through *spp_done.
 * hxr_progressive_resume is hxr_progressive_begin continued from such a checkpoint (a frame interrupted after passes_done
 * passes, possibly by another process): same scene, camera, parameters, n_passes and number of GPUs as the run that wrote it -
 * spp_done is checked against the plan, a checkpoint of another plan is refused with HXR_ERR_INVALID; the next
 * hxr_progressive_pass renders pass number passes_done. */
int hxr_progressive_begin(hxr_ctx* ctx, const hxr_render_params* p, int32_t n_passes);
int hxr_progressive_pass(hxr_ctx* ctx, float* rgb_out, hxr_stats* stats);
int hxr_progressive_state(hxr_ctx* ctx, float* sum_out, int32_t* passes_done, int32_t* spp_done);
int hxr_progressive_resume(hxr_ctx* ctx, const hxr_render_params* p, int32_t n_passes, const float* sum, int32_t passes_done, int32_t spp_done);

/* when on, every kernel launch of this context is bracketed by CUDA events and hxr_stats carries the
 * per-kernel-class device times (trace_closest_ms, trace_shadow_ms, shade_ms, other_ms) */
int hxr_set_profiling(hxr_ctx* ctx, int32_t on);

/* test hooks: explicit rays in, raycast()/visible() results out (host buffers) */
int hxr_trace_closest(hxr_ctx* ctx, const hxr_ray* rays, size_t n, hxr_hit* hits);
int hxr_trace_visible(hxr_ctx* ctx, const double* segments /* n*6: A,B */, size_t n, uint8_t* visible);
/* raytrace() (Whitted) colour for explicit rays: n*3 floats out */
int hxr_trace_color(hxr_ctx* ctx, const hxr_ray* rays, size_t n, float* rgb);

/* test hook (pure host code, no context): the walk's conservative FP32 triangle filter (csrc/device/isect.h: tri_filter)
 * next to the reference's exact double test (tri_core) on n independent (ray, triangle) pairs.
 *   rays: n*6 (origin, unit direction), tris: n*9 (vertices A, B, C), tbest: n (best parameter known so far)
 *   cls_out: 0 MISS / 1 MAYBE / 2 CERTAIN, ghi_out: the filter's upper bound for CERTAIN, exact_out: 1 if the exact test
 *   accepts the pair at a parameter <= tbest, gamma_out: that parameter */
int hxr_test_tri_filter(size_t n, const double* rays, const double* tris, const double* tbest, int32_t backface_culling,
                        int32_t* cls_out, float* ghi_out, int32_t* exact_out, double* gamma_out);
/* the same through the 32-byte packed triangle record the walk reads by default (tri_filter_packed) */
int hxr_test_tri_filter_packed(size_t n, const double* rays, const double* tris, const double* tbest, int32_t backface_culling,
                               int32_t* cls_out, float* ghi_out, int32_t* exact_out, double* gamma_out);

/* acceleration-structure facts for reporting (per mesh): nodes, leaves, max depth, tri refs, build ms */
typedef struct hxr_accel_info {
    uint64_t nodes, leaves, tri_refs, bytes_nodes, bytes_tris;
    uint32_t max_depth;
    uint32_t n_triangles;
    double build_ms;        /* of the build that produced the tree (possibly in an earlier process: see from_cache) */
    uint32_t from_cache;    /* 1: the tree came from the on-disk cache, or from another process that was building it */
    uint32_t device_build;  /* 1: the tree was built on the GPU (HXR_CFG_DEVICE_KD_BUILD) */
    double device_ms;       /* device-built trees: the part of build_ms spent in the GPU passes (the rest packs the blocks on the host) */
} hxr_accel_info;
int hxr_get_accel_info(hxr_ctx* ctx, int32_t mesh, hxr_accel_info* out);

/* ---------------------------------------------------------------- host front-end
 * Pure host code living in the same library: the .hexray parser and element model
 * (drop-in for scene.parseScene + beginRender + beginFrame) that produces the POD above. */
typedef struct hxr_scene_file hxr_scene_file;

int hxr_scene_load(const char* path, hxr_scene_file** out);
const hxr_scene* hxr_scene_file_scene(const hxr_scene_file* sf);
/* camera for a given frame size (Camera::beginFrame depends on frameWidth/Height only via the caller) */
int hxr_scene_file_camera(const hxr_scene_file* sf, hxr_camera* out);
/* procedural meshes for the large-scene configuration: replaces mesh `mesh_index`
 * (or appends when -1) with a generated one. kind: "terrain" (n = grid vertices per side) or "soup" (n = triangles) */
int hxr_scene_file_set_synthetic_mesh(hxr_scene_file* sf, int32_t mesh_index, const char* kind, int64_t n, uint64_t seed);
/* write mesh `mesh_index` (counting Mesh geometries in scene order) as a Wavefront OBJ, so that the
 * reference renderer can load a procedural mesh for side-by-side timing */
int hxr_scene_file_write_obj(const hxr_scene_file* sf, int32_t mesh_index, const char* path);
void hxr_scene_file_free(hxr_scene_file* sf);

/* Bitmap::saveImage equivalent: ".bmp" (8-bit through the reference's sRGB LUT, src/color.h:36-47 +
 * src/sdl.cpp:404-419) or ".exr" (HALF RGBA, alpha 1). rgb: w*h*3 floats top-down. */
int hxr_save_image(const char* path, const float* rgb, int32_t width, int32_t height);

/* The screenshot straight from device memory (takeScreenshot -> Bitmap::saveBMP, src/sdl.cpp:103-116, src/bitmap.cpp:202-240):
 * the sRGB-table conversion runs on the GPU, only the 8-bit pixel array is copied back. d_rgb: width*height*3 floats on the
 * context's device, or NULL for the frame of the last hxr_render / hxr_render_device call (then width/height are ignored).
 * The file is byte-identical to hxr_save_image(".bmp") of the same frame. */
int hxr_save_frame_bmp(hxr_ctx* ctx, const void* d_rgb, int32_t width, int32_t height, const char* path);
/* the same for Bitmap::saveEXR (src/bitmap.cpp:270-288): float -> half on the GPU; byte-identical to hxr_save_image(".exr") */
int hxr_save_frame_exr(hxr_ctx* ctx, const void* d_rgb, int32_t width, int32_t height, const char* path);

/* Bitmap::loadImage equivalent (".bmp" 8/24/32 bpp, ".exr" scan-line NONE/RLE/ZIPS/ZIP/PIZ): fills *width / *height;
 * when rgb_out is non-NULL and capacity_floats >= width*height*3 also the pixels (float RGB, top-down). */
int hxr_load_image(const char* path, int32_t* width, int32_t* height, float* rgb_out, size_t capacity_floats);

#ifdef __cplusplus
}
#endif
#endif /* HXR_H */
