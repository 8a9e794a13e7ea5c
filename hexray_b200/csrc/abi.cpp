// abi.cpp — extern "C" surface of include/hxr.h. No exception leaves this file.
#include <cstring>
#include <memory>
#include <string>
#include "../../include/hxr.h"
#include "host/scene.h"
#include "multi.h"
#include "device/isect.h"

struct hxr_ctx {
    hxr::MultiRenderer m;  // one or several GPUs
    std::string err;
};

struct hxr_scene_file {
    hxr::host::Scene scene;
    hxr::host::FlatScene flat;
    std::string path;
};

static thread_local std::string g_lastError;

#define HXR_GUARD_BEGIN try {
#define HXR_GUARD_END(ctxerr)                                         \
    } catch (const std::bad_alloc&) {                                 \
        ctxerr = "out of host memory";                                \
        return HXR_ERR_INVALID;                                       \
    } catch (const std::exception& e) {                               \
        ctxerr = std::string("internal error: ") + e.what();          \
        return HXR_ERR_INVALID;                                       \
    } catch (...) {                                                   \
        ctxerr = "internal error";                                    \
        return HXR_ERR_INVALID;                                       \
    }

extern "C" {

int hxr_create(const hxr_config* cfg, hxr_ctx** out)
{
    if (!out) { g_lastError = "hxr_create: null output"; return HXR_ERR_INVALID; }
    *out = nullptr;
    HXR_GUARD_BEGIN
    hxr_config c;
    memset(&c, 0, sizeof c);
    if (cfg) c = *cfg;
    std::unique_ptr<hxr_ctx> ctx(new hxr_ctx);
    int rc = ctx->m.create(c);
    if (rc != HXR_OK) { g_lastError = ctx->m.error(); return rc; }
    *out = ctx.release();
    return HXR_OK;
    HXR_GUARD_END(g_lastError)
}

void hxr_destroy(hxr_ctx* ctx)
{
    try { delete ctx; } catch (...) {}
}

const char* hxr_last_error(const hxr_ctx* ctx)
{
    if (!ctx) return g_lastError.c_str();
    return ctx->err.empty() ? ctx->m.error().c_str() : ctx->err.c_str();
}

#define HXR_CTX_CALL(expr)                                   \
    if (!ctx) { g_lastError = "null context"; return HXR_ERR_INVALID; } \
    ctx->err.clear();                                        \
    HXR_GUARD_BEGIN                                          \
    return (expr);                                           \
    HXR_GUARD_END(ctx->err)

int hxr_device_count(void) { return hxr::dev::device_count(); }
const char* hxr_reduce_backend(const hxr_ctx* ctx) { return ctx ? ctx->m.reduceName() : "none"; }

int hxr_upload_scene(hxr_ctx* ctx, const hxr_scene* scene) { HXR_CTX_CALL(ctx->m.uploadScene(scene)) }
int hxr_set_camera(hxr_ctx* ctx, const hxr_camera* cam) { HXR_CTX_CALL(ctx->m.setCamera(cam)) }

int hxr_render(hxr_ctx* ctx, const hxr_render_params* p, float* rgb_out, hxr_stats* stats)
{
    if (ctx && !p) { ctx->err = "null render params"; return HXR_ERR_INVALID; }
    if (ctx && !rgb_out) { ctx->err = "null output buffer"; return HXR_ERR_INVALID; }
    HXR_CTX_CALL(ctx->m.render(*p, rgb_out, nullptr, stats))
}
int hxr_render_device(hxr_ctx* ctx, const hxr_render_params* p, void* d_rgb, hxr_stats* stats)
{
    if (ctx && !p) { ctx->err = "null render params"; return HXR_ERR_INVALID; }
    if (ctx && !d_rgb) { ctx->err = "null output buffer"; return HXR_ERR_INVALID; }
    HXR_CTX_CALL(ctx->m.render(*p, nullptr, d_rgb, stats))
}
int hxr_resolve_device(hxr_ctx* ctx, void* d_rgb, int32_t w, int32_t h, int32_t spp) { HXR_CTX_CALL(ctx->m.primary().resolveDevice(d_rgb, w, h, spp)) }
int hxr_progressive_begin(hxr_ctx* ctx, const hxr_render_params* p, int32_t n_passes)
{
    if (ctx && !p) { ctx->err = "null render params"; return HXR_ERR_INVALID; }
    HXR_CTX_CALL(ctx->m.progressiveBegin(*p, n_passes))
}
int hxr_progressive_pass(hxr_ctx* ctx, float* rgb_out, hxr_stats* stats) { HXR_CTX_CALL(ctx->m.progressivePass(rgb_out, stats)) }
int hxr_progressive_state(hxr_ctx* ctx, float* sum_out, int32_t* passes_done, int32_t* spp_done) { HXR_CTX_CALL(ctx->m.progressiveState(sum_out, passes_done, spp_done)) }
int hxr_progressive_resume(hxr_ctx* ctx, const hxr_render_params* p, int32_t n_passes, const float* sum, int32_t passes_done, int32_t spp_done)
{
    if (ctx && !p) { ctx->err = "null render params"; return HXR_ERR_INVALID; }
    HXR_CTX_CALL(ctx->m.progressiveResume(*p, n_passes, sum, passes_done, spp_done))
}
int hxr_set_profiling(hxr_ctx* ctx, int32_t on) { HXR_CTX_CALL((ctx->m.setProfiling(on != 0), HXR_OK)) }
int hxr_trace_closest(hxr_ctx* ctx, const hxr_ray* rays, size_t n, hxr_hit* hits) { HXR_CTX_CALL(ctx->m.primary().traceClosest(rays, n, hits)) }
int hxr_trace_visible(hxr_ctx* ctx, const double* seg, size_t n, uint8_t* vis) { HXR_CTX_CALL(ctx->m.primary().traceVisible(seg, n, vis)) }
int hxr_trace_color(hxr_ctx* ctx, const hxr_ray* rays, size_t n, float* rgb) { HXR_CTX_CALL(ctx->m.primary().traceColor(rays, n, rgb)) }
int hxr_get_accel_info(hxr_ctx* ctx, int32_t mesh, hxr_accel_info* out) { HXR_CTX_CALL(ctx->m.primary().accelInfo(mesh, out)) }
int hxr_save_frame_bmp(hxr_ctx* ctx, const void* d_rgb, int32_t w, int32_t h, const char* path) { HXR_CTX_CALL(ctx->m.primary().saveFrameBmp(d_rgb, w, h, path)) }
int hxr_save_frame_exr(hxr_ctx* ctx, const void* d_rgb, int32_t w, int32_t h, const char* path) { HXR_CTX_CALL(ctx->m.primary().saveFrameExr(d_rgb, w, h, path)) }

static int test_tri_filter(bool packed, size_t n, const double* rays, const double* tris, const double* tbest, int32_t backface, int32_t* cls_out,
                           float* ghi_out, int32_t* exact_out, double* gamma_out)
{
    if (n && (!rays || !tris || !tbest || !cls_out || !ghi_out || !exact_out || !gamma_out)) { g_lastError = "hxr_test_tri_filter: null argument"; return HXR_ERR_INVALID; }
    using namespace hxr;
    for (size_t i = 0; i < n; i++) {
        const double* r = rays + 6 * i;
        const double* v = tris + 9 * i;
        TriTest tt;
        TriF32 tf;
        double amax = 0;
        for (int k = 0; k < 3; k++) {
            tt.A[k] = v[k];
            tt.AB[k] = v[3 + k] - v[k];
            tt.AC[k] = v[6 + k] - v[k];
            for (int c = 0; c < 3; c++) amax = std::max(amax, std::fabs(v[3 * c + k]));
        }
        const d3 N = cross(ld3(tt.AB), ld3(tt.AC));
        tt.N[0] = N.x; tt.N[1] = N.y; tt.N[2] = N.z;
        for (int k = 0; k < 3; k++) { tf.A[k] = (float)tt.A[k]; tf.AB[k] = (float)tt.AB[k]; tf.AC[k] = (float)tt.AC[k]; tf.N[k] = (float)tt.N[k]; }
        // what enter_mesh hands the walk (pipeline.h): float ray, error bound from |o| and the mesh extent
        const float mo = std::max(std::fabs((float)r[0]), std::max(std::fabs((float)r[1]), std::fabs((float)r[2])));
        const float err = (mo * (1.0f + 1e-6f) + std::nextafter((float)amax, INFINITY)) * (1.02f / 8388608.0f);
        float ghi = 0;
        TriPacked tp;
        if (packed && !pack_tri(tt, tp)) { g_lastError = "hxr_test_tri_filter_packed: triangle does not fit the packed form"; return HXR_ERR_INVALID; }
        cls_out[i] = packed ? tri_filter_packed(&tp, backface != 0, (float)r[0], (float)r[1], (float)r[2], (float)r[3], (float)r[4], (float)r[5], err,
                                                f32_above(tbest[i]), ghi)
                            : tri_filter(&tf, backface != 0, (float)r[0], (float)r[1], (float)r[2], (float)r[3], (float)r[4], (float)r[5], err,
                                         f32_above(tbest[i]), ghi);
        ghi_out[i] = ghi;
        Ray ray;
        ray.o = ld3(r); ray.d = ld3(r + 3); ray.depth = 0; ray.flags = 0;
        double g = 0, l2, l3;
        exact_out[i] = tri_core(&tt, backface != 0, ray, 0, tbest[i], -1, g, l2, l3) ? 1 : 0;
        gamma_out[i] = g;
    }
    return HXR_OK;
}

int hxr_test_tri_filter(size_t n, const double* rays, const double* tris, const double* tbest, int32_t backface, int32_t* cls_out, float* ghi_out,
                        int32_t* exact_out, double* gamma_out)
{
    return test_tri_filter(false, n, rays, tris, tbest, backface, cls_out, ghi_out, exact_out, gamma_out);
}
int hxr_test_tri_filter_packed(size_t n, const double* rays, const double* tris, const double* tbest, int32_t backface, int32_t* cls_out,
                               float* ghi_out, int32_t* exact_out, double* gamma_out)
{
    return test_tri_filter(true, n, rays, tris, tbest, backface, cls_out, ghi_out, exact_out, gamma_out);
}

// ---------------------------------------------------------------- host front-end
int hxr_scene_load(const char* path, hxr_scene_file** out)
{
    if (!path || !out) { g_lastError = "hxr_scene_load: null argument"; return HXR_ERR_INVALID; }
    *out = nullptr;
    HXR_GUARD_BEGIN
    std::unique_ptr<hxr_scene_file> sf(new hxr_scene_file);
    sf->path = path;
    if (!sf->scene.parseScene(path)) {
        g_lastError = sf->scene.lastError.empty() ? std::string("could not parse ") + path : sf->scene.lastError;
        return HXR_ERR_PARSE;
    }
    std::string err;
    if (!hxr::host::flattenScene(sf->scene, sf->flat, err)) { g_lastError = err; return HXR_ERR_PARSE; }
    *out = sf.release();
    return HXR_OK;
    HXR_GUARD_END(g_lastError)
}

const hxr_scene* hxr_scene_file_scene(const hxr_scene_file* sf) { return sf ? &sf->flat.pod : nullptr; }

int hxr_scene_file_camera(const hxr_scene_file* sf, hxr_camera* out)
{
    if (!sf || !out || !sf->scene.camera) { g_lastError = "hxr_scene_file_camera: null argument"; return HXR_ERR_INVALID; }
    sf->scene.camera->computeFrame(*out);
    return HXR_OK;
}

int hxr_scene_file_set_synthetic_mesh(hxr_scene_file* sf, int32_t mesh_index, const char* kind, int64_t n, uint64_t seed)
{
    if (!sf || !kind) { g_lastError = "null argument"; return HXR_ERR_INVALID; }
    HXR_GUARD_BEGIN
    using namespace hxr::host;
    // mesh_index counts Mesh geometries in scene order
    Mesh* target = nullptr;
    int k = 0;
    for (Geometry* g : sf->scene.geometries)
        if (Mesh* m = dynamic_cast<Mesh*>(g)) {
            if (k == mesh_index) target = m;
            k++;
        }
    if (!target) { g_lastError = "no such mesh"; return HXR_ERR_INVALID; }
    if (!strcmp(kind, "terrain")) target->generateTerrain((int)n, seed);
    else if (!strcmp(kind, "soup")) target->generateSoup(n, seed);
    else { g_lastError = "unknown synthetic mesh kind"; return HXR_ERR_INVALID; }
    target->recenter = false;
    target->autoSmooth = false;
    // only the mesh changed: re-flatten that one table entry in place (re-running beginRender on the
    // whole scene would differentiate BumpTexture images a second time)
    target->computeBoundingGeometry();
    int mi = 0;
    for (size_t gi = 0; gi < sf->scene.geometries.size(); gi++) {
        if (Mesh* m = dynamic_cast<Mesh*>(sf->scene.geometries[gi])) {
            if (m == target) {
                hxr_geometry g;
                memset(&g, 0, sizeof g);
                FlatScene tmp;
                m->flatten(tmp, g);
                // move the storage into the live flat scene and patch the mesh record
                for (auto& d : tmp.doubleStore) sf->flat.doubleStore.push_back(std::move(d));
                hxr_mesh rec = tmp.meshes[0];
                const size_t nd = sf->flat.doubleStore.size();
                rec.vertices = sf->flat.doubleStore[nd - 3].data();
                rec.normals = sf->flat.doubleStore[nd - 2].data();
                rec.uvs = sf->flat.doubleStore[nd - 1].data();
                sf->flat.meshes[mi] = rec;
            }
            mi++;
        }
    }
    sf->flat.finalize();
    return HXR_OK;
    HXR_GUARD_END(g_lastError)
}

int hxr_scene_file_write_obj(const hxr_scene_file* sf, int32_t mesh_index, const char* path)
{
    if (!sf || !path) { g_lastError = "null argument"; return HXR_ERR_INVALID; }
    HXR_GUARD_BEGIN
    int k = 0;
    for (hxr::host::Geometry* g : sf->scene.geometries)
        if (auto* m = dynamic_cast<hxr::host::Mesh*>(g)) {
            if (k == mesh_index) {
                if (!m->saveOBJ(path)) { g_lastError = std::string("cannot write ") + path; return HXR_ERR_IO; }
                return HXR_OK;
            }
            k++;
        }
    g_lastError = "no such mesh";
    return HXR_ERR_INVALID;
    HXR_GUARD_END(g_lastError)
}

void hxr_scene_file_free(hxr_scene_file* sf)
{
    try { delete sf; } catch (...) {}
}

int hxr_save_image(const char* path, const float* rgb, int32_t width, int32_t height)
{
    if (!path || !rgb || width <= 0 || height <= 0) { g_lastError = "hxr_save_image: bad argument"; return HXR_ERR_INVALID; }
    HXR_GUARD_BEGIN
    hxr::host::Bitmap bmp;
    bmp.generateEmptyImage(width, height);
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            const float* p = rgb + 3 * ((size_t)y * width + x);
            bmp.setPixel(x, y, hxr::host::Color3(p[0], p[1], p[2]));
        }
    if (!bmp.saveImage(path)) { g_lastError = std::string("cannot write ") + path; return HXR_ERR_IO; }
    return HXR_OK;
    HXR_GUARD_END(g_lastError)
}

int hxr_load_image(const char* path, int32_t* width, int32_t* height, float* rgb_out, size_t capacity_floats)
{
    if (!path || !width || !height) { g_lastError = "hxr_load_image: bad argument"; return HXR_ERR_INVALID; }
    HXR_GUARD_BEGIN
    hxr::host::Bitmap bmp;
    if (!bmp.loadImage(path) || !bmp.isOK()) { g_lastError = std::string("cannot load ") + path; return HXR_ERR_IO; }
    *width = bmp.getWidth();
    *height = bmp.getHeight();
    const size_t n = (size_t)bmp.getWidth() * bmp.getHeight();
    if (rgb_out) {
        if (capacity_floats < n * 3) { g_lastError = "hxr_load_image: buffer too small"; return HXR_ERR_INVALID; }
        const auto& px = bmp.data();
        for (size_t i = 0; i < n; i++) { rgb_out[i * 3] = px[i].r; rgb_out[i * 3 + 1] = px[i].g; rgb_out[i * 3 + 2] = px[i].b; }
    }
    return HXR_OK;
    HXR_GUARD_END(g_lastError)
}

}  // extern "C"
