// tests/emu/launch_emu.cpp — TEST INFRASTRUCTURE ONLY (never linked into libhexray_b200.so).
//
// Host implementation of csrc/device/launch.h: each "kernel" is a loop over the SAME per-item
// functions the CUDA kernels call (csrc/device/pipeline.h), spread over std::threads. It lets
// the CPU test tier (`pytest -m "not gpu"`) exercise the host frame driver, the flattening and
// the per-ray logic against the oracle without a GPU. The product has no CPU path: hxr_create
// in libhexray_b200.so fails with HXR_ERR_NO_DEVICE when CUDA is unavailable.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../../hexray_b200/csrc/device/launch.h"

namespace hxr {
namespace dev {

struct Context {
    int device = 0;
    int threads = 8;
    uint64_t launches[PROF_NCAT] = {};
};

template <class F> static void parallel_for(Context* c, uint32_t n, F f, bool serial = false)
{
    if (n == 0) return;
    const int nt = serial ? 1 : (int)std::min<uint32_t>((uint32_t)c->threads, (n + 63) / 64);
    if (nt <= 1) { for (uint32_t i = 0; i < n; i++) f(i); return; }
    std::atomic<uint32_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&] {
            for (;;) {
                uint32_t b = next.fetch_add(64);
                if (b >= n) break;
                uint32_t e = std::min(n, b + 64);
                for (uint32_t i = b; i < e; i++) f(i);
            }
        });
    for (auto& x : th) x.join();
}

int device_count() { return getenv("HXR_EMU_DEVICES") ? std::max(1, atoi(getenv("HXR_EMU_DEVICES"))) : 2; }
Context* create(int device, char* err, size_t errlen)
{
    if (device < 0 || device >= device_count()) { snprintf(err, errlen, "device ordinal out of range"); return nullptr; }
    Context* c = new Context;
    c->device = device;
    unsigned hc = std::thread::hardware_concurrency();
    c->threads = hc ? (int)hc : 4;
    if (const char* e = getenv("HXR_EMU_THREADS")) c->threads = std::max(1, atoi(e));
    return c;
}
void destroy(Context* c) { delete c; }
int device_of(const Context* c) { return c->device; }
void* stream_of(const Context*) { return nullptr; }
const char* backend_name() { return "host-emulation (tests only)"; }
void* alloc(Context*, size_t bytes) { return calloc(1, bytes ? bytes : 1); }
void free_(Context*, void* p) { free(p); }
void* alloc_pinned(Context*, size_t bytes) { return calloc(1, bytes ? bytes : 1); }
void free_pinned(Context*, void* p) { free(p); }
bool upload(Context*, void* d, const void* s, size_t n) { memcpy(d, s, n); return true; }
bool download(Context*, void* d, const void* s, size_t n) { memcpy(d, s, n); return true; }
bool download_async(Context*, void* d, const void* s, size_t n) { memcpy(d, s, n); return true; }
bool zero(Context*, void* p, size_t n) { memset(p, 0, n); return true; }
bool copy_d2d(Context*, void* d, const void* s, size_t n) { memcpy(d, s, n); return true; }
bool sync(Context*) { return true; }
void lane(Context*, int) {}
void fork(Context*) {}
void join(Context*) {}
const char* last_error(const Context*) { return ""; }
bool failed(const Context*) { return false; }
void clear_error(Context*) {}

struct Timer { std::chrono::steady_clock::time_point a, b; };
Timer* timer_create(Context*) { return new Timer; }
void timer_destroy(Context*, Timer* t) { delete t; }
void timer_start(Context*, Timer* t) { t->a = std::chrono::steady_clock::now(); }
void timer_stop(Context*, Timer* t) { t->b = std::chrono::steady_clock::now(); }
double timer_ms(Context*, Timer* t) { return std::chrono::duration<double, std::milli>(t->b - t->a).count(); }

void prof_enable(Context*, bool) {}
void prof_reset(Context* c) { memset(c->launches, 0, sizeof c->launches); }
void prof_collect(Context* c, double ms[PROF_NCAT], uint64_t launches[PROF_NCAT])
{
    for (int i = 0; i < PROF_NCAT; i++) { ms[i] = 0; launches[i] = c->launches[i]; }
}

int gen_primary(Context* c, const DScene& sc, const FrameParams& fp, const uint32_t* pixels, const uint32_t* pixels_count, uint32_t first_pixel,
                uint32_t n_items, uint32_t spp_pass, const RayQueue& q, CandRec* cand)
{
    RayQueue qq = q;
    if (!sc.n_big) { qq.entry = nullptr; cand = nullptr; }
    if (pixels_count) n_items = std::min(n_items, *pixels_count * spp_pass);
    parallel_for(c, n_items, [&](uint32_t i) {
        const uint32_t pi = i / spp_pass;
        const uint32_t pixel = pixels ? pixels[pi] : first_pixel + pi;
        const uint32_t sample = fp.sample_base + (i % spp_pass) * fp.sample_stride;
        if (sc.simple_inline) gen_primary_item<true>(sc, fp, pixel, sample, qq, cand, i);
        else gen_primary_item<false>(sc, fp, pixel, sample, qq, cand, i);
    });
    *q.count = n_items;
    c->launches[PROF_GEN]++;
    return 1;
}

int setup_closest(Context* c, const DScene& sc, const RayQueue& q, CandRec* cand, TravCounters* cnt, uint32_t)
{
    const uint32_t n = std::min(*q.count, q.cap);
    RayGeom* geom = q.geom;
    MeshEntry* entry = sc.n_big ? q.entry : nullptr;
    parallel_for(c, n, [&](uint32_t i) {
        MeshEntry* e = entry ? entry + i : nullptr;
        CandRec* cr = entry ? cand + i : nullptr;
        if (cnt) setup_closest_geom<true, false>(sc, geom[i], e, cr, cnt);
        else if (sc.simple_inline) setup_closest_geom<false, true>(sc, geom[i], e, cr, nullptr);
        else setup_closest_geom<false, false>(sc, geom[i], e, cr, nullptr);
    }, cnt != nullptr);
    c->launches[PROF_SETUP]++;
    return 1;
}

int setup_shadow(Context* c, const DScene& sc, const ShadowQueue& q, CandRec* cand, float* accum, TravCounters* cnt, uint32_t)
{
    const uint32_t n = std::min(*q.count, q.cap);
    RayGeom* geom = q.geom;
    MeshEntry* entry = sc.n_big ? q.entry : nullptr;
    parallel_for(c, n, [&](uint32_t i) {
        MeshEntry* e = entry ? entry + i : nullptr;
        CandRec* cr = entry ? cand + i : nullptr;
        if (cnt) setup_shadow_geom<true, false>(sc, geom[i], e, cr, q.aux + i, accum, cnt);
        else if (sc.simple_inline) setup_shadow_geom<false, true>(sc, geom[i], e, cr, q.aux + i, accum, nullptr);
        else setup_shadow_geom<false, false>(sc, geom[i], e, cr, q.aux + i, accum, nullptr);
    }, cnt != nullptr);
    c->launches[PROF_SETUP]++;
    return 1;
}

static std::atomic<unsigned long long>* as_atomic(unsigned long long* p) { return reinterpret_cast<std::atomic<unsigned long long>*>(p); }

int walk(Context* c, const DScene& sc, bool shadow, const RayGeom* geom, const MeshEntry* entry, const uint32_t* count, uint32_t cap, const WalkBuffers& wb,
         FrameTotals* totals, TravCounters* cnt, uint32_t)
{
    if (sc.n_big == 0) return 0;
    const uint32_t n = std::min(*count, cap);
    CandRec* cand = wb.cand;
    parallel_for(c, n, [&](uint32_t i) {
        OverflowEntry oe;
        if (entry[i].kup < 0.0f) return;  // dead: its candidate record was written by the setup kernel
        if (shadow) cand[i] = cnt ? walk_ray_item<true, true>(sc, geom[i], entry[i], oe, cnt) : walk_ray_item<true, false>(sc, geom[i], entry[i], oe, nullptr);
        else cand[i] = cnt ? walk_ray_item<false, true>(sc, geom[i], entry[i], oe, cnt) : walk_ray_item<false, false>(sc, geom[i], entry[i], oe, nullptr);
        if (oe.slot >= 0) {
            oe.ray = i;
            oe.pad = 0;
            wb.ovf_list[__atomic_fetch_add(wb.ovf_count, 1u, __ATOMIC_RELAXED)] = oe;
        }
    }, cnt != nullptr);
    const uint32_t m = std::min(*wb.ovf_count, cap);
    parallel_for(c, m, [&](uint32_t k) {
        const OverflowEntry oe = wb.ovf_list[k];
        const uint32_t i = oe.ray;
        if (shadow) cand[i] = cnt ? finish_overflowed_ray<true, true>(sc, geom[i], cand[i], oe, cnt) : finish_overflowed_ray<true, false>(sc, geom[i], cand[i], oe, nullptr);
        else cand[i] = cnt ? finish_overflowed_ray<false, true>(sc, geom[i], cand[i], oe, cnt) : finish_overflowed_ray<false, false>(sc, geom[i], cand[i], oe, nullptr);
    }, cnt != nullptr);
    if (totals && m) as_atomic(&totals->cand_overflow)->fetch_add(m);
    c->launches[shadow ? PROF_WALK_SHADOW : PROF_WALK_CLOSEST]++;
    c->launches[PROF_FINISH]++;
    return 2;
}

int shade(Context* c, const DScene& sc, const FrameParams& fp, const RayQueue& q, const CandRec* cand, uint32_t begin, uint32_t end, const Sinks& sinks,
          FrameTotals* totals, TravCounters* cnt)
{
    const uint32_t e = std::min(end, std::min(*q.count, q.cap));
    if (e > begin) {
        parallel_for(c, e - begin, [&](uint32_t k) {
            const uint32_t i = begin + k;
            EmitCounters ec = {0, 0};
            CandRec cr;
            cr.tri[0] = cr.tri[1] = cr.tri[2] = 0;
            cr.meta = 0;
            if (sc.n_big) cr = cand[i];
            if (fp.gi) {
                if (cnt) shade_item<true, true, false>(sc, fp, q.geom[i], q.aux[i], cr, sinks, ec, cnt);
                else if (sc.simple_inline) shade_item<true, false, true>(sc, fp, q.geom[i], q.aux[i], cr, sinks, ec, nullptr);
                else shade_item<true, false, false>(sc, fp, q.geom[i], q.aux[i], cr, sinks, ec, nullptr);
            } else {
                if (cnt) shade_item<false, true, false>(sc, fp, q.geom[i], q.aux[i], cr, sinks, ec, cnt);
                else if (sc.simple_inline) shade_item<false, false, true>(sc, fp, q.geom[i], q.aux[i], cr, sinks, ec, nullptr);
                else shade_item<false, false, false>(sc, fp, q.geom[i], q.aux[i], cr, sinks, ec, nullptr);
            }
            if (totals) {
                if (ec.shadow_rays) as_atomic(&totals->rays_shadow)->fetch_add(ec.shadow_rays);
                if (ec.cand_overflow) as_atomic(&totals->cand_overflow)->fetch_add(ec.cand_overflow);
            }
        }, cnt != nullptr);
        if (totals) totals->rays_closest += e - begin;
    }
    c->launches[PROF_SHADE]++;
    return 1;
}

int resolve_shadow(Context* c, const DScene& sc, const ShadowQueue& q, const CandRec* cand, float* accum, uint8_t* visible, FrameTotals* totals,
                   TravCounters* cnt, uint32_t)
{
    const uint32_t n = std::min(*q.count, q.cap);
    parallel_for(c, n, [&](uint32_t i) {
        EmitCounters ec = {0, 0};
        CandRec cr;
        cr.tri[0] = cr.tri[1] = cr.tri[2] = 0;
        cr.meta = 0;
        if (sc.n_big) cr = cand[i];
        if (visible) visible[i] = (cnt ? resolve_visible<true>(sc, q.geom + i, cr, ec, cnt) : resolve_visible<false>(sc, q.geom + i, cr, ec, nullptr)) ? 1 : 0;
        else if (cnt) resolve_shadow_item<true>(sc, q.geom + i, q.aux + i, cr, accum, ec, cnt);
        else resolve_shadow_item<false>(sc, q.geom + i, q.aux + i, cr, accum, ec, nullptr);
        if (totals && ec.cand_overflow) as_atomic(&totals->cand_overflow)->fetch_add(ec.cand_overflow);
    }, cnt != nullptr);
    c->launches[PROF_SHADOW_RESOLVE]++;
    return 1;
}

int hit_records(Context* c, const DScene& sc, const RayQueue& q, const CandRec* cand, HitRec* hits, uint32_t)
{
    const uint32_t n = std::min(*q.count, q.cap);
    parallel_for(c, n, [&](uint32_t i) {
        CandRec cr;
        cr.tri[0] = cr.tri[1] = cr.tri[2] = 0;
        cr.meta = 0;
        if (sc.n_big) cr = cand[i];
        hit_record_item<false, false>(sc, q.geom[i], cr, hits[i], nullptr);
    });
    c->launches[PROF_OTHER]++;
    return 1;
}

int aa_detect(Context* c, const float* vfb, int W, int H, int shard_index, int shard_count, uint32_t* list, uint32_t* n_out, uint8_t* mask)
{
    uint32_t n = 0;
    for (int y = 0; y < H; y++) {
        const bool mine = shard_count <= 1 || ((y / HXR_ROW_BAND) % shard_count) == shard_index;
        for (int x = 0; x < W; x++) {
            const bool f = mine && aa_detect_item(vfb, W, H, x, y);
            mask[(size_t)y * W + x] = f;
            if (f) list[n++] = (uint32_t)(y * W + x);
        }
    }
    *n_out = n;
    c->launches[PROF_OTHER]++;
    return 1;
}

int scale_listed(Context* c, float* vfb, const uint32_t* list, const uint32_t* n, uint32_t cap, float mul)
{
    const uint32_t m = std::min(*n, cap);
    for (uint32_t i = 0; i < m; i++)
        for (int k = 0; k < 3; k++) vfb[3 * (size_t)list[i] + k] *= mul;
    c->launches[PROF_OTHER]++;
    return 1;
}
int scale_all(Context* c, float* buf, size_t n, float mul)
{
    for (size_t i = 0; i < n; i++) buf[i] *= mul;
    c->launches[PROF_OTHER]++;
    return 1;
}
int add_into(Context* c, float* dst, const float* src, size_t n)
{
    for (size_t i = 0; i < n; i++) dst[i] += src[i];
    c->launches[PROF_OTHER]++;
    return 1;
}
int to_bmp_rows(Context* c, const float* rgb, int W, int H, int rowsz, const uint8_t* lut, uint8_t* out)
{
    memset(out, 0, (size_t)rowsz * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) bmp_pixel_item(rgb, W, H, rowsz, lut, out, x, y);
    c->launches[PROF_OTHER]++;
    return 1;
}
int to_exr_rows(Context* c, const float* rgb, int W, int H, uint16_t* out)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) exr_pixel_item(rgb, W, out, x, y);
    c->launches[PROF_OTHER]++;
    return 1;
}
int stereo_mix(Context* c, float* out, const float* left, const float* right, size_t n_pixels)
{
    for (size_t i = 0; i < n_pixels; i++) stereo_mix_item(out, left, right, i);
    c->launches[PROF_OTHER]++;
    return 1;
}
bool enable_peer(Context*, Context*) { return true; }
int reduce_peers(Context* c, float* dst, const float* const* srcs, int n_src, size_t n, float scale)
{
    for (size_t i = 0; i < n; i++) {
        float a = dst[i];
        for (int k = 0; k < n_src; k++) a += srcs[k][i];
        dst[i] = a * scale;
    }
    c->launches[PROF_OTHER]++;
    return 1;
}
struct Comm { int n; };
Comm* comm_create(Context* const*, int, char* err, size_t errlen) { snprintf(err, errlen, "no NCCL in the host emulation"); return nullptr; }
void comm_destroy(Comm* c) { delete c; }
bool comm_reduce_sum(Comm*, float* const*, size_t) { return false; }

// ---- device KD-tree build: the passes of kdbuild.h as loops
int kd_bounds(Context*, const double* vertices, const int32_t* triV, uint32_t nTris, double* tb)
{
    for (uint32_t t = 0; t < nTris; t++) kdb::bounds_item(t, vertices, triV, tb, nTris);
    return 1;
}
int kd_iota(Context*, uint32_t* refTri, uint32_t* refNode, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) { refTri[i] = i; refNode[i] = 0; }
    return 1;
}
int kd_bin(Context*, const kdb::Params& P, const kdb::NodeWork* work, const uint32_t* refTri, const uint32_t* refNode, uint32_t nRefs, const double* tb,
           uint32_t nTris, uint32_t* hist)
{
    for (uint32_t i = 0; i < nRefs; i++) {
        const uint32_t node = refNode[i];
        const kdb::NodeWork& w = work[node];
        if ((int)w.count <= P.binnedAbove) continue;
        const uint32_t t = refTri[i];
        for (int a = 0; a < 3; a++) {
            if (!(w.mx[a] - w.mn[a] > 0)) continue;
            int b0, b1;
            kdb::bin_range(tb[(size_t)a * nTris + t], tb[(size_t)(3 + a) * nTris + t], w.mn[a], w.mx[a], b0, b1);
            hist[((size_t)node * 3 + a) * 2 * HXR_KDB_BINS + b0]++;
            hist[((size_t)node * 3 + a) * 2 * HXR_KDB_BINS + HXR_KDB_BINS + b1]++;
        }
    }
    return 1;
}
int kd_choose(Context*, const kdb::Params& P, const kdb::NodeWork* work, uint32_t nNodes, int depth, const uint32_t* hist, const uint32_t* refTri,
              const double* tb, uint32_t nTris, kdb::Decision* dec)
{
    for (uint32_t n = 0; n < nNodes; n++)
        kdb::choose_serial(P, work[n], depth, hist ? hist + (size_t)n * 3 * 2 * HXR_KDB_BINS : nullptr, refTri, tb, nTris, dec[n]);
    return 1;
}
int kd_classify(Context*, const uint32_t* refTri, const uint32_t* refNode, uint32_t nRefs, const kdb::Decision* dec, const double* tb, uint32_t nTris,
                uint32_t* flagL, uint32_t* flagR)
{
    for (uint32_t i = 0; i < nRefs; i++) kdb::classify_item(i, refTri, refNode, dec, tb, nTris, flagL, flagR);
    flagL[nRefs] = flagR[nRefs] = 0;
    return 1;
}
int scan_u32(Context*, uint32_t* data, uint32_t n)
{
    uint32_t run = 0;
    for (uint32_t i = 0; i < n; i++) { const uint32_t v = data[i]; data[i] = run; run += v; }
    return 1;
}
int kd_plan(Context*, const kdb::NodeWork* work, uint32_t nNodes, kdb::Decision* dec, const uint32_t* scanL, const uint32_t* scanR, uint32_t* childRefs,
            uint32_t* isSplit, uint32_t* leafRefs, uint32_t* levelMax)
{
    for (uint32_t n = 0; n < nNodes; n++) {
        kdb::plan_item(n, work, dec, scanL, scanR, childRefs, isSplit, leafRefs);
        if (isSplit[n]) *levelMax = std::max(*levelMax, std::max(dec[n].nl, childRefs[n] - dec[n].nl));
    }
    childRefs[nNodes] = isSplit[nNodes] = leafRefs[nNodes] = 0;
    return 1;
}
int kd_emit(Context*, const kdb::NodeWork* work, uint32_t nNodes, const kdb::Decision* dec, const uint32_t* childRefs, const uint32_t* isSplit,
            const uint32_t* leafRefs, uint32_t outCount, uint32_t leafBase, kdb::OutNode* out, kdb::NodeWork* next)
{
    for (uint32_t n = 0; n < nNodes; n++) kdb::emit_item(n, work, dec, childRefs, isSplit, leafRefs, outCount, leafBase, out, next);
    return 1;
}
int kd_scatter(Context*, const uint32_t* refTri, const uint32_t* refNode, uint32_t nRefs, const kdb::NodeWork* work, const kdb::Decision* dec,
               const uint32_t* scanL, const uint32_t* scanR, const uint32_t* childRefs, const uint32_t* isSplit, const uint32_t* leafRefs,
               uint32_t leafBase, uint32_t* nextTri, uint32_t* nextNode, uint32_t* leafOut)
{
    for (uint32_t i = 0; i < nRefs; i++)
        kdb::scatter_item(i, refTri, refNode, work, dec, scanL, scanR, childRefs, isSplit, leafRefs, leafBase, nextTri, nextNode, leafOut);
    return 1;
}

}  // namespace dev
}  // namespace hxr
