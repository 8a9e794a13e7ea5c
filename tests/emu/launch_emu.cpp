// tests/emu/launch_emu.cpp — TEST INFRASTRUCTURE ONLY (never linked into libhexray_b200.so).
//
// Host implementation of csrc/device/launch.h: each "kernel" is a loop over the SAME per-item
// functions the CUDA kernels call (csrc/device/pipeline.h), spread over std::threads. It lets
// the CPU test tier (`pytest -m "not gpu"`) exercise the host frame driver, the flattening and
// the per-ray logic against the oracle without a GPU. The product has no CPU path: hxr_create
// in libhexray_b200.so fails with HXR_ERR_NO_DEVICE when CUDA is unavailable.
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../../hexray_b200/csrc/device/launch.h"

namespace hxr {
namespace dev {

static int g_threads = 8;
static uint64_t g_launches[PROF_NCAT];

template <class F> static void parallel_for(uint32_t n, F f)
{
    if (n == 0) return;
    const int nt = (int)std::min<uint32_t>((uint32_t)g_threads, (n + 63) / 64);
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

bool init(int, char*, size_t)
{
    unsigned hc = std::thread::hardware_concurrency();
    g_threads = hc ? (int)hc : 4;
    if (const char* e = getenv("HXR_EMU_THREADS")) g_threads = std::max(1, atoi(e));
    return true;
}
const char* backend_name() { return "host-emulation (tests only)"; }
void* alloc(size_t bytes) { return calloc(1, bytes ? bytes : 1); }
void free_(void* p) { free(p); }
bool upload(void* d, const void* s, size_t n) { memcpy(d, s, n); return true; }
bool download(void* d, const void* s, size_t n) { memcpy(d, s, n); return true; }
bool upload_pinned_async(void* d, const void* s, size_t n) { memcpy(d, s, n); return true; }
bool zero(void* p, size_t n) { memset(p, 0, n); return true; }
bool copy_d2d(void* d, const void* s, size_t n) { memcpy(d, s, n); return true; }
bool sync() { return true; }
const char* last_error() { return ""; }
bool set_u32(uint32_t* p, uint32_t v) { *p = v; return true; }

struct Timer { std::chrono::steady_clock::time_point a, b; };
Timer* timer_create() { return new Timer; }
void timer_destroy(Timer* t) { delete t; }
void timer_start(Timer* t) { t->a = std::chrono::steady_clock::now(); }
void timer_stop(Timer* t) { t->b = std::chrono::steady_clock::now(); }
double timer_ms(Timer* t) { return std::chrono::duration<double, std::milli>(t->b - t->a).count(); }

void prof_enable(bool) {}
void prof_reset() { memset(g_launches, 0, sizeof g_launches); }
void prof_collect(double ms[PROF_NCAT], uint64_t launches[PROF_NCAT])
{
    for (int i = 0; i < PROF_NCAT; i++) { ms[i] = 0; launches[i] = g_launches[i]; }
}

int gen_primary(const DScene& sc, const FrameParams& fp, const uint32_t* pixels, uint32_t first_pixel, uint32_t n_items,
                uint32_t spp_pass, RayTask* q, uint32_t* q_count)
{
    parallel_for(n_items, [&](uint32_t i) {
        const uint32_t pi = i / spp_pass;
        const uint32_t pixel = pixels ? pixels[pi] : first_pixel + pi;
        q[i] = gen_primary_item(sc, fp, pixel, fp.sample_base + (i % spp_pass) * fp.sample_stride);
    });
    *q_count = n_items;
    g_launches[PROF_OTHER]++;
    return 1;
}

int trace_closest(const DScene& sc, const RayTask* q, const uint32_t* q_count, uint32_t, HitRec* hits, const TraceScratch& ts, TravCounters* cnt, uint32_t)
{
    const uint32_t n = *q_count;
    *ts.task_count = 0;
    *ts.pair_count = 0;
    auto stage = [&](uint32_t m, auto f) {
        if (cnt) { for (uint32_t i = 0; i < m; i++) f(i); } else parallel_for(m, f);
    };
    const bool simple = sc.simple_inline && !cnt;
    if (cnt) stage(n, [&](uint32_t i) { setup_closest_item<true, false>(sc, task_ray(q[i]), i, ts, cnt); });
    else if (simple) stage(n, [&](uint32_t i) { setup_closest_item<false, true>(sc, task_ray(q[i]), i, ts, nullptr); });
    else stage(n, [&](uint32_t i) { setup_closest_item<false, false>(sc, task_ray(q[i]), i, ts, nullptr); });
    const uint32_t nt = std::min(*ts.task_count, ts.task_cap);
    if (cnt) stage(nt, [&](uint32_t k) { walk_item<false, true>(sc, k, ts, cnt); });
    else stage(nt, [&](uint32_t k) { walk_item<false, false>(sc, k, ts, nullptr); });
    const uint32_t np = std::min(*ts.pair_count, ts.pair_cap);
    stage(np, [&](uint32_t i) { confirm_closest_a_item(sc, q, i, ts); });
    stage(np, [&](uint32_t i) { confirm_closest_b_item(sc, i, ts); });
    if (cnt) stage(n, [&](uint32_t i) { finalize_closest_item<true, false>(sc, task_ray(q[i]), i, ts, hits[i], cnt); });
    else if (simple) stage(n, [&](uint32_t i) { finalize_closest_item<false, true>(sc, task_ray(q[i]), i, ts, hits[i], nullptr); });
    else stage(n, [&](uint32_t i) { finalize_closest_item<false, false>(sc, task_ray(q[i]), i, ts, hits[i], nullptr); });
    g_launches[PROF_TRACE_CLOSEST] += 5;
    return 5;
}

int shade(const DScene& sc, const FrameParams& fp, const RayTask* q, const uint32_t* q_count, const HitRec* hits, uint32_t begin,
          uint32_t end, const Sinks& sinks)
{
    const uint32_t e = std::min(end, *q_count);
    if (e > begin)
        parallel_for(e - begin, [&](uint32_t k) {
            const uint32_t i = begin + k;
            if (fp.gi) shade_gi_item(sc, fp, q[i], hits[i], sinks);
            else shade_whitted_item(sc, fp, q[i], hits[i], sinks);
        });
    g_launches[PROF_SHADE]++;
    return 1;
}

int trace_shadow(const DScene& sc, const ShadowTask* shadow, const uint32_t* count, uint32_t cap, float* accum, const TraceScratch& ts,
                 TravCounters* cnt, unsigned long long* total, uint32_t)
{
    const uint32_t n = std::min(*count, cap);
    *ts.task_count = 0;
    *ts.pair_count = 0;
    auto stage = [&](uint32_t m, auto f) {
        if (cnt) { for (uint32_t i = 0; i < m; i++) f(i); } else parallel_for(m, f);
    };
    if (cnt) stage(n, [&](uint32_t i) { setup_shadow_item<true, false>(sc, shadow[i], i, ts, cnt); });
    else if (sc.simple_inline) stage(n, [&](uint32_t i) { setup_shadow_item<false, true>(sc, shadow[i], i, ts, nullptr); });
    else stage(n, [&](uint32_t i) { setup_shadow_item<false, false>(sc, shadow[i], i, ts, nullptr); });
    const uint32_t nt = std::min(*ts.task_count, ts.task_cap);
    if (cnt) stage(nt, [&](uint32_t k) { walk_item<true, true>(sc, k, ts, cnt); });
    else stage(nt, [&](uint32_t k) { walk_item<true, false>(sc, k, ts, nullptr); });
    const uint32_t np = std::min(*ts.pair_count, ts.pair_cap);
    stage(np, [&](uint32_t i) { confirm_shadow_item(sc, shadow, i, ts); });
    if (accum) stage(n, [&](uint32_t i) { accumulate_shadow_item(shadow[i], i, ts, accum); });
    if (total) *total += n;
    g_launches[PROF_TRACE_SHADOW] += 4;
    return 4;
}

int aa_detect(const float* vfb, int W, int H, int shard_index, int shard_count, uint32_t* list, uint32_t* n_out, uint8_t* mask)
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
    g_launches[PROF_OTHER]++;
    return 1;
}

int scale_listed(float* vfb, const uint32_t* list, const uint32_t* n, uint32_t cap, float mul)
{
    const uint32_t m = std::min(*n, cap);
    for (uint32_t i = 0; i < m; i++)
        for (int c = 0; c < 3; c++) vfb[3 * (size_t)list[i] + c] *= mul;
    g_launches[PROF_OTHER]++;
    return 1;
}
int scale_all(float* buf, size_t n, float mul)
{
    for (size_t i = 0; i < n; i++) buf[i] *= mul;
    g_launches[PROF_OTHER]++;
    return 1;
}
int add_into(float* dst, const float* src, size_t n)
{
    for (size_t i = 0; i < n; i++) dst[i] += src[i];
    g_launches[PROF_OTHER]++;
    return 1;
}
int to_bmp_rows(const float* rgb, int W, int H, int rowsz, const uint8_t* lut, uint8_t* out)
{
    memset(out, 0, (size_t)rowsz * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) bmp_pixel_item(rgb, W, H, rowsz, lut, out, x, y);
    g_launches[PROF_OTHER]++;
    return 1;
}
int stereo_mix(float* out, const float* left, const float* right, size_t n_pixels)
{
    for (size_t i = 0; i < n_pixels; i++) stereo_mix_item(out, left, right, i);
    g_launches[PROF_OTHER]++;
    return 1;
}
}  // namespace dev
}  // namespace hxr
