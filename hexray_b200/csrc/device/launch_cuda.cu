// launch_cuda.cu — the sm_100a kernels of the render hot path and their launchers
// (implements device/launch.h; the only translation unit compiled by nvcc).
//
// One bounce of the wavefront is SIX launches + a small one after each walk; a ray lives in a 64-byte geometry record + a 24-byte
// payload between them:
//   k_setup<closest>     the inline part of raycast(), in place on the queue: depth guard + every inline node (analytic
//                        primitives, CSG, heightfields, quads) in scene order, in double; reads 64 B, writes back 16 B
//   k_walk<closest>      the KD-tree walk, FP32 only: PERSISTENT warps; a lane owns one RAY and walks, one after the other,
//                        every big mesh whose box the ray enters (the world -> object transform of the ray is the only double
//                        arithmetic here, done once per (ray, mesh) at refill); idle lanes are refilled with __ballot_sync +
//                        one atomicAdd per warp + __shfl_sync; a bounded phase of block steps (32-byte block = two tree
//                        levels) alternates with a warp-cooperative phase that filters the triangles of all leaves held
//                        by the warp, 32 (ray, triangle) pairs at a time; pairs the filter cannot rule out are collected
//                        in the owner's shared-memory slots and leave the kernel as ONE 16-byte candidate record per ray
//   k_finish_warp        the few rays whose candidate record filled up (a 4th candidate): the walk resumed where it stopped, ONE WARP
//                        per ray, exact double tests on the spot (a leaf's triangles split over the lanes, shuffle argmin)
//   k_shade<GI>          per ray: exact (double) test of its candidates, winner across inline nodes and meshes,
//                        IntersectionInfo, lights, environment, bump, then the Whitted shader tree or the path-tracing vertex;
//                        pushes child rays and shadow rays; radiance lands with RED.ADD.F32
//   k_setup<shadow>      the inline part of visible(): inline nodes and lights, marks blocked rays
//   k_walk<shadow>       any-hit form of the same walk
//   k_resolve_shadow     exact test of the undecided shadow pairs, then the carried colour
// (Round 2 first fused the two setup kernels into k_shade: the result ran at 7.7 of 32 lanes per instruction and stalled on
// instruction fetch - profiles/ncu_fused_shade_r2_v1.txt - so the inline part is its own lean, convergent kernel again.)
// plus k_gen_primary (camera rays with their inline part), k_hit_records (test hook) and the frame-buffer passes
// (k_aa_detect, k_scale_*, k_add_into, k_stereo_mix, k_to_bmp_rows, k_to_exr_rows).
// Grid sizing: the persistent walk launches (SM count x resident blocks/SM) blocks - 148 SMs on B200 - the per-item
// kernels are grid-stride loops over device-side counts, sized by a host-side upper bound of the count: the host never
// reads a count back between bounces.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <dlfcn.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "launch.h"

#ifndef HXR_FINISH_WARP
#define HXR_FINISH_WARP 1 /* overflowed rays are finished by one WARP each (k_finish_warp); 0: one thread each (k_finish). HXR_FINISH_WARP in the environment overrides */
#endif

namespace hxr {
namespace dev {

struct Context {
    int device = -1;
    int sms = 0;
    cudaStream_t stream = nullptr;  // the CURRENT lane's stream (lane()): what every launch below uses
    cudaStream_t lanes[2] = {nullptr, nullptr};
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    std::string err;  // first failure since clear_error
    bool prof = false;
    // walk-loop tunables (uniform kernel arguments; environment overrides for A/B runs: HXR_WALK_STEPS, HXR_REFILL_MIN, HXR_SSTACK,
    // HXR_NO_MAILBOX, HXR_BRANCHY_PUSH, HXR_WALK_CARVEOUT, HXR_WALK_BLOCKS_PER_SM; measured optima are the defaults)
    int walkSteps = 4, refillMin = 8, sstack = 10, useMail = 1, bfPush = 1, walkCarveout = -1, walkBlocksPerSm = 0;
    int finishWarp = HXR_FINISH_WARP;  // overflowed rays: one warp per ray (k_finish_warp) or one thread per ray (k_finish)
    uint64_t launches[PROF_NCAT] = {};
    std::vector<cudaEvent_t> evPool;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evPairs[PROF_NCAT];
    int walkGrid[2][2][3][2] = {};  // cached occupancy-derived grid per k_walk instantiation
    void* scanTmp = nullptr;  // scratch of the prefix sums of the device KD build
    size_t scanTmpBytes = 0;
};

static bool ck(Context* c, cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return true;
    if (c->err.empty()) c->err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
static bool use(Context* c) { return c && c->device >= 0 && ck(c, cudaSetDevice(c->device), "cudaSetDevice"); }

int device_count()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

Context* create(int device, char* err, size_t errlen)
{
    auto fail = [&](const std::string& m) -> Context* {
        snprintf(err, errlen, "%s", m.c_str());
        return nullptr;
    };
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                    "); hexray_b200 has no CPU fallback");
    if (device < 0 || device >= n) return fail("CUDA device ordinal out of range");
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed");
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return fail("cudaGetDeviceProperties failed");
    if (p.major < 10) return fail(std::string("device '") + p.name + "' is not Blackwell (sm_100a code only)");
    Context* c = new Context;
    c->device = device;
    c->sms = p.multiProcessorCount;
    if (const char* v = getenv("HXR_WALK_STEPS")) c->walkSteps = std::max(1, atoi(v));
    if (const char* v = getenv("HXR_SSTACK")) c->sstack = atoi(v);
    if (getenv("HXR_NO_MAILBOX")) c->useMail = 0;
    if (getenv("HXR_BRANCHY_PUSH")) c->bfPush = 0;
    if (const char* v = getenv("HXR_WALK_CARVEOUT")) c->walkCarveout = std::min(100, std::max(0, atoi(v)));
    if (const char* v = getenv("HXR_WALK_BLOCKS_PER_SM")) c->walkBlocksPerSm = std::max(1, atoi(v));
    if (const char* v = getenv("HXR_REFILL_MIN")) c->refillMin = std::min(32, std::max(1, atoi(v)));
    if (const char* v = getenv("HXR_FINISH_WARP")) c->finishWarp = atoi(v) != 0;
    if (cudaStreamCreateWithFlags(&c->lanes[0], cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&c->lanes[1], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming) != cudaSuccess) {
        delete c;
        return fail("cudaStreamCreate failed");
    }
    c->stream = c->lanes[0];
    return c;
}
void destroy(Context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->lanes[0]);
    cudaStreamSynchronize(c->lanes[1]);
    for (int k = 0; k < PROF_NCAT; k++)
        for (auto& pr : c->evPairs[k]) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (cudaEvent_t e : c->evPool) cudaEventDestroy(e);
    if (c->scanTmp) cudaFree(c->scanTmp);
    cudaEventDestroy(c->evFork);
    cudaEventDestroy(c->evJoin);
    cudaStreamDestroy(c->lanes[0]);
    cudaStreamDestroy(c->lanes[1]);
    delete c;
}
int device_of(const Context* c) { return c->device; }
void* stream_of(const Context* c) { return (void*)c->lanes[0]; }
const char* backend_name() { return "cuda sm_100a"; }
const char* last_error(const Context* c) { return c->err.c_str(); }
bool failed(const Context* c) { return !c->err.empty(); }
void clear_error(Context* c) { c->err.clear(); }

void* alloc(Context* c, size_t bytes)
{
    if (!use(c)) return nullptr;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();  // an allocation failure is reported to the caller, it does not poison the context
        return nullptr;
    }
    return p;
}
void free_(Context* c, void* p)
{
    if (p && use(c)) cudaFree(p);
}
void* alloc_pinned(Context* c, size_t bytes)
{
    if (!use(c)) return nullptr;
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void free_pinned(Context* c, void* p)
{
    if (p && use(c)) cudaFreeHost(p);
}
bool upload(Context* c, void* d, const void* s, size_t n)
{
    return use(c) && ck(c, cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, c->stream), "H2D copy") && ck(c, cudaStreamSynchronize(c->stream), "H2D sync");
}
bool download(Context* c, void* d, const void* s, size_t n)
{
    return use(c) && ck(c, cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, c->stream), "D2H copy") && ck(c, cudaStreamSynchronize(c->stream), "D2H sync");
}
bool download_async(Context* c, void* d, const void* s, size_t n) { return use(c) && ck(c, cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, c->stream), "D2H copy"); }
bool zero(Context* c, void* p, size_t n) { return use(c) && ck(c, cudaMemsetAsync(p, 0, n, c->stream), "memset"); }
bool copy_d2d(Context* c, void* d, const void* s, size_t n) { return use(c) && ck(c, cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, c->stream), "D2D copy"); }
bool sync(Context* c) { return use(c) && ck(c, cudaStreamSynchronize(c->lanes[1]), "stream sync") && ck(c, cudaStreamSynchronize(c->lanes[0]), "stream sync"); }
void lane(Context* c, int which) { c->stream = c->lanes[which ? 1 : 0]; }
void fork(Context* c)
{
    use(c);
    ck(c, cudaEventRecord(c->evFork, c->lanes[0]), "fork");
    ck(c, cudaStreamWaitEvent(c->lanes[1], c->evFork, 0), "fork");
}
void join(Context* c)
{
    use(c);
    ck(c, cudaEventRecord(c->evJoin, c->lanes[1]), "join");
    ck(c, cudaStreamWaitEvent(c->lanes[0], c->evJoin, 0), "join");
}

struct Timer { cudaEvent_t a, b; };
Timer* timer_create(Context* c)
{
    use(c);
    Timer* t = new Timer;
    cudaEventCreate(&t->a);
    cudaEventCreate(&t->b);
    return t;
}
void timer_destroy(Context* c, Timer* t)
{
    if (!t) return;
    use(c);
    cudaEventDestroy(t->a);
    cudaEventDestroy(t->b);
    delete t;
}
void timer_start(Context* c, Timer* t) { use(c); cudaEventRecord(t->a, c->stream); }
void timer_stop(Context* c, Timer* t) { use(c); cudaEventRecord(t->b, c->stream); }
double timer_ms(Context* c, Timer* t)
{
    float ms = 0;
    use(c);
    cudaEventSynchronize(t->b);
    cudaEventElapsedTime(&ms, t->a, t->b);
    return ms;
}

// ---- per-launch profiling -------------------------------------------------------------
void prof_enable(Context* c, bool on) { c->prof = on; }
static cudaEvent_t ev_get(Context* c)
{
    if (!c->evPool.empty()) { cudaEvent_t e = c->evPool.back(); c->evPool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_reset(Context* c)
{
    for (int k = 0; k < PROF_NCAT; k++) {
        c->launches[k] = 0;
        for (auto& pr : c->evPairs[k]) { c->evPool.push_back(pr.first); c->evPool.push_back(pr.second); }
        c->evPairs[k].clear();
    }
}
void prof_collect(Context* c, double ms[PROF_NCAT], uint64_t launches[PROF_NCAT])
{
    use(c);
    cudaStreamSynchronize(c->lanes[0]);
    cudaStreamSynchronize(c->lanes[1]);
    for (int k = 0; k < PROF_NCAT; k++) {
        double s = 0;
        for (auto& pr : c->evPairs[k]) {
            float m = 0;
            cudaEventElapsedTime(&m, pr.first, pr.second);
            s += m;
        }
        ms[k] = s;
        launches[k] = c->launches[k];
    }
}
// brackets one launch: counts it, times it when profiling is on, and records a launch failure in the context (sticky)
struct LaunchScope {
    Context* c;
    int cat;
    cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(Context* ctx, int category) : c(ctx), cat(category)
    {
        use(c);
        c->launches[cat]++;
        if (c->prof) { a = ev_get(c); b = ev_get(c); cudaEventRecord(a, c->stream); }
    }
    ~LaunchScope()
    {
        if (c->prof) { cudaEventRecord(b, c->stream); c->evPairs[cat].emplace_back(a, b); }
        ck(c, cudaGetLastError(), "kernel launch");
    }
};

// ---- small device helpers ------------------------------------------------------------------
__device__ __forceinline__ RayGeom load_geom(const RayGeom* p)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    RayGeom g;
    g.o[0] = a.x; g.o[1] = a.y; g.o[2] = b.x;
    g.d[0] = b.y; g.d[1] = c.x; g.d[2] = c.y;
    g.limit = d.x;
    const long long t = __double_as_longlong(d.y);
    g.pre = (int32_t)(uint32_t)(t & 0xFFFFFFFFll);
    g.depth_flags = (uint32_t)((unsigned long long)t >> 32);
    return g;
}
// same with plain (coherent) loads: for the kernel that updates the record in place
__device__ __forceinline__ RayGeom load_geom_rw(const RayGeom* p)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = q[0], b = q[1], c = q[2], d = q[3];
    RayGeom g;
    g.o[0] = a.x; g.o[1] = a.y; g.o[2] = b.x;
    g.d[0] = b.y; g.d[1] = c.x; g.d[2] = c.y;
    g.limit = d.x;
    const long long t = __double_as_longlong(d.y);
    g.pre = (int32_t)(uint32_t)(t & 0xFFFFFFFFll);
    g.depth_flags = (uint32_t)((unsigned long long)t >> 32);
    return g;
}
__device__ __forceinline__ RayAux load_aux(const RayAux* p)
{
    const uint2* q = reinterpret_cast<const uint2*>(p);
    const uint2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    RayAux x;
    x.w[0] = __uint_as_float(a.x); x.w[1] = __uint_as_float(a.y); x.w[2] = __uint_as_float(b.x);
    x.pixel = b.y; x.sample = c.x; x.stream = c.y;
    return x;
}
__device__ __forceinline__ CandRec load_cand(const CandRec* p)
{
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    CandRec c;
    c.tri[0] = u.x; c.tri[1] = u.y; c.tri[2] = u.z; c.meta = u.w;
    return c;
}
// per-thread tallies -> one atomic per warp (all threads of the block reach this together, after their grid-stride loop)
__device__ __forceinline__ void flush_totals(FrameTotals* totals, const EmitCounters& ec)
{
    if (!totals) return;
    __syncwarp();
    const unsigned s = __reduce_add_sync(0xffffffffu, ec.shadow_rays), o = __reduce_add_sync(0xffffffffu, ec.cand_overflow);
    if ((threadIdx.x & 31u) == 0) {
        if (s) atomicAdd(&totals->rays_shadow, (unsigned long long)s);
        if (o) atomicAdd(&totals->cand_overflow, (unsigned long long)o);
    }
}
__device__ __forceinline__ void flush_trav(TravCounters* cnt, const TravCounters& local)
{
    if (!cnt) return;
    if (local.kd_inner) atomicAdd(&cnt->kd_inner, local.kd_inner);
    if (local.kd_leaves) atomicAdd(&cnt->kd_leaves, local.kd_leaves);
    if (local.tri_tests) atomicAdd(&cnt->tri_tests, local.tri_tests);
    if (local.mesh_queries) atomicAdd(&cnt->mesh_queries, local.mesh_queries);

}

// ---- kernels ----------------------------------------------------------------------------
#define HXR_WALK_BLOCK 128
// resident blocks per SM of the per-ray kernels (register caps 65536 / (128 * blocks)); A/B-ed on B200, see profiles/README.md
#ifndef HXR_GEN_BLOCKS
#define HXR_GEN_BLOCKS 5
#endif
#ifndef HXR_SETUP_BLOCKS
#define HXR_SETUP_BLOCKS 6
#endif
#ifndef HXR_SHADE_GI_BLOCKS
#define HXR_SHADE_GI_BLOCKS 4
#endif
#ifndef HXR_SHADE_WH_BLOCKS
#define HXR_SHADE_WH_BLOCKS 3
#endif
#ifndef HXR_WALK_MIN_BLOCKS
#define HXR_WALK_MIN_BLOCKS 8
#endif

template <bool SIMPLE>
__global__ void __launch_bounds__(128, SIMPLE ? HXR_GEN_BLOCKS : 1) k_gen_primary(DScene sc, FrameParams fp, const uint32_t* __restrict__ pixels,
                                                                                  const uint32_t* __restrict__ pixels_count, uint32_t first_pixel,
                                                                                  uint32_t n_items, uint32_t spp_pass, RayQueue q, CandRec* cand)
{
    if (pixels_count) n_items = min(n_items, *pixels_count * spp_pass);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride) {
        const uint32_t pi = i / spp_pass;
        const uint32_t pixel = pixels ? pixels[pi] : first_pixel + pi;
        gen_primary_item<SIMPLE>(sc, fp, pixel, fp.sample_base + (i % spp_pass) * fp.sample_stride, q, cand, i);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *q.count = n_items;
}

// the inline part of queued rays, in place (reads the 64-byte record, writes back its last 16 bytes), and their slot-0 entry
// records for the walk
template <bool SHADOW, bool COUNT, bool SIMPLE>
__global__ void __launch_bounds__(128, SIMPLE ? HXR_SETUP_BLOCKS : 1) k_setup(DScene sc, RayGeom* geom, MeshEntry* entry, CandRec* cand, const ShadowAux* aux,
                                                                              float* accum, const uint32_t* __restrict__ count, uint32_t cap, TravCounters* cnt)
{
    const uint32_t n = min(*count, cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    TravCounters local = {0, 0, 0, 0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        RayGeom g = load_geom_rw(geom + i);
        const int32_t pre0 = g.pre;
        if (SHADOW) setup_shadow_geom<COUNT, SIMPLE>(sc, g, entry ? entry + i : nullptr, cand ? cand + i : nullptr, aux + i, accum, COUNT ? &local : nullptr);
        else setup_closest_geom<COUNT, SIMPLE>(sc, g, entry ? entry + i : nullptr, cand ? cand + i : nullptr, COUNT ? &local : nullptr);
        if (SHADOW && g.pre == pre0) continue;  // nothing to write back
        double2 tail;
        tail.x = g.limit;
        tail.y = __longlong_as_double((long long)(((unsigned long long)g.depth_flags << 32) | (unsigned long long)(uint32_t)g.pre));
        reinterpret_cast<double2*>(geom + i)[3] = tail;
    }
    if (COUNT) flush_trav(cnt, local);
}

// ---- the KD walk -------------------------------------------------------------------------------------
// One lane = one ray; the lane walks the big meshes the ray enters one after the other. Lanes that run out of work are
// refilled from the ray queue (__ballot_sync finds them, one atomicAdd per warp, __shfl_sync broadcast). Each round has:
//   0. refill (only when enough lanes are idle): finished rays write their candidate record, empty lanes take new rays,
//      lanes between two meshes transform their ray into the next mesh whose box it enters (the kernel's only double
//      arithmetic: the reference's Node::intersect transform, rounded to float together with its error bound);
//   1. `walkSteps` block steps: every lane whose cursor is a tree block pops / steps (block_step: one 32-byte
//      fetch = two tree levels, up to four grandchildren front to back, the far ones pushed on the stack);
//   2. the leaves reached so far are filtered WARP-COOPERATIVELY: the (ray, triangle) pairs of all lanes' leaves are
//      dealt out evenly over the 32 lanes (prefix sum of the leaf sizes + binary search by shuffle), so a lane
//      with a long leaf does not hold up the others.
// Bounding phase 1 keeps lanes from idling while one ray of the warp descends a long path (the measured SIMD
// efficiency of an unbounded while-while loop on incoherent GI rays was 20 %).
//
// The walk itself is FP32 only: conservative plane arithmetic (isect.h: block_step) and the conservative triangle
// filter (tri_filter). Pairs the filter cannot rule out are the ray's CANDIDATES for the exact double test, which runs
// in the kernel that consumes the ray (k_shade / k_resolve_shadow); certain hits shorten the walk.
//
// State per lane in shared memory: the float ray (36 B), the best-hit bound (4 B, lowered with atomicMin by whichever
// lane filters a certain hit), the candidate slots (3 x 4 B + count/slot word), a two-entry mailbox and the first HXR_SSTACK
// stack entries (12 B each); deeper entries overflow to local memory (rare: the stack is shallow for almost all rays).
#define HXR_POP 0x7FFFFFFFu /* cursor value: take the next entry from the stack */
#ifndef HXR_LEAF_BREAK
// phase 1 ends early once this many lanes of the warp hold a leaf (0: always walkSteps steps). A/B on B200 (profiles/r2/ab/r2w_*,
// r2x_*, r2p_*): 12 against 0 - terrain walk 143.1 -> 141.7 ms per 32-spp wave (1856 -> 1868 Mrays/s), beer / meshes / kdtree_test
// at 1080p-class +1.8 / +1.6 / +1.9 %; 8 and 10 the same within noise, 16 and 20 half of it. The opposite rule - extra steps
// while fewer than 4 / 8 lanes hold a leaf - loses (walk 142.2 / 143.1 ms)
#define HXR_LEAF_BREAK 12
#endif

template <int SSTACK>
struct WalkShared {
    float ray[9][HXR_WALK_BLOCK];  // rows 0-2 origin, 3-5 1/direction, 6-8 direction (a leaf child's "axis 3" reads the next row: finite, unused)
    uint32_t tb[HXR_WALK_BLOCK];  // bits of the (non-negative) float bound; 0 = shadow ray certainly blocked
    uint32_t mail[2][HXR_WALK_BLOCK];  // the last two triangles of the current mesh already among the candidates (a triangle sits in several leaves)
    uint32_t cand[HXR_CAND_MAX][HXR_WALK_BLOCK];
    uint32_t meta[HXR_WALK_BLOCK];  // CandRec::meta under construction
    uint32_t stRef[SSTACK][HXR_WALK_BLOCK];
    float stMin[SSTACK][HXR_WALK_BLOCK];
    float stMax[SSTACK][HXR_WALK_BLOCK];
};

// SSTACK = stack entries kept in shared memory: every entry costs 1.5 KB of the SM's 256 KB L1/shared array per block
template <bool SHADOW, bool COUNT, int SSTACK, bool PACKED>
__global__ void __launch_bounds__(HXR_WALK_BLOCK, HXR_WALK_MIN_BLOCKS) k_walk(DScene sc, const RayGeom* __restrict__ geom,
                                                                              const MeshEntry* __restrict__ entry, const uint32_t* __restrict__ count,
                                                                              uint32_t cap, CandRec* __restrict__ cand, uint32_t* head, OverflowEntry* ovf_list,
                                                                              uint32_t* ovf_count, TravCounters* cnt, int walkSteps, int refillMin, int useMail,
                                                                              int branchFreePush)
{
    __shared__ WalkShared<SSTACK> sh;
    constexpr int HXR_SSTACK = SSTACK;
    const unsigned FULL = 0xffffffffu;
    const uint32_t n = min(*count, cap);
    // The grid is sized from a host-side upper bound of the count. Blocks that the actual count does not need leave at once
    // (the blocks before this one have at least n lanes between them): an almost empty queue must not cost a same-address
    // atomic per warp of a full grid (measured: 60-120 us per launch on the deep, nearly empty levels of small Whitted frames)
    if ((uint64_t)blockIdx.x * HXR_WALK_BLOCK >= n) return;
    const int nBig = sc.n_big;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warpBase = tid & ~31u;
    uint32_t ovRef[HXR_KD_STACK - HXR_SSTACK];
    float ovMin[HXR_KD_STACK - HXR_SSTACK], ovMax[HXR_KD_STACK - HXR_SSTACK];
    bool active = false, hasRay = false, drained = false;
    struct SharedRay {  // o(axis) / inv(axis) straight from this lane's shared-memory column: no selects, no registers
        const float* col;
        uint32_t par;
        __device__ __forceinline__ float o(uint32_t axis) const { return col[axis * HXR_WALK_BLOCK]; }
        __device__ __forceinline__ float inv(uint32_t axis) const { return col[(3 + axis) * HXR_WALK_BLOCK]; }
    } wr;
    wr.col = &sh.ray[0][tid];
    wr.par = 0;
    for (int r = 0; r < 9; r++) sh.ray[r][tid] = 0.0f;
    // walked slot 0: what almost every ray walks (its entry arrives precomputed)
    const int mesh0 = slot_mesh(sc, 0);
    const KdBlock* const blocks0 = sc.meshes[mesh0].blocks;
    const uint32_t* const leafTris0 = sc.meshes[mesh0].leaf_tris;
    const void* const tris0 = PACKED ? (const void*)sc.meshes[mesh0].tri_pk : (const void*)sc.meshes[mesh0].tri_f32;
    const bool backface0 = sc.meshes[mesh0].backface != 0;
    const KdBlock* blocks = nullptr;  // the current mesh
    const uint32_t* leafTris = nullptr;
    const void* tris = nullptr;  // TriPacked (32 B) or TriF32 (48 B) records of the current mesh
    bool backface = false;
    int meshIdx = -1;
    uint32_t cur = HXR_POP, leafCnt = 0;
    float tmin = 0, tmax = 0, tbest = 0, err = 0, occ = 0;
    float wcap = 0, kup = 0;  // world-distance cap of this ray (tightened by certain hits), world distance per unit parameter of the current mesh
    int sp = 0;
    uint32_t rayIdx = 0;
    int slot = 0;  // next walked-mesh slot this ray has to try (the mesh being walked is slot - 1)
    TravCounters local = {0, 0, 0, 0, 0};

    auto push = [&](const WalkEnt& e) {
        if (sp < HXR_SSTACK) {
            sh.stRef[sp][tid] = e.ref; sh.stMin[sp][tid] = e.lo; sh.stMax[sp][tid] = e.hi;
        } else if (sp < HXR_KD_STACK) {
            ovRef[sp - HXR_SSTACK] = e.ref; ovMin[sp - HXR_SSTACK] = e.lo; ovMax[sp - HXR_SSTACK] = e.hi;
        } else {
            return;  // the build caps the depth at HXR_KD_MAX_DEPTH (<= 1.5 pushes per level): never reached
        }
        sp++;
    };

    for (;;) {
        // ---- phase 0: refill
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle == FULL || __popc(idle) >= refillMin) {
            // rays that have tried all their meshes leave their candidate record
            if (!active && hasRay && slot >= nBig) {
                reinterpret_cast<uint4*>(cand)[rayIdx] = make_uint4(sh.cand[0][tid], sh.cand[1][tid], sh.cand[2][tid], sh.meta[tid]);
                hasRay = false;
            }
            // empty lanes take new rays
            const unsigned empty = __ballot_sync(FULL, !active && !hasRay);
            if (!drained && empty) {
                const int c = __popc(empty);
                const int leader = __ffs(empty) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(head, (uint32_t)c);
                base = __shfl_sync(FULL, base, leader);
                if (base + (uint32_t)c >= n) drained = true;
                if (!active && !hasRay) {
                    const uint32_t k = base + __popc(empty & ((1u << lane) - 1u));
                    if (k < n) {
                        // the ray's slot-0 entry record, computed by the kernel that set the ray up: three 128-bit loads
                        const uint4* ep = reinterpret_cast<const uint4*>(entry + k);
                        const uint4 w0 = __ldg(ep), w1 = __ldg(ep + 1), w2 = __ldg(ep + 2);
                        if (__uint_as_float(w2.w) >= 0.0f) {  // (a dead ray has its candidate record already)
                            rayIdx = k;
                            hasRay = true;
                            slot = 1;
                            kup = __uint_as_float(w2.w);
                            tbest = __uint_as_float(w2.x);
                            wcap = tbest < 3e38f ? fminf(tbest * kup, 3.4e38f) : 3.4e38f;
                            sh.meta[tid] = 0u;
                            sh.cand[0][tid] = 0u; sh.cand[1][tid] = 0u; sh.cand[2][tid] = 0u;
                            tmin = __uint_as_float(w1.z);
                            tmax = __uint_as_float(w1.w);
                            if (tmin <= tmax) {
                                const float ox = __uint_as_float(w0.x), oy = __uint_as_float(w0.y), oz = __uint_as_float(w0.z);
                                const float dx = __uint_as_float(w0.w), dy = __uint_as_float(w1.x), dz = __uint_as_float(w1.y);
                                const WalkRay w = walk_ray_f(ox, oy, oz, dx, dy, dz);
                                wr.par = w.par;
                                sh.ray[0][tid] = ox; sh.ray[1][tid] = oy; sh.ray[2][tid] = oz;
                                sh.ray[3][tid] = w.ix; sh.ray[4][tid] = w.iy; sh.ray[5][tid] = w.iz;
                                sh.ray[6][tid] = dx; sh.ray[7][tid] = dy; sh.ray[8][tid] = dz;
                                occ = __uint_as_float(w2.y);
                                err = __uint_as_float(w2.z);
                                meshIdx = mesh0;
                                sh.tb[tid] = __float_as_uint(tbest);
                                sh.mail[0][tid] = 0xFFFFFFFFu;
                                sh.mail[1][tid] = 0xFFFFFFFFu;
                                blocks = blocks0;
                                leafTris = leafTris0;
                                tris = tris0;
                                backface = backface0;
                                sp = 0;
                                cur = 0;
                                active = true;
                                if (COUNT) local.mesh_queries++;
                            }
                        }
                    }
                }
            }
            // lanes between two meshes (scenes with several walked nodes): into the next mesh whose box the ray enters -
            // the kernel's only double arithmetic (the reference's Node::intersect transform, rounded to float at the end)
            if (nBig > 1 && !active && hasRay && slot < nBig) {
                const RayGeom g = load_geom(geom + rayIdx);
                {
                    const float of[3] = {(float)g.o[0], (float)g.o[1], (float)g.o[2]}, df[3] = {(float)g.d[0], (float)g.d[1], (float)g.d[2]};
                    while (slot < nBig && box_certainly_missed(sc.big_box + 6 * slot, of, df, wcap)) slot++;
                }
                if (slot < nBig) {
                    Ray ray;
                    ray.o = ld3(g.o);
                    ray.d = ld3(g.d);
                    ray.depth = 0;
                    ray.flags = 0;
                    MeshEntry e;
                    if (enter_mesh<SHADOW>(sc, slot, ray, fmin(g.limit, (double)wcap), e)) {
                        const WalkRay w = walk_ray_f(e.ox, e.oy, e.oz, e.dx, e.dy, e.dz);
                        wr.par = w.par;
                        sh.ray[0][tid] = e.ox; sh.ray[1][tid] = e.oy; sh.ray[2][tid] = e.oz;
                        sh.ray[3][tid] = w.ix; sh.ray[4][tid] = w.iy; sh.ray[5][tid] = w.iz;
                        sh.ray[6][tid] = e.dx; sh.ray[7][tid] = e.dy; sh.ray[8][tid] = e.dz;
                        tmin = e.tmin;
                        tmax = e.tmax;
                        tbest = e.tlimit;
                        occ = e.occ;
                        err = e.err;
                        kup = e.kup;
                        meshIdx = slot_mesh(sc, slot);
                        sh.tb[tid] = __float_as_uint(tbest);
                        sh.mail[0][tid] = 0xFFFFFFFFu;
                        sh.mail[1][tid] = 0xFFFFFFFFu;
                        const DMesh& M = sc.meshes[meshIdx];
                        blocks = M.blocks;
                        leafTris = M.leaf_tris;
                        tris = PACKED ? (const void*)M.tri_pk : (const void*)M.tri_f32;
                        backface = M.backface != 0;
                        sp = 0;
                        cur = 0;
                        active = true;
                        if (COUNT) local.mesh_queries++;
                    }
                    slot++;
                }
            }
        }
        if (__ballot_sync(FULL, active) == 0) {
            if (drained && __ballot_sync(FULL, hasRay) == 0) break;
            continue;
        }
        // ---- phase 1: a few block steps for every lane whose cursor is not a leaf
#pragma unroll 1
        for (int it = 0; it < walkSteps; it++) {
            const bool stepping = active && !(cur >> 31);
            if (stepping && cur == HXR_POP) {
                if (sp == 0) {
                    // this mesh is done: a certain hit at parameter <= tbest lies within world distance tbest * kup, the
                    // ray's later meshes need not look farther (without a certain hit the product is >= the cap: no change)
                    active = false;
                    wcap = fminf(wcap, tbest * kup);
                } else {
                    sp--;
                    WalkEnt e;
                    if (sp < HXR_SSTACK) { e.ref = sh.stRef[sp][tid]; e.lo = sh.stMin[sp][tid]; e.hi = sh.stMax[sp][tid]; }
                    else { e.ref = ovRef[sp - HXR_SSTACK]; e.lo = ovMin[sp - HXR_SSTACK]; e.hi = ovMax[sp - HXR_SSTACK]; }
                    // else it cannot hold a closer hit: keep popping. (One pop per step: draining such entries in a loop right here
                    // was A/B-ed - 1.35 failing pops per closest-hit ray on the terrain - and lost: walk 143.1 -> 145.0 ms, r2p.)
                    if (e.lo <= tbest) { cur = e.ref; tmin = e.lo; tmax = e.hi; }
                }
            }
            __syncwarp();  // lanes that popped and lanes that did not take the block step together
            if (stepping && active && cur < HXR_POP) {
                // one block = a node and both its children: up to four grandchildren, front to back
                if (COUNT) local.kd_inner++;
                const KdBlock B = load_block(blocks + cur);
                WalkEnt e0, e1, e2, e3;
                block_step(B, wr, tmin, tmax, tbest, e0, e1, e2, e3);
                // nearest valid entry becomes the cursor, the others are pushed far-to-near. Branch-free: an entry is pushed iff it
                // is valid and a nearer one is too; its slot follows from the pushes before it; the three stores are predicated.
                // (The chained "if valid { if have push; c = e }" form compiled to ~75 issue slots per step at 2-6 active lanes.)
                const bool v0 = ent_valid(e0), v1 = ent_valid(e1), v2 = ent_valid(e2), v3 = ent_valid(e3);
                if (branchFreePush) {
                    const bool p3 = v3 && (v0 || v1 || v2), p2 = v2 && (v0 || v1), p1 = v1 && v0;
                    const int s3 = sp, s2 = s3 + (p3 ? 1 : 0), s1 = s2 + (p2 ? 1 : 0);
                    sp = s1 + (p1 ? 1 : 0);
                    if (p3 && s3 < HXR_SSTACK) { sh.stRef[s3][tid] = e3.ref; sh.stMin[s3][tid] = e3.lo; sh.stMax[s3][tid] = e3.hi; }
                    if (p2 && s2 < HXR_SSTACK) { sh.stRef[s2][tid] = e2.ref; sh.stMin[s2][tid] = e2.lo; sh.stMax[s2][tid] = e2.hi; }
                    if (p1 && s1 < HXR_SSTACK) { sh.stRef[s1][tid] = e1.ref; sh.stMin[s1][tid] = e1.lo; sh.stMax[s1][tid] = e1.hi; }
                    if (sp > HXR_SSTACK) {  // rare: some of them belong to the overflow part of the stack (local memory)
                        if (p3 && s3 >= HXR_SSTACK && s3 < HXR_KD_STACK) { ovRef[s3 - HXR_SSTACK] = e3.ref; ovMin[s3 - HXR_SSTACK] = e3.lo; ovMax[s3 - HXR_SSTACK] = e3.hi; }
                        if (p2 && s2 >= HXR_SSTACK && s2 < HXR_KD_STACK) { ovRef[s2 - HXR_SSTACK] = e2.ref; ovMin[s2 - HXR_SSTACK] = e2.lo; ovMax[s2 - HXR_SSTACK] = e2.hi; }
                        if (p1 && s1 >= HXR_SSTACK && s1 < HXR_KD_STACK) { ovRef[s1 - HXR_SSTACK] = e1.ref; ovMin[s1 - HXR_SSTACK] = e1.lo; ovMax[s1 - HXR_SSTACK] = e1.hi; }
                        if (sp > HXR_KD_STACK) sp = HXR_KD_STACK;  // never reached: the build caps the depth (see push)
                    }
                    cur = v0 ? e0.ref : v1 ? e1.ref : v2 ? e2.ref : v3 ? e3.ref : HXR_POP;
                    tmin = v0 ? e0.lo : v1 ? e1.lo : v2 ? e2.lo : e3.lo;
                    tmax = v0 ? e0.hi : v1 ? e1.hi : v2 ? e2.hi : e3.hi;
                } else {
                    WalkEnt c;
                    c.ref = HXR_POP; c.lo = 0; c.hi = 0;
                    bool have = false;
                    if (v3) { c = e3; have = true; }
                    if (v2) { if (have) push(c); c = e2; have = true; }
                    if (v1) { if (have) push(c); c = e1; have = true; }
                    if (v0) { if (have) push(c); c = e0; have = true; }
                    cur = c.ref; tmin = c.lo; tmax = c.hi;
                }
            }
            if (stepping && active && (cur >> 31)) leafCnt = __ldg(leafTris + (cur & ~HXR_KD_LEAF));  // in flight while the others keep stepping
            __syncwarp();
#if HXR_LEAF_BREAK
            // enough lanes wait at a leaf to fill the cooperative rounds of phase 2: do not let them idle through the remaining steps
            if (__popc(__ballot_sync(FULL, active && (cur >> 31))) >= HXR_LEAF_BREAK) break;
#endif
        }
        // ---- phase 2: all (ray, triangle) pairs of the leaves held by this warp, dealt out over its 32 lanes
        const bool hasLeaf = active && (cur >> 31);
        if (__ballot_sync(FULL, hasLeaf) == 0) continue;
        const uint32_t cntMine = hasLeaf ? leafCnt : 0u;
        uint32_t incl = cntMine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, incl, d);
            if ((int)lane >= d) incl += v;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        const uint32_t excl = incl - cntMine;
        const uint32_t myFirst = (cur & ~HXR_KD_LEAF) + 1u;
        for (uint32_t base = 0; base < total; base += 32u) {
            const uint32_t pair = base + lane;
            int o = 0;  // owner of this pair: the first lane whose inclusive prefix exceeds it
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const uint32_t v = __shfl_sync(FULL, incl, o + step - 1);
                if (v <= pair) o += step;
            }
            const uint32_t oExcl = __shfl_sync(FULL, excl, o);
            const uint32_t oFirst = __shfl_sync(FULL, myFirst, o);
            const int oMesh = __shfl_sync(FULL, meshIdx, o);
            const float oErr = __shfl_sync(FULL, err, o);
            const int oSlot = __shfl_sync(FULL, slot, o) - 1;
            const float oOcc = SHADOW ? __shfl_sync(FULL, occ, o) : 0.0f;
            if (pair < total) {
                const uint32_t* lt = leafTris;
                const void* tt = tris;
                bool bf = backface;
                if (oMesh != meshIdx) {
                    const DMesh& M = sc.meshes[oMesh];
                    lt = M.leaf_tris; tt = PACKED ? (const void*)M.tri_pk : (const void*)M.tri_f32; bf = M.backface != 0;
                }
                const uint32_t ti = __ldg(lt + oFirst + (pair - oExcl));
                const unsigned ot = warpBase | (unsigned)o;
                const float oBest = __uint_as_float(sh.tb[ot]);  // the freshest bound (other lanes may have lowered it this round)
                float ghi = 0;
                int cls = HXR_TF_MISS;
                if (ti != sh.mail[0][ot] && ti != sh.mail[1][ot]) {  // not already a candidate from a neighbouring leaf
                    if (PACKED) cls = tri_filter_packed(static_cast<const TriPacked*>(tt) + ti, bf, sh.ray[0][ot], sh.ray[1][ot], sh.ray[2][ot], sh.ray[6][ot], sh.ray[7][ot], sh.ray[8][ot], oErr, oBest, ghi);
                    else cls = tri_filter(static_cast<const TriF32*>(tt) + ti, bf, sh.ray[0][ot], sh.ray[1][ot], sh.ray[2][ot], sh.ray[6][ot], sh.ray[7][ot], sh.ray[8][ot], oErr, oBest, ghi);
                }
                bool emit = false;
                if (cls == HXR_TF_CERTAIN) {
                    if (SHADOW && ghi < oOcc) atomicMin(&sh.tb[ot], 0u);  // certainly blocked: no exact test needed
                    else { atomicMin(&sh.tb[ot], __float_as_uint(ghi)); emit = true; }
                } else if (cls == HXR_TF_MAYBE) {
                    emit = true;
                }
                // a surviving pair becomes a candidate of its owner's ray: a slot in the owner's shared-memory record (lanes
                // serving the same owner take different slots through the atomic; the count saturates far below its 8-bit field)
                if (emit) {
                    const uint32_t m0 = sh.meta[ot];
                    const uint32_t n0 = min(m0 & 0xFFu, (uint32_t)HXR_CAND_MAX);
                    bool dup = false;  // already recorded (from a neighbouring leaf, longer ago than the mailbox remembers)
#pragma unroll
                    for (uint32_t i = 0; i < HXR_CAND_MAX; i++) dup = dup || (i < n0 && sh.cand[i][ot] == ti && ((m0 >> (8u + 8u * i)) & 0xFFu) == (uint32_t)oSlot);
                    if (!dup && (m0 & 0xFFu) < 100u) {
                        const uint32_t pos = atomicAdd(&sh.meta[ot], 1u) & 0xFFu;
                        if (pos < HXR_CAND_MAX) {
                            sh.cand[pos][ot] = ti;
                            if (oSlot) atomicOr(&sh.meta[ot], (uint32_t)oSlot << (8u + 8u * pos));
                        }
                    }
                    if (useMail) {
                        sh.mail[1][ot] = sh.mail[0][ot];  // (lanes emitting for the same owner race here: any of their triangles is a valid entry)
                        sh.mail[0][ot] = ti;
                    }
                }
            }
        }
        __syncwarp();
        if (hasLeaf) {
            if (COUNT) { local.kd_leaves++; local.tri_tests += leafCnt; }
            const uint32_t tb = sh.tb[tid];
            if (!(SHADOW && tb == 0u) && (sh.meta[tid] & 0xFFu) > HXR_CAND_MAX) {
                // the record is full (rare: distant grazing rays on which the float bounds decide nothing): the ray stops here and
                // is listed for k_finish, which resumes at this leaf with exact tests on the spot
                OverflowEntry oe;
                oe.ray = rayIdx;
                oe.slot = slot - 1;
                oe.tstop = tmin;
                oe.pad = 0;
                ovf_list[atomicAdd(ovf_count, 1u)] = oe;
                active = false;
                slot = nBig;
            }
            cur = HXR_POP;
            tbest = __uint_as_float(tb);
            if (SHADOW && tb == 0u) {  // certainly blocked: the ray is finished
                sh.meta[tid] = HXR_CAND_BLOCKED;
                active = false;
                slot = nBig;
            }
        }
    }
    if (COUNT) flush_trav(cnt, local);
}

// the rays whose candidate record filled up: resumed where they stopped, exact tests on the spot (one ray per thread)
template <bool SHADOW, bool COUNT>
__global__ void __launch_bounds__(128, 2) k_finish(DScene sc, const RayGeom* __restrict__ geom, CandRec* cand, const OverflowEntry* __restrict__ ovf_list,
                                                   const uint32_t* __restrict__ ovf_count, uint32_t cap, FrameTotals* totals, TravCounters* cnt)
{
    const uint32_t n = min(*ovf_count, cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    TravCounters local = {0, 0, 0, 0, 0};
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(ovf_list + k));
        OverflowEntry oe;
        oe.ray = u.x; oe.slot = (int32_t)u.y; oe.tstop = __uint_as_float(u.z); oe.pad = 0;
        const uint4 c = reinterpret_cast<const uint4*>(cand)[oe.ray];
        CandRec rec;
        rec.tri[0] = c.x; rec.tri[1] = c.y; rec.tri[2] = c.z; rec.meta = c.w;
        const CandRec r = finish_overflowed_ray<SHADOW, COUNT>(sc, load_geom(geom + oe.ray), rec, oe, COUNT ? &local : nullptr);
        reinterpret_cast<uint4*>(cand)[oe.ray] = make_uint4(r.tri[0], r.tri[1], r.tri[2], r.meta);
    }
    if (totals && blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&totals->cand_overflow, (unsigned long long)n);
    if (COUNT) flush_trav(cnt, local);
}

// The same, ONE WARP per overflowed ray (HXR_FINISH_WARP, the default). The rays that arrive here are the long grazing ones on
// which the float bounds decide nothing: they cross hundreds of leaves and every triangle of every leaf needs the exact double
// test, i.e. a dependent chain of index -> filter record -> 96-byte exact record per TRIANGLE when one thread does it. ncu on the
// one-thread form (profiles/r2/ncu_deep_bounce_r2_v1.txt): 2.97 lanes per instruction, 4.5 % warps active, the launch lasts as
// long as its longest ray (0.78 ms for 41 k rays). Here the 32 lanes walk the tree together (the block steps are uniform: the
// same loads, broadcast) and split each leaf's triangles among them, so a leaf costs one such chain, not one per triangle;
// the leaf's winner is found with a shuffle reduction under the rule of tri_core (smallest parameter, then highest index), which
// does not depend on the order of the tests - the verdict equals finish_overflowed_ray's (tests/test_gpu_parity.py:
// test_finish_forms_agree). Rays are fetched one at a time through an atomic cursor: the long ones do not pile up on one warp.
template <bool SHADOW, bool COUNT>
__device__ __forceinline__ CandRec finish_overflowed_ray_warp(const DScene& sc, const RayGeom& g, const CandRec& rec, const OverflowEntry& ov, TravCounters* cnt,
                                                              unsigned lane, uint32_t* stRef, float* stLo, float* stHi)
{
    const unsigned FULL = 0xffffffffu;
    CandRec out;
    out.tri[0] = out.tri[1] = out.tri[2] = 0;
    out.meta = 0;
    const Ray ray = geom_ray(g);
    double bestDist = g.limit;  // closest: inline winner so far; shadow: |AB|
    int bestNode = SHADOW ? 0x7FFFFFFF : g.pre;
    MeshWinner win;
    win.node = -1;
    win.mb.tri = -1;
    // true: the hit just recorded in mb blocks a shadow ray
    auto blocks = [&](const hxr_node& nd, const Ray& t, const MeshBest& mb) -> bool {
        const d3 ipw = node_mul_m(nd, t.o + mb.gamma * t.d) + ld3(nd.T.offset);
        return distance3(ray.o, ipw) < g.limit;
    };
    // the recorded candidates (at most three): every lane runs them, uniformly - see finish_overflowed_ray for the grouping
    MeshBest cur;
    bool haveCur = false;
    cur.tri = -1;
    cur.gamma = 0;
    cur.l2 = cur.l3 = 0;
    {
        int gs = -1;
        MeshBest mb;
        mb.tri = -1;
        mb.gamma = 0;
        mb.l2 = mb.l3 = 0;
        Ray t = ray;
        for (uint32_t i = 0; i <= HXR_CAND_MAX; i++) {
            const int s = i < HXR_CAND_MAX ? (int)((rec.meta >> (8u + 8u * i)) & 0xFFu) : -2;
            if (s != gs) {
                if (gs >= 0) {
                    if (gs == ov.slot) { cur = mb; haveCur = true; }
                    else if (!SHADOW) fold_mesh_hit(sc.nodes[sc.big_nodes[gs]], sc.big_nodes[gs], ray, t, mb, bestDist, bestNode, win);
                }
                if (s < 0) break;
                gs = s;
                const hxr_node& nd = sc.nodes[sc.big_nodes[s]];
                t = object_ray(nd, ray);
                mb.gamma = gamma_limit_for(nd, t, g.limit);
                mb.tri = -1;
                mb.l2 = mb.l3 = 0;
            }
            const hxr_node& nd = sc.nodes[sc.big_nodes[s]];
            const DMesh& M = node_mesh(sc, nd);
            if (tri_test(M.tri_test, M.backface != 0, t, rec.tri[i], mb) && SHADOW && blocks(nd, t, mb)) {
                out.meta = HXR_CAND_BLOCKED;
                return out;
            }
        }
    }
    float wcap = f32_cap(bestDist);
    const float of[3] = {(float)g.o[0], (float)g.o[1], (float)g.o[2]}, df[3] = {(float)g.d[0], (float)g.d[1], (float)g.d[2]};
    for (int slot = ov.slot; slot < sc.n_big; slot++) {
        const int node = sc.big_nodes[slot];
        const hxr_node& nd = sc.nodes[node];
        MeshBest mb;
        mb.tri = -1;
        mb.l2 = mb.l3 = 0;
        bool walked = false;
        MeshEntry e;
        if ((slot == ov.slot || !box_certainly_missed(sc.big_box + 6 * slot, of, df, wcap)) && enter_mesh<SHADOW>(sc, slot, ray, fmin(g.limit, (double)wcap), e)) {
            walked = true;
            const DMesh& M = sc.meshes[slot_mesh(sc, slot)];
            const Ray t = object_ray(nd, ray);
            mb.gamma = gamma_limit_for(nd, t, fmin(g.limit, (double)wcap));
            if (slot == ov.slot && haveCur) mb = cur;
            if (COUNT && cnt) cnt->mesh_queries++;
            const WalkRay w = walk_ray_f(e.ox, e.oy, e.oz, e.dx, e.dy, e.dz);
            const bool bf = M.backface != 0;
            // (segments that end before the stop point were examined by the first walk)
            const float tskip = slot == ov.slot ? ov.tstop * (ov.tstop > 0 ? 1.0f - 1e-6f : 1.0f + 1e-6f) - 1e-30f : -INFINITY;
            float tmin = e.tmin, tmax = e.tmax, tbest = fminf(e.tlimit, f32_above(mb.gamma));
            int sp = 0;
            uint32_t curRef = 0;
            for (;;) {
                if (curRef & HXR_KD_LEAF) {
                    const uint32_t* list = M.leaf_tris + (curRef & ~HXR_KD_LEAF);
                    const uint32_t nT = __ldg(list);
                    if (COUNT && cnt) { cnt->tri_tests += nT; cnt->kd_leaves++; }
                    for (uint32_t base = 0; base < nT; base += 32u) {
                        // lane L: triangle base + L of the leaf, filtered in float, then the exact test against the leaf's starting best
                        bool hit = false;
                        double gm = 0, l2 = 0, l3 = 0;
                        uint32_t ti = 0;
                        if (base + lane < nT) {
                            ti = __ldg(list + 1u + base + lane);
                            float ghi;
                            const int cls = sc.walk_packed ? tri_filter_packed(M.tri_pk + ti, bf, e.ox, e.oy, e.oz, e.dx, e.dy, e.dz, e.err, tbest, ghi)
                                                           : tri_filter(M.tri_f32 + ti, bf, e.ox, e.oy, e.oz, e.dx, e.dy, e.dz, e.err, tbest, ghi);
                            if (cls != HXR_TF_MISS) hit = tri_core(M.tri_test, bf, t, ti, mb.gamma, mb.tri, gm, l2, l3);
                        }
                        if (__ballot_sync(FULL, hit) == 0u) continue;
                        // the winner among the lanes' hits: smallest parameter, then highest index (tri_core's own rule, so the
                        // result is what testing them one after the other leaves in mb)
                        bool bv = hit;
                        double bg = gm;
                        uint32_t bt = ti;
                        unsigned bl = lane;
#pragma unroll
                        for (int d = 16; d >= 1; d >>= 1) {
                            const bool ov_ = __shfl_xor_sync(FULL, (int)bv, d) != 0;
                            const double og = __shfl_xor_sync(FULL, bg, d);
                            const uint32_t ot = __shfl_xor_sync(FULL, bt, d);
                            const unsigned ol = __shfl_xor_sync(FULL, bl, d);
                            if (ov_ && (!bv || og < bg || (og == bg && ot > bt))) { bv = true; bg = og; bt = ot; bl = ol; }
                        }
                        mb.gamma = bg;
                        mb.tri = (int)bt;
                        mb.l2 = __shfl_sync(FULL, l2, bl);
                        mb.l3 = __shfl_sync(FULL, l3, bl);
                        if (SHADOW && blocks(nd, t, mb)) {
                            out.meta = HXR_CAND_BLOCKED;
                            return out;
                        }
                        tbest = fminf(tbest, f32_above(mb.gamma));
                    }
                } else {
                    if (COUNT && cnt) cnt->kd_inner++;
                    WalkEnt en[4];
                    block_step(load_block(M.blocks + curRef), w, tmin, tmax, tbest, en[0], en[1], en[2], en[3]);
                    bool have = false;
                    WalkEnt c;
                    c.ref = 0; c.lo = c.hi = 0;
#pragma unroll
                    for (int k = 3; k >= 0; k--) {
                        if (!ent_valid(en[k]) || en[k].hi < tskip) continue;
                        if (have && sp < HXR_KD_STACK) { stRef[sp] = c.ref; stLo[sp] = c.lo; stHi[sp] = c.hi; sp++; }  // (every lane writes the same entry)
                        c = en[k];
                        have = true;
                    }
                    if (have) { curRef = c.ref; tmin = c.lo; tmax = c.hi; continue; }
                }
                bool found = false;
                while (sp > 0) {
                    sp--;
                    if (stLo[sp] <= tbest) { curRef = stRef[sp]; tmin = stLo[sp]; tmax = stHi[sp]; found = true; break; }
                }
                if (!found) break;
            }
        } else if (slot == ov.slot && haveCur) {
            mb = cur;  // (cannot happen: the first walk entered this mesh)
            walked = true;
        }
        if (walked && !SHADOW) {
            const Ray t = object_ray(nd, ray);
            fold_mesh_hit(nd, node, ray, t, mb, bestDist, bestNode, win);
            wcap = fminf(wcap, f32_cap(bestDist));
        }
    }
    if (SHADOW) return out;
    if (win.node >= 0 && bestNode == win.node) {
        out.tri[0] = (uint32_t)win.mb.tri;
        out.meta = 1u | ((uint32_t)sc.node_slot[win.node] << 8);
    }
    return out;
}

template <bool SHADOW, bool COUNT>
__global__ void __launch_bounds__(128, 4) k_finish_warp(DScene sc, const RayGeom* __restrict__ geom, CandRec* cand, const OverflowEntry* __restrict__ ovf_list,
                                                        const uint32_t* __restrict__ ovf_count, uint32_t* fetch, uint32_t cap, FrameTotals* totals,
                                                        TravCounters* cnt)
{
    __shared__ uint32_t sRef[4][HXR_KD_STACK];  // one traversal stack per warp
    __shared__ float sLo[4][HXR_KD_STACK], sHi[4][HXR_KD_STACK];
    const unsigned FULL = 0xffffffffu;
    const uint32_t n = min(*ovf_count, cap);
    const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    TravCounters local = {0, 0, 0, 0, 0};
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(fetch, 1u);
        k = __shfl_sync(FULL, k, 0);
        if (k >= n) break;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(ovf_list + k));
        OverflowEntry oe;
        oe.ray = u.x; oe.slot = (int32_t)u.y; oe.tstop = __uint_as_float(u.z); oe.pad = 0;
        const uint4 c = reinterpret_cast<const uint4*>(cand)[oe.ray];
        CandRec rec;
        rec.tri[0] = c.x; rec.tri[1] = c.y; rec.tri[2] = c.z; rec.meta = c.w;
        __syncwarp();  // (the previous ray's stack is dead before this one's first push)
        const CandRec r = finish_overflowed_ray_warp<SHADOW, COUNT>(sc, load_geom(geom + oe.ray), rec, oe, COUNT && lane == 0 ? &local : nullptr, lane, sRef[wib],
                                                                    sLo[wib], sHi[wib]);
        if (lane == 0) reinterpret_cast<uint4*>(cand)[oe.ray] = make_uint4(r.tri[0], r.tri[1], r.tri[2], r.meta);
    }
    if (totals && blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&totals->cand_overflow, (unsigned long long)n);
    if (COUNT) flush_trav(cnt, local);
}

// GI (one path vertex) and Whitted (shader tree with its stacks) are separate compilations, and so are scenes whose inline
// nodes are only planes / spheres / cubes / quads (SIMPLE: no CSG, heightfield or inline tree-walk code, no stack frame).
//
// Material-sorted shading (the reference dispatches per hit through Shader::computeColor virtuals, src/shading.cpp:67-343,
// src/main.cpp:105-113). SORT variants work on 128 queued rays per block at a time, in two phases:
//   1. every thread resolves its own ray (exact tests, winner across nodes); rays that end here (miss, light, depth guard) add
//      their colour and drop out;
//   2. if some warp holds hits of more than one material, the surviving hits are ranked by the shader of the node they hit,
//      through shared memory (the IntersectionInfo travels as 15 doubles, one column per ray: conflict-free), and thread t
//      shades the hit of rank t; otherwise every thread shades its own hit straight from registers.
// So a warp of phase 2 runs ONE material's code (Lambert + its light loop, Phong, the Layered tree of one object, Refl, Refr)
// unless a material's run straddles it, and the lanes freed by the dropped rays are compacted into whole idle warps. No extra
// pass over DRAM and no extra launch: the sort key only exists once the hit is resolved, and it stays on chip.
#ifndef HXR_SHADE_SORT
#define HXR_SHADE_SORT 1  // bit 0: Whitted, bit 1: GI
#endif
struct ShadeSortShared {
    double hit[16][128];  // dist, ip, norm, dNdx, dNdy, u, v of each resolved hit
    int32_t node[128];
    uint32_t ray[128];    // queue index of the ray
    int32_t key[128];     // shader index of the node hit, or -1
    uint16_t order[128];
};

template <bool GI, bool COUNT, bool SIMPLE>
__global__ void __launch_bounds__(128, SIMPLE ? (GI ? HXR_SHADE_GI_BLOCKS : HXR_SHADE_WH_BLOCKS) : 1)
    k_shade(DScene sc, FrameParams fp, RayQueue q, const CandRec* __restrict__ cand, uint32_t begin, uint32_t end, Sinks sinks, FrameTotals* totals,
            TravCounters* cnt)
{
    const uint32_t e = min(end, min(*q.count, q.cap));
    const uint32_t stride = gridDim.x * blockDim.x;
    EmitCounters ec = {0, 0};
    TravCounters local = {0, 0, 0, 0, 0};
    if ((HXR_SHADE_SORT >> (GI ? 1 : 0)) & 1) {
        __shared__ ShadeSortShared sh;
        const uint32_t t = threadIdx.x;
        for (uint32_t base = begin + blockIdx.x * blockDim.x; base < e; base += stride) {  // (block-uniform trip count)
            const uint32_t i = base + t;
            int key = -1;
            PathState p;
            Hit info;
            int node = -1;
            if (i < e) {
                const RayGeom g = load_geom(q.geom + i);
                const RayAux a = load_aux(q.aux + i);
                CandRec cr;
                cr.meta = 0;
                if (sc.n_big) cr = load_cand(cand + i);
                p = path_state(g, a);
                f3 early;
                if (!resolve_closest<COUNT, SIMPLE>(sc, p.ray, g.limit, g.pre, cr, info, node, early, ec, COUNT ? &local : nullptr)) accum_pixel(sinks.accum, p.pixel, p.w * early);
                else key = sc.nodes[node].shader;
            }
            // a warp whose hits already share one material has nothing to gain; the block only permutes when some warp is mixed
            const unsigned hits = __ballot_sync(0xffffffffu, key >= 0);
            const int first = __shfl_sync(0xffffffffu, key, hits ? __ffs(hits) - 1 : 0);
            sh.key[t] = key;
            if (__syncthreads_or(key >= 0 && key != first)) {
                if (key >= 0) {
                    sh.hit[0][t] = info.dist;
                    sh.hit[1][t] = info.ip.x; sh.hit[2][t] = info.ip.y; sh.hit[3][t] = info.ip.z;
                    sh.hit[4][t] = info.norm.x; sh.hit[5][t] = info.norm.y; sh.hit[6][t] = info.norm.z;
                    sh.hit[7][t] = info.dNdx.x; sh.hit[8][t] = info.dNdx.y; sh.hit[9][t] = info.dNdx.z;
                    sh.hit[10][t] = info.dNdy.x; sh.hit[11][t] = info.dNdy.y; sh.hit[12][t] = info.dNdy.z;
                    sh.hit[13][t] = info.u; sh.hit[14][t] = info.v;
                    sh.node[t] = node;
                    sh.ray[t] = i;
                    // rank among the hits: by key, ties by thread (128 broadcast reads)
                    uint32_t rank = 0;
                    for (uint32_t j = 0; j < 128; j++) {
                        const int kj = sh.key[j];
                        rank += (kj >= 0 && (kj < key || (kj == key && j < t))) ? 1u : 0u;
                    }
                    sh.order[rank] = (uint16_t)t;
                }
                const uint32_t nHits = (uint32_t)__syncthreads_count(key >= 0);
                key = -1;
                if (t < nHits) {
                    const uint32_t s = sh.order[t];
                    const uint32_t ri = sh.ray[s];
                    p = path_state(load_geom(q.geom + ri), load_aux(q.aux + ri));
                    info.dist = sh.hit[0][s];
                    info.ip = mk3(sh.hit[1][s], sh.hit[2][s], sh.hit[3][s]);
                    info.norm = mk3(sh.hit[4][s], sh.hit[5][s], sh.hit[6][s]);
                    info.dNdx = mk3(sh.hit[7][s], sh.hit[8][s], sh.hit[9][s]);
                    info.dNdy = mk3(sh.hit[10][s], sh.hit[11][s], sh.hit[12][s]);
                    info.u = sh.hit[13][s];
                    info.v = sh.hit[14][s];
                    node = sh.node[s];
                    key = 0;
                }
            }
            if (key >= 0) {
                info.geom = -1;
                if (GI) shade_gi_item<COUNT, SIMPLE>(sc, fp, p, info, node, sinks, ec, COUNT ? &local : nullptr);
                else shade_whitted_item<COUNT, SIMPLE>(sc, fp, p, info, node, sinks, ec, COUNT ? &local : nullptr);
            }
            // (no barrier here: sh.key is last read before the count barrier above, the columns and sh.order before this thread
            // reaches the next batch's first barrier)
        }
    } else {
        for (uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < e; i += stride) {
            const RayGeom g = load_geom(q.geom + i);
            const RayAux a = load_aux(q.aux + i);
            CandRec cr;
            cr.meta = 0;
            if (sc.n_big) cr = load_cand(cand + i);
            // (A/B-ed and dropped, r2r: requesting the thread's next candidate record here and prefetching the exact records and
            // attributes it names into L2 - shade 45.5 -> 48.0 ms; the kernel is bound by issue and code size, not by the chain)
            shade_item<GI, COUNT, SIMPLE>(sc, fp, g, a, cr, sinks, ec, COUNT ? &local : nullptr);
        }
    }
    flush_totals(totals, ec);
    if (totals && blockIdx.x == 0 && threadIdx.x == 0 && e > begin) atomicAdd(&totals->rays_closest, (unsigned long long)(e - begin));
    if (COUNT) flush_trav(cnt, local);
}

template <bool COUNT>
__global__ void __launch_bounds__(128, COUNT ? 1 : 4) k_resolve_shadow(DScene sc, ShadowQueue q, const CandRec* __restrict__ cand, float* accum, uint8_t* visible,
                                                                       FrameTotals* totals, TravCounters* cnt)
{
    const uint32_t n = min(*q.count, q.cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    EmitCounters ec = {0, 0};
    TravCounters local = {0, 0, 0, 0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        CandRec cr;
        cr.meta = 0;
        if (sc.n_big) cr = load_cand(cand + i);
        if (visible) visible[i] = resolve_visible<COUNT>(sc, q.geom + i, cr, ec, COUNT ? &local : nullptr) ? 1 : 0;
        else resolve_shadow_item<COUNT>(sc, q.geom + i, q.aux + i, cr, accum, ec, COUNT ? &local : nullptr);
    }
    flush_totals(totals, ec);
    if (COUNT) flush_trav(cnt, local);
}

__global__ void __launch_bounds__(128, 1) k_hit_records(DScene sc, RayQueue q, const CandRec* __restrict__ cand, HitRec* __restrict__ hits)
{
    const uint32_t n = min(*q.count, q.cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        CandRec cr;
        cr.meta = 0;
        if (sc.n_big) cr = load_cand(cand + i);
        HitRec h;
        hit_record_item<false, false>(sc, load_geom(q.geom + i), cr, h, nullptr);
        hits[i] = h;
    }
}

__global__ void k_aa_detect(const float* __restrict__ vfb, int W, int H, int shard_index, int shard_count, uint32_t* list, uint32_t* n_out,
                            uint8_t* mask)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const bool mine = shard_count <= 1 || ((y / HXR_ROW_BAND) % shard_count) == shard_index;
    const bool f = mine && aa_detect_item(vfb, W, H, x, y);
    mask[(size_t)y * W + x] = f;
    if (f) list[atomicAdd(n_out, 1u)] = (uint32_t)(y * W + x);
}

__global__ void k_scale_listed(float* vfb, const uint32_t* __restrict__ list, const uint32_t* __restrict__ n, uint32_t cap, float mul)
{
    const uint32_t m = min(*n, cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        float* p = vfb + 3 * (size_t)list[i];
        p[0] *= mul; p[1] *= mul; p[2] *= mul;
    }
}
__global__ void k_scale_all(float* buf, size_t n, float mul)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] *= mul;
}
__global__ void k_add_into(float* dst, const float* __restrict__ src, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += src[i];
}
__global__ void k_stereo_mix(float* out, const float* __restrict__ L, const float* __restrict__ R, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) stereo_mix_item(out, L, R, i);
}
__global__ void k_to_bmp_rows(const float* __restrict__ rgb, int W, int H, int rowsz, const uint8_t* __restrict__ lut, uint8_t* out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < W && y < H) bmp_pixel_item(rgb, W, H, rowsz, lut, out, x, y);
}
__global__ void k_to_exr_rows(const float* __restrict__ rgb, int W, int H, uint16_t* out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < W && y < H) exr_pixel_item(rgb, W, out, x, y);
}

// ---- launchers --------------------------------------------------------------------------
// per-item kernels are grid-stride loops over a count that only the device knows
static uint32_t stage_grid(const Context* c, uint32_t n, int perSm = 16)
{
    return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)n + 127) / 128, (uint64_t)c->sms * perSm));
}

int gen_primary(Context* c, const DScene& sc, const FrameParams& fp, const uint32_t* pixels, const uint32_t* pixels_count, uint32_t first_pixel,
                uint32_t n_items, uint32_t spp_pass, const RayQueue& q, CandRec* cand)
{
    LaunchScope ls(c, PROF_GEN);
    const uint32_t blocks = stage_grid(c, n_items, 32);
    RayQueue qq = q;
    if (!sc.n_big) { qq.entry = nullptr; cand = nullptr; }
    if (sc.simple_inline) k_gen_primary<true><<<blocks, 128, 0, c->stream>>>(sc, fp, pixels, pixels_count, first_pixel, n_items, spp_pass, qq, cand);
    else k_gen_primary<false><<<blocks, 128, 0, c->stream>>>(sc, fp, pixels, pixels_count, first_pixel, n_items, spp_pass, qq, cand);
    return 1;
}

template <bool SHADOW> static void launch_setup(Context* c, const DScene& sc, RayGeom* geom, MeshEntry* entry, CandRec* cand, const ShadowAux* aux, float* accum,
                                                const uint32_t* count, uint32_t cap, TravCounters* cnt, uint32_t n_hint)
{
    const uint32_t blocks = stage_grid(c, n_hint);
    if (!sc.n_big) { entry = nullptr; cand = nullptr; }
    // counting builds and scenes with CSG / heightfield / inline tree walks take the generic variant
    if (cnt) k_setup<SHADOW, true, false><<<blocks, 128, 0, c->stream>>>(sc, geom, entry, cand, aux, accum, count, cap, cnt);
    else if (sc.simple_inline) k_setup<SHADOW, false, true><<<blocks, 128, 0, c->stream>>>(sc, geom, entry, cand, aux, accum, count, cap, nullptr);
    else k_setup<SHADOW, false, false><<<blocks, 128, 0, c->stream>>>(sc, geom, entry, cand, aux, accum, count, cap, nullptr);
}
int setup_closest(Context* c, const DScene& sc, const RayQueue& q, CandRec* cand, TravCounters* cnt, uint32_t n_hint)
{
    LaunchScope ls(c, PROF_SETUP);
    launch_setup<false>(c, sc, q.geom, q.entry, cand, nullptr, nullptr, q.count, q.cap, cnt, n_hint);
    return 1;
}
int setup_shadow(Context* c, const DScene& sc, const ShadowQueue& q, CandRec* cand, float* accum, TravCounters* cnt, uint32_t n_hint)
{
    LaunchScope ls(c, PROF_SETUP);
    launch_setup<true>(c, sc, q.geom, q.entry, cand, q.aux, accum, q.count, q.cap, cnt, n_hint);
    return 1;
}

template <bool SHADOW, bool COUNT, int SSTACK, bool PACKED>
static void launch_walk_s(Context* c, const DScene& sc, const RayGeom* geom, const MeshEntry* entry, const uint32_t* count, uint32_t cap, const WalkBuffers& wb,
                          FrameTotals* totals, TravCounters* cnt, uint32_t n_hint)
{
    int& full = c->walkGrid[SHADOW][COUNT][SSTACK == 9 ? 0 : (SSTACK == 12 ? 2 : 1)][PACKED];
    if (!full) {
        // fewer resident blocks than fit + a smaller shared-memory carve-out leave the L1 more room for the top of the tree
        if (c->walkCarveout >= 0) cudaFuncSetAttribute(k_walk<SHADOW, COUNT, SSTACK, PACKED>, cudaFuncAttributePreferredSharedMemoryCarveout, c->walkCarveout);
        int perSm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_walk<SHADOW, COUNT, SSTACK, PACKED>, HXR_WALK_BLOCK, 0) != cudaSuccess || perSm < 1) perSm = 1;
        full = c->sms * perSm;
        if (c->walkBlocksPerSm > 0) full = std::min(full, c->sms * c->walkBlocksPerSm);
    }
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)full, ((uint64_t)n_hint + HXR_WALK_BLOCK - 1) / HXR_WALK_BLOCK));
    k_walk<SHADOW, COUNT, SSTACK, PACKED><<<grid, HXR_WALK_BLOCK, 0, c->stream>>>(sc, geom, entry, count, cap, wb.cand, wb.head, wb.ovf_list, wb.ovf_count, cnt, c->walkSteps,
                                                                                  c->refillMin, c->useMail, c->bfPush);
}
template <bool SHADOW, bool PACKED>
static void launch_walk_p(Context* c, const DScene& sc, const RayGeom* geom, const MeshEntry* entry, const uint32_t* count, uint32_t cap, const WalkBuffers& wb,
                          FrameTotals* totals, TravCounters* cnt, uint32_t n_hint)
{
    if (cnt) { launch_walk_s<SHADOW, true, 10, PACKED>(c, sc, geom, entry, count, cap, wb, totals, cnt, n_hint); return; }
    switch (c->sstack) {
        case 9: launch_walk_s<SHADOW, false, 9, PACKED>(c, sc, geom, entry, count, cap, wb, totals, nullptr, n_hint); break;
        case 12: launch_walk_s<SHADOW, false, 12, PACKED>(c, sc, geom, entry, count, cap, wb, totals, nullptr, n_hint); break;
        default: launch_walk_s<SHADOW, false, 10, PACKED>(c, sc, geom, entry, count, cap, wb, totals, nullptr, n_hint); break;
    }
}

int walk(Context* c, const DScene& sc, bool shadow, const RayGeom* geom, const MeshEntry* entry, const uint32_t* count, uint32_t cap, const WalkBuffers& wb,
         FrameTotals* totals, TravCounters* cnt, uint32_t n_hint)
{
    if (sc.n_big == 0) return 0;
    {
        LaunchScope ls(c, shadow ? PROF_WALK_SHADOW : PROF_WALK_CLOSEST);
        if (shadow) {
            if (sc.walk_packed) launch_walk_p<true, true>(c, sc, geom, entry, count, cap, wb, totals, cnt, n_hint);
            else launch_walk_p<true, false>(c, sc, geom, entry, count, cap, wb, totals, cnt, n_hint);
        } else {
            if (sc.walk_packed) launch_walk_p<false, true>(c, sc, geom, entry, count, cap, wb, totals, cnt, n_hint);
            else launch_walk_p<false, false>(c, sc, geom, entry, count, cap, wb, totals, cnt, n_hint);
        }
    }
    {
        LaunchScope ls(c, PROF_FINISH);
        // a fraction of a percent of the rays: a small grid (it loops over whatever the list holds)
        const uint32_t blocks = stage_grid(c, std::max<uint32_t>(1, n_hint / 32), 8);
        if (c->finishWarp) {
            // one warp per listed ray, fetched through wb.fetch (zeroed together with the walk's cursor)
            const uint32_t wblocks = stage_grid(c, std::max<uint32_t>(1, n_hint / 32) * 8, 4);
            if (shadow) {
                if (cnt) k_finish_warp<true, true><<<wblocks, 128, 0, c->stream>>>(sc, geom, wb.cand, wb.ovf_list, wb.ovf_count, wb.fetch, cap, totals, cnt);
                else k_finish_warp<true, false><<<wblocks, 128, 0, c->stream>>>(sc, geom, wb.cand, wb.ovf_list, wb.ovf_count, wb.fetch, cap, totals, nullptr);
            } else {
                if (cnt) k_finish_warp<false, true><<<wblocks, 128, 0, c->stream>>>(sc, geom, wb.cand, wb.ovf_list, wb.ovf_count, wb.fetch, cap, totals, cnt);
                else k_finish_warp<false, false><<<wblocks, 128, 0, c->stream>>>(sc, geom, wb.cand, wb.ovf_list, wb.ovf_count, wb.fetch, cap, totals, nullptr);
            }
        } else if (shadow) {
            if (cnt) k_finish<true, true><<<blocks, 128, 0, c->stream>>>(sc, geom, wb.cand, wb.ovf_list, wb.ovf_count, cap, totals, cnt);
            else k_finish<true, false><<<blocks, 128, 0, c->stream>>>(sc, geom, wb.cand, wb.ovf_list, wb.ovf_count, cap, totals, nullptr);
        } else {
            if (cnt) k_finish<false, true><<<blocks, 128, 0, c->stream>>>(sc, geom, wb.cand, wb.ovf_list, wb.ovf_count, cap, totals, cnt);
            else k_finish<false, false><<<blocks, 128, 0, c->stream>>>(sc, geom, wb.cand, wb.ovf_list, wb.ovf_count, cap, totals, nullptr);
        }
    }
    return 2;
}

int shade(Context* c, const DScene& sc, const FrameParams& fp, const RayQueue& q, const CandRec* cand, uint32_t begin, uint32_t end, const Sinks& sinks,
          FrameTotals* totals, TravCounters* cnt)
{
    if (end <= begin) return 0;
    LaunchScope ls(c, PROF_SHADE);
    const uint32_t blocks = stage_grid(c, end - begin);
    // counting builds and scenes with CSG / heightfield / inline tree walks take the generic variant
    if (fp.gi) {
        if (cnt) k_shade<true, true, false><<<blocks, 128, 0, c->stream>>>(sc, fp, q, cand, begin, end, sinks, totals, cnt);
        else if (sc.simple_inline) k_shade<true, false, true><<<blocks, 128, 0, c->stream>>>(sc, fp, q, cand, begin, end, sinks, totals, nullptr);
        else k_shade<true, false, false><<<blocks, 128, 0, c->stream>>>(sc, fp, q, cand, begin, end, sinks, totals, nullptr);
    } else {
        if (cnt) k_shade<false, true, false><<<blocks, 128, 0, c->stream>>>(sc, fp, q, cand, begin, end, sinks, totals, cnt);
        else if (sc.simple_inline) k_shade<false, false, true><<<blocks, 128, 0, c->stream>>>(sc, fp, q, cand, begin, end, sinks, totals, nullptr);
        else k_shade<false, false, false><<<blocks, 128, 0, c->stream>>>(sc, fp, q, cand, begin, end, sinks, totals, nullptr);
    }
    return 1;
}

int resolve_shadow(Context* c, const DScene& sc, const ShadowQueue& q, const CandRec* cand, float* accum, uint8_t* visible, FrameTotals* totals,
                   TravCounters* cnt, uint32_t n_hint)
{
    LaunchScope ls(c, PROF_SHADOW_RESOLVE);
    const uint32_t blocks = stage_grid(c, n_hint);
    if (cnt) k_resolve_shadow<true><<<blocks, 128, 0, c->stream>>>(sc, q, cand, accum, visible, totals, cnt);
    else k_resolve_shadow<false><<<blocks, 128, 0, c->stream>>>(sc, q, cand, accum, visible, totals, nullptr);
    return 1;
}

int hit_records(Context* c, const DScene& sc, const RayQueue& q, const CandRec* cand, HitRec* hits, uint32_t n_hint)
{
    LaunchScope ls(c, PROF_OTHER);
    k_hit_records<<<stage_grid(c, n_hint), 128, 0, c->stream>>>(sc, q, cand, hits);
    return 1;
}

int aa_detect(Context* c, const float* vfb, int W, int H, int shard_index, int shard_count, uint32_t* list, uint32_t* n_out, uint8_t* mask)
{
    LaunchScope ls(c, PROF_OTHER);
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    k_aa_detect<<<g, b, 0, c->stream>>>(vfb, W, H, shard_index, shard_count, list, n_out, mask);
    return 1;
}
int scale_listed(Context* c, float* vfb, const uint32_t* list, const uint32_t* n, uint32_t cap, float mul)
{
    LaunchScope ls(c, PROF_OTHER);
    k_scale_listed<<<c->sms * 4, 256, 0, c->stream>>>(vfb, list, n, cap, mul);
    return 1;
}
int scale_all(Context* c, float* buf, size_t n, float mul)
{
    LaunchScope ls(c, PROF_OTHER);
    k_scale_all<<<c->sms * 8, 256, 0, c->stream>>>(buf, n, mul);
    return 1;
}
int add_into(Context* c, float* dst, const float* src, size_t n)
{
    LaunchScope ls(c, PROF_OTHER);
    k_add_into<<<c->sms * 8, 256, 0, c->stream>>>(dst, src, n);
    return 1;
}
int to_bmp_rows(Context* c, const float* rgb, int W, int H, int rowsz, const uint8_t* lut, uint8_t* out)
{
    LaunchScope ls(c, PROF_OTHER);
    cudaMemsetAsync(out, 0, (size_t)rowsz * H, c->stream);
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    k_to_bmp_rows<<<g, b, 0, c->stream>>>(rgb, W, H, rowsz, lut, out);
    return 1;
}
int to_exr_rows(Context* c, const float* rgb, int W, int H, uint16_t* out)
{
    LaunchScope ls(c, PROF_OTHER);
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    k_to_exr_rows<<<g, b, 0, c->stream>>>(rgb, W, H, out);
    return 1;
}
int stereo_mix(Context* c, float* out, const float* left, const float* right, size_t n_pixels)
{
    LaunchScope ls(c, PROF_OTHER);
    k_stereo_mix<<<c->sms * 8, 256, 0, c->stream>>>(out, left, right, n_pixels);
    return 1;
}
// ---- several devices ---------------------------------------------------------------------------------------------
bool enable_peer(Context* a, Context* b)
{
    if (a->device == b->device) return true;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, a->device, b->device) != cudaSuccess || !can) { cudaGetLastError(); return false; }
    if (cudaSetDevice(a->device) != cudaSuccess) return false;
    const cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return false; }
    cudaGetLastError();
    return true;
}

struct PeerPtrs { const float* p[HXR_MAX_PEERS]; };
// the partial frames of the other GPUs, read where they lie (peer memory over NVLink), summed into this GPU's and resolved
__global__ void __launch_bounds__(256) k_reduce_peers(float* __restrict__ dst, PeerPtrs srcs, int n_src, size_t n, float scale)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n4 = n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 a = reinterpret_cast<const float4*>(dst)[i];
        for (int k = 0; k < n_src; k++) {
            const float4 b = reinterpret_cast<const float4*>(srcs.p[k])[i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
        reinterpret_cast<float4*>(dst)[i] = a;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float a = dst[i];
        for (int k = 0; k < n_src; k++) a += srcs.p[k][i];
        dst[i] = a * scale;
    }
}
int reduce_peers(Context* c, float* dst, const float* const* srcs, int n_src, size_t n, float scale)
{
    if (n_src > HXR_MAX_PEERS) return 0;
    LaunchScope ls(c, PROF_OTHER);
    PeerPtrs pp;
    for (int k = 0; k < HXR_MAX_PEERS; k++) pp.p[k] = k < n_src ? srcs[k] : nullptr;
    k_reduce_peers<<<c->sms * 8, 256, 0, c->stream>>>(dst, pp, n_src, n, scale);
    return 1;
}

// NCCL, bound at run time: the library must load (and render on one GPU) where libnccl is absent
struct Comm {
    void* lib = nullptr;
    int n = 0;
    std::vector<void*> comms;  // ncclComm_t
    std::vector<Context*> ctxs;
    int (*groupStart)() = nullptr;
    int (*groupEnd)() = nullptr;
    int (*reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    int (*commDestroy)(void*) = nullptr;
    const char* (*errString)(int) = nullptr;
};
Comm* comm_create(Context* const* ctxs, int n, char* err, size_t errlen)
{
    auto fail = [&](const std::string& m) -> Comm* { snprintf(err, errlen, "%s", m.c_str()); return nullptr; };
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail("libnccl.so.2 not found");
    Comm* c = new Comm;
    c->lib = lib;
    c->n = n;
    auto initAll = (int (*)(void**, int, const int*))dlsym(lib, "ncclCommInitAll");
    c->groupStart = (int (*)())dlsym(lib, "ncclGroupStart");
    c->groupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
    c->reduce = (int (*)(const void*, void*, size_t, int, int, int, void*, cudaStream_t))dlsym(lib, "ncclReduce");
    c->commDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    c->errString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!initAll || !c->groupStart || !c->groupEnd || !c->reduce || !c->commDestroy) { delete c; return fail("libnccl lacks the expected symbols"); }
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) {
        devs[i] = ctxs[i]->device;
        for (int j = 0; j < i; j++)
            if (devs[j] == devs[i]) { delete c; return fail("NCCL needs distinct devices"); }
        c->ctxs.push_back(ctxs[i]);
    }
    c->comms.resize(n, nullptr);
    const int rc = initAll(c->comms.data(), n, devs.data());
    if (rc != 0) {
        const std::string m = std::string("ncclCommInitAll: ") + (c->errString ? c->errString(rc) : "failed");
        delete c;
        return fail(m);
    }
    return c;
}
void comm_destroy(Comm* c)
{
    if (!c) return;
    for (void* m : c->comms)
        if (m) c->commDestroy(m);
    delete c;
}
bool comm_reduce_sum(Comm* c, float* const* bufs, size_t n)
{
    // ncclFloat = 7, ncclSum = 0 (nccl.h); one group: every rank of this process enqueues its part on its own stream
    bool ok = c->groupStart() == 0;
    for (int i = 0; i < c->n && ok; i++) {
        cudaSetDevice(c->ctxs[i]->device);
        ok = c->reduce(bufs[i], bufs[i], n, 7, 0, 0, c->comms[i], c->ctxs[i]->stream) == 0;
    }
    ok = (c->groupEnd() == 0) && ok;
    return ok;
}

#include "kdbuild_kernels.inl"

}  // namespace dev
}  // namespace hxr
