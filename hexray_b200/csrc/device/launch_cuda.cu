// launch_cuda.cu — the sm_100a kernels of the render hot path and their launchers
// (implements device/launch.h; the only translation unit compiled by nvcc).
//
// Kernels (one per wavefront stage; per-item bodies live in pipeline.h / isect.h / shade.h):
//   k_gen_primary        primary-ray generator (pixel jitter / AA offsets / DOF lens / stereo eye)  -> ray queue
//   k_setup_closest      per ray, in double: every inline node (analytic primitives, CSG, heightfield, quads) in scene
//                        order; each mesh whose box the ray enters becomes a ready-to-walk float task
//   k_walk               the KD-tree walk, FP32 only: PERSISTENT warps, idle lanes refilled with __ballot_sync + one
//                        atomicAdd per warp + __shfl_sync; a bounded phase of block steps (32-byte block = two tree
//                        levels) alternates with a warp-cooperative phase that filters the triangles of all leaves held
//                        by the warp, 32 (ray, triangle) pairs at a time; undecided pairs go to a list (slots reserved 64
//                        per warp: one same-address atomic per append serialised in L2)
//   k_confirm_closest_a/b, k_confirm_shadow   the exact (double) triangle test on the listed pairs
//   k_finalize_closest   winner across inline nodes and walked meshes, IntersectionInfo, lights, environment, bump
//   k_shade<GI>          Whitted shader tree or path-tracing vertex: pushes child/shadow tasks,
//                        accumulates radiance with RED.ADD.F32
//   k_setup_shadow / k_walk<shadow> / k_confirm_shadow / k_accum_shadow   visible() in the same stages (any-hit walk)
//   k_aa_detect / k_scale_* / k_add_into / k_stereo_mix / k_to_bmp_rows   frame-buffer passes
// Grid sizing: the persistent walk launches (SM count x resident blocks/SM) blocks - 148 SMs on B200 - the per-item
// stage kernels are grid-stride loops over device-side counts, sized by a host-side upper bound of the count.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "launch.h"

namespace hxr {
namespace dev {

static int g_device = -1;
static int g_sms = 0;
static cudaStream_t g_stream = nullptr;
static std::string g_err;
static bool g_prof = false;
// walk-loop tunables (uniform kernel arguments; environment overrides for A/B runs: HXR_WALK_STEPS, HXR_REFILL_MIN, HXR_SSTACK,
// HXR_PAIR_CHUNK, HXR_NO_MAILBOX, HXR_BRANCHY_PUSH, HXR_WALK_CARVEOUT, HXR_WALK_BLOCKS_PER_SM; measured optima are the defaults)
#define HXR_PAIR_CHUNK 64 /* pair-list slots a warp of k_walk reserves per atomic (0: one atomic per append) */
static int g_walkSteps = 3, g_refillMin = 8, g_sstack = 10, g_useMail = 1, g_pairChunk = HXR_PAIR_CHUNK, g_bfPush = 1, g_walkCarveout = -1, g_walkBlocksPerSm = 0;
static uint64_t g_launches[PROF_NCAT];
static std::vector<cudaEvent_t> g_evPool;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_evPairs[PROF_NCAT];

static bool ck(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return true;
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
static bool use() { return g_device >= 0 && ck(cudaSetDevice(g_device), "cudaSetDevice"); }

bool init(int device, char* err, size_t errlen)
{
    auto fail = [&](const std::string& m) {
        snprintf(err, errlen, "%s", m.c_str());
        g_err = m;
        return false;
    };
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                    "); hexray_b200 has no CPU fallback");
    if (device < 0 || device >= n) return fail("CUDA device ordinal out of range");
    if (g_device >= 0 && g_device != device) return fail("this process is already bound to another device (one process per GPU)");
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed");
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return fail("cudaGetDeviceProperties failed");
    if (p.major < 10) return fail(std::string("device '") + p.name + "' is not Blackwell (sm_100a code only)");
    g_device = device;
    g_sms = p.multiProcessorCount;
    if (const char* e = getenv("HXR_WALK_STEPS")) g_walkSteps = std::max(1, atoi(e));
    if (const char* e = getenv("HXR_SSTACK")) g_sstack = atoi(e);
    if (getenv("HXR_NO_MAILBOX")) g_useMail = 0;
    if (getenv("HXR_BRANCHY_PUSH")) g_bfPush = 0;
    if (const char* e = getenv("HXR_WALK_CARVEOUT")) g_walkCarveout = std::min(100, std::max(0, atoi(e)));
    if (const char* e = getenv("HXR_WALK_BLOCKS_PER_SM")) g_walkBlocksPerSm = std::max(1, atoi(e));
    if (const char* e = getenv("HXR_PAIR_CHUNK")) g_pairChunk = std::min(1024, std::max(0, atoi(e)));
    if (const char* e = getenv("HXR_REFILL_MIN")) g_refillMin = std::min(32, std::max(1, atoi(e)));
    if (!g_stream && cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate failed");
    return true;
}
const char* backend_name() { return "cuda sm_100a"; }
const char* last_error() { return g_err.c_str(); }

void* alloc(size_t bytes)
{
    if (!use()) return nullptr;
    void* p = nullptr;
    if (!ck(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc")) return nullptr;
    return p;
}
void free_(void* p)
{
    if (p && use()) cudaFree(p);
}
bool upload(void* d, const void* s, size_t n)
{
    return use() && ck(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, g_stream), "H2D copy") && ck(cudaStreamSynchronize(g_stream), "H2D sync");
}
bool upload_pinned_async(void* d, const void* s, size_t n) { return use() && ck(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, g_stream), "H2D copy"); }
bool download(void* d, const void* s, size_t n)
{
    return use() && ck(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, g_stream), "D2H copy") && ck(cudaStreamSynchronize(g_stream), "D2H sync");
}
bool zero(void* p, size_t n) { return use() && ck(cudaMemsetAsync(p, 0, n, g_stream), "memset"); }
bool copy_d2d(void* d, const void* s, size_t n) { return use() && ck(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, g_stream), "D2D copy"); }
bool sync() { return use() && ck(cudaStreamSynchronize(g_stream), "stream sync"); }

struct Timer { cudaEvent_t a, b; };
Timer* timer_create()
{
    use();
    Timer* t = new Timer;
    cudaEventCreate(&t->a);
    cudaEventCreate(&t->b);
    return t;
}
void timer_destroy(Timer* t)
{
    if (!t) return;
    cudaEventDestroy(t->a);
    cudaEventDestroy(t->b);
    delete t;
}
void timer_start(Timer* t) { cudaEventRecord(t->a, g_stream); }
void timer_stop(Timer* t) { cudaEventRecord(t->b, g_stream); }
double timer_ms(Timer* t)
{
    float ms = 0;
    cudaEventSynchronize(t->b);
    cudaEventElapsedTime(&ms, t->a, t->b);
    return ms;
}

// ---- per-launch profiling -------------------------------------------------------------
void prof_enable(bool on) { g_prof = on; }
static cudaEvent_t ev_get()
{
    if (!g_evPool.empty()) { cudaEvent_t e = g_evPool.back(); g_evPool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_reset()
{
    for (int c = 0; c < PROF_NCAT; c++) {
        g_launches[c] = 0;
        for (auto& pr : g_evPairs[c]) { g_evPool.push_back(pr.first); g_evPool.push_back(pr.second); }
        g_evPairs[c].clear();
    }
}
void prof_collect(double ms[PROF_NCAT], uint64_t launches[PROF_NCAT])
{
    cudaStreamSynchronize(g_stream);
    for (int c = 0; c < PROF_NCAT; c++) {
        double s = 0;
        for (auto& pr : g_evPairs[c]) {
            float m = 0;
            cudaEventElapsedTime(&m, pr.first, pr.second);
            s += m;
        }
        ms[c] = s;
        launches[c] = g_launches[c];
    }
}
struct ProfScope {
    int cat;
    cudaEvent_t a = nullptr, b = nullptr;
    explicit ProfScope(int c) : cat(c)
    {
        use();
        g_launches[c]++;
        if (g_prof) { a = ev_get(); b = ev_get(); cudaEventRecord(a, g_stream); }
    }
    ~ProfScope()
    {
        if (g_prof) { cudaEventRecord(b, g_stream); g_evPairs[cat].emplace_back(a, b); }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) g_err = std::string("kernel launch: ") + cudaGetErrorString(e);
    }
};

// ---- kernels ----------------------------------------------------------------------------
#define HXR_TRACE_BLOCK 128
#define HXR_SHADE_BLOCK 128

__global__ void k_set_u32(uint32_t* p, uint32_t v) { *p = v; }

bool set_u32(uint32_t* p, uint32_t v)
{
    if (!use()) return false;
    if (v == 0) return ck(cudaMemsetAsync(p, 0, sizeof(uint32_t), g_stream), "memset");
    k_set_u32<<<1, 1, 0, g_stream>>>(p, v);
    return true;
}

__global__ void __launch_bounds__(128) k_gen_primary(DScene sc, FrameParams fp, const uint32_t* __restrict__ pixels, uint32_t first_pixel,
                                                     uint32_t n_items, uint32_t spp_pass, RayTask* __restrict__ q, uint32_t* q_count)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride) {
        const uint32_t pi = i / spp_pass;
        const uint32_t pixel = pixels ? pixels[pi] : first_pixel + pi;
        q[i] = gen_primary_item(sc, fp, pixel, fp.sample_base + (i % spp_pass) * fp.sample_stride);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *q_count = n_items;
}

#define HXR_WALK_BLOCK 128
// resident blocks per SM of the lean per-ray kernels (register caps 65536 / (128 * blocks)); A/B-ed on B200, see profiles/README.md
#ifndef HXR_SETUP_BLOCKS
#define HXR_SETUP_BLOCKS 6
#endif
#ifndef HXR_FIN_BLOCKS
#define HXR_FIN_BLOCKS 4
#endif
#ifndef HXR_SHADE_GI_BLOCKS
#define HXR_SHADE_GI_BLOCKS 5
#endif
#ifndef HXR_WALK_MIN_BLOCKS
#define HXR_WALK_MIN_BLOCKS 8
#endif
#ifndef HXR_REFILL_MIN
#define HXR_REFILL_MIN 8  /* refill as soon as this many lanes of a warp are idle */
#endif
#ifndef HXR_WALK_STEPS
#define HXR_WALK_STEPS 3  /* block steps per round of the walk loop */
#endif

template <bool COUNT, bool SIMPLE>
__global__ void __launch_bounds__(128, SIMPLE ? HXR_SETUP_BLOCKS : 1) k_setup_closest(DScene sc, const RayTask* __restrict__ q, const uint32_t* __restrict__ q_count, uint32_t cap,
                                                       TraceScratch ts, TravCounters* cnt)
{
    const uint32_t n = min(*q_count, cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    TravCounters local = {0, 0, 0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) setup_closest_item<COUNT, SIMPLE>(sc, task_ray(q[i]), i, ts, &local);
    if (COUNT && local.mesh_queries) atomicAdd(&cnt->mesh_queries, local.mesh_queries);
}

template <bool COUNT, bool SIMPLE>
__global__ void __launch_bounds__(128, SIMPLE ? HXR_FIN_BLOCKS : 2) k_finalize_closest(DScene sc, const RayTask* __restrict__ q, const uint32_t* __restrict__ q_count, uint32_t cap,
                                                          TraceScratch ts, HitRec* __restrict__ hits, TravCounters* cnt)
{
    const uint32_t n = min(*q_count, cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    TravCounters local = {0, 0, 0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        HitRec h;
        finalize_closest_item<COUNT, SIMPLE>(sc, task_ray(q[i]), i, ts, h, &local);
        hits[i] = h;
    }
}

template <bool COUNT, bool SIMPLE>
__global__ void __launch_bounds__(128, SIMPLE ? HXR_SETUP_BLOCKS : 1) k_setup_shadow(DScene sc, const ShadowTask* __restrict__ shadow, const uint32_t* __restrict__ count, uint32_t cap,
                                                      TraceScratch ts, TravCounters* cnt, unsigned long long* total)
{
    const uint32_t n = min(*count, cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    if (blockIdx.x == 0 && threadIdx.x == 0 && total) atomicAdd(total, (unsigned long long)n);
    TravCounters local = {0, 0, 0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) setup_shadow_item<COUNT, SIMPLE>(sc, shadow[i], i, ts, &local);
}

__global__ void __launch_bounds__(256) k_accum_shadow(const ShadowTask* __restrict__ shadow, const uint32_t* __restrict__ count, uint32_t cap, TraceScratch ts,
                                                      float* accum)
{
    const uint32_t n = min(*count, cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) accumulate_shadow_item(shadow[i], i, ts, accum);
}

// ---- the KD walk -------------------------------------------------------------------------------------
// One lane = one walk task (a ray inside one big mesh); lanes that run out of work are refilled from the task queue
// (__ballot_sync finds idle lanes, one atomicAdd per warp, __shfl_sync broadcast). Each round of the loop has two phases:
//   1. `walkSteps` block steps: every lane whose cursor is a tree block pops / steps (block_step: one 32-byte
//      fetch = two tree levels, up to four grandchildren front to back, the far ones pushed on the stack);
//   2. the leaves reached so far are filtered WARP-COOPERATIVELY: the (ray, triangle) pairs of all lanes' leaves are
//      dealt out evenly over the 32 lanes (prefix sum of the leaf sizes + binary search by shuffle), so a lane
//      with a long leaf does not hold up the others.
// Bounding phase 1 keeps lanes from idling while one ray of the warp descends a long path (the measured SIMD
// efficiency of an unbounded while-while loop on incoherent GI rays was 20 %).
//
// The kernel is FP32 only: conservative plane arithmetic (isect.h: block_step) and the conservative triangle
// filter (tri_filter). Pairs the filter cannot rule out are appended to a list for the exact double test
// (k_confirm_*); certain hits shorten the walk. No double arithmetic keeps the kernel at <= 64 registers
// (8 blocks = 1024 lanes per SM; A/B: 9 or 10 blocks at 55 / 48 registers and 5-7 blocks are all slower) and cuts the
// triangle bytes (48-byte TriF32, or 32-byte TriPacked on scenes whose triangles outgrow the L2, instead of 96 B).
//
// State per lane in shared memory: the float ray (24 B), the best-hit bound (4 B, lowered with atomicMin by whichever
// lane filters a certain hit) and the first HXR_SSTACK stack entries (12 B each); deeper entries overflow to local
// memory (rare: the stack is shallow for almost all rays).
#define HXR_POP 0x7FFFFFFFu /* cursor value: take the next entry from the stack */

template <int SSTACK>
struct WalkShared {
    float ray[9][HXR_WALK_BLOCK];  // rows 0-2 origin, 3-5 1/direction, 6-8 direction (a leaf child's "axis 3" reads the next row: finite, unused)
    uint32_t tb[HXR_WALK_BLOCK];  // bits of the (non-negative) float bound; 0 = shadow ray certainly blocked
    uint32_t mail[2][HXR_WALK_BLOCK];  // the last two triangles this task already put on the pair list (a triangle sits in several leaves)
    uint32_t stRef[SSTACK][HXR_WALK_BLOCK];
    float stMin[SSTACK][HXR_WALK_BLOCK];
    float stMax[SSTACK][HXR_WALK_BLOCK];
};

// SSTACK = stack entries kept in shared memory: every entry costs 1.5 KB of the SM's 256 KB L1/shared array per block
template <bool SHADOW, bool COUNT, int SSTACK, bool PACKED>
__global__ void __launch_bounds__(HXR_WALK_BLOCK, HXR_WALK_MIN_BLOCKS) k_walk(DScene sc, TraceScratch ts, TravCounters* cnt, int walkSteps,
                                                                              int refillMin, int useMail, int pairChunk, int branchFreePush)
{
    __shared__ WalkShared<SSTACK> sh;
    constexpr int HXR_SSTACK = SSTACK;
    const unsigned FULL = 0xffffffffu;
    const uint32_t n = min(*ts.task_count, ts.task_cap);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warpBase = tid & ~31u;
    uint32_t ovRef[HXR_KD_STACK - HXR_SSTACK];
    float ovMin[HXR_KD_STACK - HXR_SSTACK], ovMax[HXR_KD_STACK - HXR_SSTACK];
    bool active = false, drained = false;
    struct SharedRay {  // o(axis) / inv(axis) straight from this lane's shared-memory column: no selects, no registers
        const float* col;
        uint32_t par;
        __device__ __forceinline__ float o(uint32_t axis) const { return col[axis * HXR_WALK_BLOCK]; }
        __device__ __forceinline__ float inv(uint32_t axis) const { return col[(3 + axis) * HXR_WALK_BLOCK]; }
    } wr;
    wr.col = &sh.ray[0][tid];
    wr.par = 0;
    for (int r = 0; r < 9; r++) sh.ray[r][tid] = 0.0f;
    const KdBlock* blocks = nullptr;  // the current mesh
    const uint32_t* leafTris = nullptr;
    const void* tris = nullptr;  // TriPacked (32 B) or TriF32 (48 B) records of the current mesh
    bool backface = false;
    int meshIdx = -1;
    uint32_t cur = HXR_POP, leafCnt = 0;
    float tmin = 0, tmax = 0, tbest = 0, err = 0, occ = 0;
    int sp = 0;
    uint32_t taskIdx = 0, taskRay = 0;
    uint32_t pkBase = 0, pkLeft = 0;  // this warp's reserved slots of the pair list (warp-uniform)
    TravCounters local = {0, 0, 0, 0};

    auto push = [&](const WalkEnt& e) {
        if (sp < HXR_SSTACK) {
            sh.stRef[sp][tid] = e.ref; sh.stMin[sp][tid] = e.lo; sh.stMax[sp][tid] = e.hi;
        } else if (sp < HXR_KD_STACK) {
            ovRef[sp - HXR_SSTACK] = e.ref; ovMin[sp - HXR_SSTACK] = e.lo; ovMax[sp - HXR_SSTACK] = e.hi;
        } else {
            return;  // unreachable: the build caps the depth at HXR_KD_MAX_DEPTH (<= 1.5 pushes per level)
        }
        sp++;
    };

    for (;;) {
        // ---- refill: idle lanes take new tasks
        const unsigned idle = __ballot_sync(FULL, !active);
        if (!drained && (idle == FULL || __popc(idle) >= refillMin)) {
            const int c = __popc(idle);
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(ts.head, (uint32_t)c);
            base = __shfl_sync(FULL, base, leader);
            if (base + (uint32_t)c >= n) drained = true;
            if (!active) {
                const uint32_t k = base + __popc(idle & ((1u << lane) - 1u));
                if (k < n) {
                    const uint4* tp = reinterpret_cast<const uint4*>(ts.tasks + k);
                    const uint4 w0 = __ldg(tp), w1 = __ldg(tp + 1), w2 = __ldg(tp + 2), w3 = __ldg(tp + 3);
                    taskRay = w0.x;
                    if (!(SHADOW && ts.occluded[taskRay])) {
                        taskIdx = k;
                        const float ox = __uint_as_float(w0.z), oy = __uint_as_float(w0.w), oz = __uint_as_float(w1.x);
                        const float dx = __uint_as_float(w1.y), dy = __uint_as_float(w1.z), dz = __uint_as_float(w1.w);
                        const WalkRay w = walk_ray_f(ox, oy, oz, dx, dy, dz);
                        wr.par = w.par;
                        sh.ray[0][tid] = ox; sh.ray[1][tid] = oy; sh.ray[2][tid] = oz;
                        sh.ray[3][tid] = w.ix; sh.ray[4][tid] = w.iy; sh.ray[5][tid] = w.iz;
                        sh.ray[6][tid] = dx; sh.ray[7][tid] = dy; sh.ray[8][tid] = dz;
                        tmin = __uint_as_float(w2.x);
                        tmax = __uint_as_float(w2.y);
                        tbest = __uint_as_float(w2.z);
                        occ = __uint_as_float(w2.w);
                        err = __uint_as_float(w3.x);
                        meshIdx = (int)w3.y;
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
                }
            }
        }
        if (__ballot_sync(FULL, active) == 0) {
            if (drained) break;
            continue;
        }
        // ---- phase 1: a few block steps for every lane whose cursor is not a leaf
#pragma unroll 1
        for (int it = 0; it < walkSteps; it++) {
            const bool stepping = active && !(cur >> 31);
            if (stepping && cur == HXR_POP) {
                if (sp == 0) {
                    active = false;  // nothing left: this task is done (its hits are in the pair list)
                } else {
                    sp--;
                    WalkEnt e;
                    if (sp < HXR_SSTACK) { e.ref = sh.stRef[sp][tid]; e.lo = sh.stMin[sp][tid]; e.hi = sh.stMax[sp][tid]; }
                    else { e.ref = ovRef[sp - HXR_SSTACK]; e.lo = ovMin[sp - HXR_SSTACK]; e.hi = ovMax[sp - HXR_SSTACK]; }
                    if (e.lo <= tbest) { cur = e.ref; tmin = e.lo; tmax = e.hi; }  // else it cannot hold a closer hit: keep popping
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
                        if (sp > HXR_KD_STACK) sp = HXR_KD_STACK;  // unreachable: the build caps the depth (see push)
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
            const uint32_t oTask = __shfl_sync(FULL, taskIdx, o);
            const float oOcc = SHADOW ? __shfl_sync(FULL, occ, o) : 0.0f;
            bool emit = false;
            uint32_t ti = 0;
            if (pair < total) {
                const uint32_t* lt = leafTris;
                const void* tt = tris;
                bool bf = backface;
                if (oMesh != meshIdx) {
                    const DMesh& M = sc.meshes[oMesh];
                    lt = M.leaf_tris; tt = PACKED ? (const void*)M.tri_pk : (const void*)M.tri_f32; bf = M.backface != 0;
                }
                ti = __ldg(lt + oFirst + (pair - oExcl));
                const unsigned ot = warpBase | (unsigned)o;
                const float oBest = __uint_as_float(sh.tb[ot]);  // the freshest bound (other lanes may have lowered it this round)
                float ghi = 0;
                int cls = HXR_TF_MISS;
                if (ti != sh.mail[0][ot] && ti != sh.mail[1][ot]) {  // not already on the list from a neighbouring leaf
                    if (PACKED) cls = tri_filter_packed(static_cast<const TriPacked*>(tt) + ti, bf, sh.ray[0][ot], sh.ray[1][ot], sh.ray[2][ot], sh.ray[6][ot], sh.ray[7][ot], sh.ray[8][ot], oErr, oBest, ghi);
                    else cls = tri_filter(static_cast<const TriF32*>(tt) + ti, bf, sh.ray[0][ot], sh.ray[1][ot], sh.ray[2][ot], sh.ray[6][ot], sh.ray[7][ot], sh.ray[8][ot], oErr, oBest, ghi);
                }
                if (cls == HXR_TF_CERTAIN) {
                    if (SHADOW && ghi < oOcc) atomicMin(&sh.tb[ot], 0u);  // certainly blocked: no exact test needed
                    else { atomicMin(&sh.tb[ot], __float_as_uint(ghi)); emit = true; }
                } else if (cls == HXR_TF_MAYBE) {
                    emit = true;
                }
            }
            // append the surviving pairs to the confirmation list. Every warp of the GPU appends all the time: one atomic per
            // append on the single list counter serialises in L2 (measured: 20 % of this kernel's stall samples sat here), so
            // a warp reserves HXR_PAIR_CHUNK slots at a time and hands them out itself; slots it does not use are marked
            // invalid (task = HXR_PAIR_NONE) and skipped by the confirmation kernels
            const unsigned em = __ballot_sync(FULL, emit);
            if (em) {
                const uint32_t need = (uint32_t)__popc(em);
                if (need > pkLeft) {
                    if (lane < pkLeft && pkBase + lane < ts.pair_cap) ts.pairs[pkBase + lane].task = HXR_PAIR_NONE;  // pkLeft < need <= 32
                    uint32_t nb = 0;
                    const uint32_t take = pairChunk ? (uint32_t)pairChunk : need;  // 0: one atomic per append (A/B)
                    if (lane == 0) nb = atomicAdd(ts.pair_count, take);
                    pkBase = __shfl_sync(FULL, nb, 0);
                    pkLeft = take;
                }
                const uint32_t pb = pkBase;
                pkBase += need;
                pkLeft -= need;
                if (emit) {
                    const unsigned ot = warpBase | (unsigned)o;
                    if (useMail) {
                        sh.mail[1][ot] = sh.mail[0][ot];  // (lanes emitting for the same owner race here: any of their triangles is a valid entry)
                        sh.mail[0][ot] = ti;
                    }
                    const uint32_t k = pb + __popc(em & ((1u << lane) - 1u));
                    if (k < ts.pair_cap) { PairRec pr; pr.task = oTask; pr.tri = ti; ts.pairs[k] = pr; }
                    else atomicExch(ts.overflow, 1u);
                }
            }
        }
        __syncwarp();
        if (hasLeaf) {
            if (COUNT) { local.kd_leaves++; local.tri_tests += leafCnt; }
            cur = HXR_POP;
            const uint32_t tb = sh.tb[tid];
            tbest = __uint_as_float(tb);
            if (SHADOW && tb == 0u) {
                ts.occluded[taskRay] = 1;
                active = false;
            }
        }
    }
    for (uint32_t i = lane; i < pkLeft; i += 32u)  // the unused tail of this warp's last chunk
        if (pkBase + i < ts.pair_cap) ts.pairs[pkBase + i].task = HXR_PAIR_NONE;
    if (COUNT) {
        atomicAdd(&cnt->kd_inner, local.kd_inner);
        atomicAdd(&cnt->kd_leaves, local.kd_leaves);
        atomicAdd(&cnt->tri_tests, local.tri_tests);
        atomicAdd(&cnt->mesh_queries, local.mesh_queries);
    }
}

// the exact test of the pairs the walk left undecided (one thread per pair; the count lives on the device)
__global__ void __launch_bounds__(128) k_confirm_closest_a(DScene sc, const RayTask* __restrict__ q, TraceScratch ts)
{
    const uint32_t n = min(*ts.pair_count, ts.pair_cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) confirm_closest_a_item(sc, q, i, ts);
}
__global__ void __launch_bounds__(256) k_confirm_closest_b(DScene sc, TraceScratch ts)
{
    const uint32_t n = min(*ts.pair_count, ts.pair_cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) confirm_closest_b_item(sc, i, ts);
}
__global__ void __launch_bounds__(128) k_confirm_shadow(DScene sc, const ShadowTask* __restrict__ shadows, TraceScratch ts)
{
    const uint32_t n = min(*ts.pair_count, ts.pair_cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) confirm_shadow_item(sc, shadows, i, ts);
}

// GI (one path vertex) and Whitted (shader tree with its stacks) are separate compilations: the path-tracing vertex
// needs far fewer registers than the tree walk, and occupancy is what hides this kernel's gather latency
template <bool GI>
__global__ void __launch_bounds__(HXR_SHADE_BLOCK, GI ? HXR_SHADE_GI_BLOCKS : 3) k_shade(DScene sc, FrameParams fp, const RayTask* __restrict__ q,
                                                                       const uint32_t* __restrict__ q_count, const HitRec* __restrict__ hits,
                                                                       uint32_t begin, uint32_t end, Sinks sinks)
{
    const uint32_t e = min(end, *q_count);
    const uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e) return;
    if (GI) shade_gi_item(sc, fp, q[i], hits[i], sinks);
    else shade_whitted_item(sc, fp, q[i], hits[i], sinks);
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
// ---- launchers --------------------------------------------------------------------------
template <class K> static int persistent_grid(K kernel, int block)
{
    int perSm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, block, 0) != cudaSuccess || perSm < 1) perSm = 1;
    return g_sms * perSm;
}

int gen_primary(const DScene& sc, const FrameParams& fp, const uint32_t* pixels, uint32_t first_pixel, uint32_t n_items,
                uint32_t spp_pass, RayTask* q, uint32_t* q_count)
{
    ProfScope ps(PROF_OTHER);
    const uint32_t blocks = n_items ? (uint32_t)std::min<uint64_t>(((uint64_t)n_items + 127) / 128, (uint64_t)g_sms * 32) : 1;
    k_gen_primary<<<blocks, 128, 0, g_stream>>>(sc, fp, pixels, first_pixel, n_items, spp_pass, q, q_count);
    return 1;
}

// per-item stage kernels are grid-stride loops over a count that only the device knows
static uint32_t stage_grid(uint32_t n) { return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)n + 127) / 128, (uint64_t)g_sms * 16)); }

template <class K> static int walk_grid(K kernel)
{
    int perSm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, HXR_WALK_BLOCK, 0) != cudaSuccess || perSm < 1) perSm = 1;
    return g_sms * perSm;
}

// max_tasks: host-side upper bound of the task count (a small wave gets a small grid: one lane per task at most)
template <bool SHADOW, bool COUNT, int SSTACK, bool PACKED> static void launch_walk_s(const DScene& sc, const TraceScratch& ts, TravCounters* cnt, uint64_t max_tasks)
{
    static int full = 0;
    if (!full) {
        // fewer resident blocks than fit + a smaller shared-memory carve-out leave the L1 more room for the top of the tree
        if (g_walkCarveout >= 0) cudaFuncSetAttribute(k_walk<SHADOW, COUNT, SSTACK, PACKED>, cudaFuncAttributePreferredSharedMemoryCarveout, g_walkCarveout);
        full = walk_grid(k_walk<SHADOW, COUNT, SSTACK, PACKED>);
        if (g_walkBlocksPerSm > 0) full = std::min(full, g_sms * g_walkBlocksPerSm);
    }
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)full, (max_tasks + HXR_WALK_BLOCK - 1) / HXR_WALK_BLOCK));
    k_walk<SHADOW, COUNT, SSTACK, PACKED><<<grid, HXR_WALK_BLOCK, 0, g_stream>>>(sc, ts, cnt, g_walkSteps, g_refillMin, g_useMail, g_pairChunk, g_bfPush);
}
template <bool SHADOW, bool PACKED> static void launch_walk_p(const DScene& sc, const TraceScratch& ts, TravCounters* cnt, uint64_t max_tasks)
{
    if (cnt) { launch_walk_s<SHADOW, true, 10, PACKED>(sc, ts, cnt, max_tasks); return; }
    switch (g_sstack) {
        case 9: launch_walk_s<SHADOW, false, 9, PACKED>(sc, ts, nullptr, max_tasks); break;
        case 12: launch_walk_s<SHADOW, false, 12, PACKED>(sc, ts, nullptr, max_tasks); break;
        default: launch_walk_s<SHADOW, false, 10, PACKED>(sc, ts, nullptr, max_tasks); break;
    }
}
template <bool SHADOW> static void launch_walk(const DScene& sc, const TraceScratch& ts, TravCounters* cnt, uint64_t max_tasks)
{
    if (sc.walk_packed) launch_walk_p<SHADOW, true>(sc, ts, cnt, max_tasks);
    else launch_walk_p<SHADOW, false>(sc, ts, cnt, max_tasks);
}

int trace_closest(const DScene& sc, const RayTask* q, const uint32_t* q_count, uint32_t q_cap, HitRec* hits, const TraceScratch& ts,
                  TravCounters* cnt, uint32_t n_hint)
{
    ProfScope ps(PROF_TRACE_CLOSEST);
    cudaMemsetAsync(ts.task_count, 0, sizeof(uint32_t), g_stream);
    cudaMemsetAsync(ts.head, 0, sizeof(uint32_t), g_stream);
    cudaMemsetAsync(ts.pair_count, 0, sizeof(uint32_t), g_stream);
    const uint32_t nb = stage_grid(std::min(q_cap, n_hint));
    const uint64_t maxTasks = (uint64_t)std::min(q_cap, n_hint) * (uint64_t)std::max(1, sc.n_big);
    const uint32_t nbPairs = stage_grid((uint32_t)std::min<uint64_t>(ts.pair_cap, maxTasks * 4));  // a ray rarely leaves more than a few pairs
    int launches = 2;
    // counting builds and scenes with CSG / heightfield / inline tree walks take the generic variant
    const bool simple = sc.simple_inline && !cnt;
    if (cnt) k_setup_closest<true, false><<<nb, 128, 0, g_stream>>>(sc, q, q_count, q_cap, ts, cnt);
    else if (simple) k_setup_closest<false, true><<<nb, 128, 0, g_stream>>>(sc, q, q_count, q_cap, ts, nullptr);
    else k_setup_closest<false, false><<<nb, 128, 0, g_stream>>>(sc, q, q_count, q_cap, ts, nullptr);
    if (sc.n_big) {
        {
            ProfScope pw(PROF_WALK);
            launch_walk<false>(sc, ts, cnt, maxTasks);
        }
        k_confirm_closest_a<<<nbPairs, 128, 0, g_stream>>>(sc, q, ts);
        k_confirm_closest_b<<<nbPairs, 256, 0, g_stream>>>(sc, ts);
        launches += 3;
    }
    if (cnt) k_finalize_closest<true, false><<<nb, 128, 0, g_stream>>>(sc, q, q_count, q_cap, ts, hits, cnt);
    else if (simple) k_finalize_closest<false, true><<<nb, 128, 0, g_stream>>>(sc, q, q_count, q_cap, ts, hits, nullptr);
    else k_finalize_closest<false, false><<<nb, 128, 0, g_stream>>>(sc, q, q_count, q_cap, ts, hits, nullptr);
    g_launches[PROF_TRACE_CLOSEST] += launches - 1;
    return launches;
}

int shade(const DScene& sc, const FrameParams& fp, const RayTask* q, const uint32_t* q_count, const HitRec* hits, uint32_t begin,
          uint32_t end, const Sinks& sinks)
{
    if (end <= begin) return 0;
    ProfScope ps(PROF_SHADE);
    const uint32_t blocks = (end - begin + HXR_SHADE_BLOCK - 1) / HXR_SHADE_BLOCK;
    if (fp.gi) k_shade<true><<<blocks, HXR_SHADE_BLOCK, 0, g_stream>>>(sc, fp, q, q_count, hits, begin, end, sinks);
    else k_shade<false><<<blocks, HXR_SHADE_BLOCK, 0, g_stream>>>(sc, fp, q, q_count, hits, begin, end, sinks);
    return 1;
}

int trace_shadow(const DScene& sc, const ShadowTask* shadow, const uint32_t* count, uint32_t cap, float* accum, const TraceScratch& ts,
                 TravCounters* cnt, unsigned long long* total, uint32_t n_hint)
{
    ProfScope ps(PROF_TRACE_SHADOW);
    cudaMemsetAsync(ts.task_count, 0, sizeof(uint32_t), g_stream);
    cudaMemsetAsync(ts.head, 0, sizeof(uint32_t), g_stream);
    cudaMemsetAsync(ts.pair_count, 0, sizeof(uint32_t), g_stream);
    const uint32_t nb = stage_grid(std::min(cap, n_hint));
    const uint64_t maxTasks = (uint64_t)std::min(cap, n_hint) * (uint64_t)std::max(1, sc.n_big);
    const uint32_t nbPairs = stage_grid((uint32_t)std::min<uint64_t>(ts.pair_cap, maxTasks * 4));
    int launches = 1;
    if (cnt) k_setup_shadow<true, false><<<nb, 128, 0, g_stream>>>(sc, shadow, count, cap, ts, cnt, total);
    else if (sc.simple_inline) k_setup_shadow<false, true><<<nb, 128, 0, g_stream>>>(sc, shadow, count, cap, ts, nullptr, total);
    else k_setup_shadow<false, false><<<nb, 128, 0, g_stream>>>(sc, shadow, count, cap, ts, nullptr, total);
    if (sc.n_big) {
        {
            ProfScope pw(PROF_WALK);
            launch_walk<true>(sc, ts, cnt, maxTasks);
        }
        k_confirm_shadow<<<nbPairs, 128, 0, g_stream>>>(sc, shadow, ts);
        launches += 2;
    }
    if (accum) { k_accum_shadow<<<nb, 256, 0, g_stream>>>(shadow, count, cap, ts, accum); launches++; }
    g_launches[PROF_TRACE_SHADOW] += launches - 1;
    return launches;
}

int aa_detect(const float* vfb, int W, int H, int shard_index, int shard_count, uint32_t* list, uint32_t* n_out, uint8_t* mask)
{
    ProfScope ps(PROF_OTHER);
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    k_aa_detect<<<g, b, 0, g_stream>>>(vfb, W, H, shard_index, shard_count, list, n_out, mask);
    return 1;
}
int scale_listed(float* vfb, const uint32_t* list, const uint32_t* n, uint32_t cap, float mul)
{
    ProfScope ps(PROF_OTHER);
    k_scale_listed<<<g_sms * 4, 256, 0, g_stream>>>(vfb, list, n, cap, mul);
    return 1;
}
int scale_all(float* buf, size_t n, float mul)
{
    ProfScope ps(PROF_OTHER);
    k_scale_all<<<g_sms * 8, 256, 0, g_stream>>>(buf, n, mul);
    return 1;
}
int add_into(float* dst, const float* src, size_t n)
{
    ProfScope ps(PROF_OTHER);
    k_add_into<<<g_sms * 8, 256, 0, g_stream>>>(dst, src, n);
    return 1;
}
int to_bmp_rows(const float* rgb, int W, int H, int rowsz, const uint8_t* lut, uint8_t* out)
{
    ProfScope ps(PROF_OTHER);
    cudaMemsetAsync(out, 0, (size_t)rowsz * H, g_stream);
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    k_to_bmp_rows<<<g, b, 0, g_stream>>>(rgb, W, H, rowsz, lut, out);
    return 1;
}
int stereo_mix(float* out, const float* left, const float* right, size_t n_pixels)
{
    ProfScope ps(PROF_OTHER);
    k_stereo_mix<<<g_sms * 8, 256, 0, g_stream>>>(out, left, right, n_pixels);
    return 1;
}
}  // namespace dev
}  // namespace hxr
