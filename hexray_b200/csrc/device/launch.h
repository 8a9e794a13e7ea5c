// launch.h — the narrow interface between the host-side frame driver (renderer.cpp) and the
// device. The product implements it in launch_cuda.cu (sm_100a kernels, CUDA streams/events).
// tests/emu/launch_emu.cpp implements the same interface as plain host loops over the SAME
// per-item functions (pipeline.h) — test infrastructure for the CPU tier, never shipped.
//
// Everything goes through a Context: one per GPU, holding the device ordinal, its stream, its error state and its
// profiling events. A process may hold any number of contexts on any number of devices (the multi-GPU renderer of
// multi.cpp drives one per device from one host thread each); there is no process-wide device state.
#pragma once
#include "pipeline.h"
#include "kdbuild.h"

namespace hxr {
namespace dev {

struct Context;

// ---- device lifetime / memory -------------------------------------------------------
// nullptr (and err filled) if the device is unusable (product: no CUDA device => hard failure, there is no CPU fallback)
Context* create(int device, char* err, size_t errlen);
void destroy(Context*);
int device_count();                    // usable devices in this process (0 if none)
int device_of(const Context*);
void* stream_of(const Context*);       // cudaStream_t of the context (nullptr in the emulation)
const char* backend_name();
void* alloc(Context*, size_t bytes);   // nullptr on failure
void free_(Context*, void* p);
void* alloc_pinned(Context*, size_t bytes);  // page-locked host memory (plain malloc in the emulation)
void free_pinned(Context*, void* p);
bool upload(Context*, void* dst, const void* src, size_t bytes);
bool download(Context*, void* dst, const void* src, size_t bytes);       // blocking
bool download_async(Context*, void* dst, const void* src, size_t bytes); // stream ordered; dst should be pinned
bool zero(Context*, void* p, size_t bytes);
bool copy_d2d(Context*, void* dst, const void* src, size_t bytes);
bool sync(Context*);  // both lanes
// Two stream-ordered lanes per context: lane 0 (main) and lane 1 (aux). Every launch, memset and copy below goes to the lane
// selected by lane(); fork makes the aux lane wait for everything queued on the main lane so far, join makes the main lane
// wait for the aux lane. The frame driver runs the shadow chain of one bounce on the aux lane while the closest-hit chain of
// the next bounce runs on the main lane (they share no buffers): small frames, whose launches do not fill the GPU, overlap.
void lane(Context*, int which);
void fork(Context*);
void join(Context*);
// text of the first failed call or kernel launch since the last clear_error ("" if none): launches are asynchronous, so
// the frame driver checks this once per frame instead of after every launch
const char* last_error(const Context*);
bool failed(const Context*);
void clear_error(Context*);

// timers on the context's stream (CUDA events in the product)
struct Timer;
Timer* timer_create(Context*);
void timer_destroy(Context*, Timer*);
void timer_start(Context*, Timer*);
void timer_stop(Context*, Timer*);
double timer_ms(Context*, Timer*);     // blocks until the stop event has happened

// per-category device time of the launches below (CUDA events around every launch when enabled)
enum ProfCat { PROF_WALK_CLOSEST = 0, PROF_WALK_SHADOW = 1, PROF_SHADE = 2, PROF_SHADOW_RESOLVE = 3, PROF_GEN = 4, PROF_OTHER = 5, PROF_SETUP = 6, PROF_FINISH = 7, PROF_NCAT = 8 };
void prof_enable(Context*, bool on);
void prof_reset(Context*);
void prof_collect(Context*, double ms[PROF_NCAT], uint64_t launches[PROF_NCAT]);  // blocks; launches are counted even when disabled

// device-side totals of one frame (64-bit, accumulated by the kernels)
struct FrameTotals {
    unsigned long long rays_closest;   // closest-hit queries past the depth guard
    unsigned long long rays_shadow;    // visible() queries
    unsigned long long cand_overflow;  // rays whose candidate record filled up (finished by the second, exact-on-the-spot walk)
    unsigned long long pad;
};

// ---- kernels -------------------------------------------------------------------------
// Every launcher returns the number of kernel launches it issued. Counts live on the device; n_hint is a host-side upper
// bound of the count that only sizes the grid (a small wave does not pay for a full-size launch).

// primary rays for `n_items` (pixel, sample) pairs, written to q[0 .. n_items) with their inline part decided; *q.count = n_items
//   item i -> pixel = pixels ? pixels[i / spp_pass] : first_pixel + i / spp_pass,
//             sample = fp.sample_base + (i % spp_pass) * fp.sample_stride
// pixels_count (device, may be null): when given, only the first *pixels_count * spp_pass items exist
int gen_primary(Context*, const DScene& sc, const FrameParams& fp, const uint32_t* pixels, const uint32_t* pixels_count, uint32_t first_pixel,
                uint32_t n_items, uint32_t spp_pass, const RayQueue& q, CandRec* cand);

// The inline part of queued rays, in place: geom[i].limit / .pre for closest-hit rays (every inline node in scene order: analytic
// primitives, CSG, heightfields, quads, in double), geom[i].pre = -2 for shadow rays an inline node or a light blocks; and the
// rays' slot-0 entry records for the walk (q.entry). Rays with nothing to walk get a dead entry and their candidate record
// (cand[i]) here. accum (shadow rays, may be null): shadow rays that nothing can block add their colour right away.
int setup_closest(Context*, const DScene& sc, const RayQueue& q, CandRec* cand, TravCounters* cnt, uint32_t n_hint);
int setup_shadow(Context*, const DScene& sc, const ShadowQueue& q, CandRec* cand, float* accum, TravCounters* cnt, uint32_t n_hint);

// the KD walk of the big meshes for the rays geom[0 .. *count) with slot-0 entries entry[]: one candidate record per live ray
// (wb.head zeroed by the caller).
// A ray whose record fills up stops there and is appended to ovf_list (capacity cap, count *ovf_count, zeroed by the caller);
// a second, small launch finishes those rays (finish_overflowed_ray) and replaces their records. totals->cand_overflow += their number.
struct WalkBuffers {
    CandRec* cand;
    uint32_t* head;       // work-fetch cursor of the persistent kernel
    OverflowEntry* ovf_list;
    uint32_t* ovf_count;
    uint32_t* fetch;      // work-fetch cursor of the kernel that finishes the listed rays (zero when the walk is queued)
};
int walk(Context*, const DScene& sc, bool shadow, const RayGeom* geom, const MeshEntry* entry, const uint32_t* count, uint32_t cap, const WalkBuffers& wb,
         FrameTotals* totals, TravCounters* cnt, uint32_t n_hint);

// shade q[begin .. min(end, *q.count)): exact test of the candidates, winner, shading; pushes child rays and shadow rays
// (raw: their inline part is decided by setup_closest / setup_shadow).
// totals->rays_closest += the rays shaded, totals->rays_shadow += the visible() queries issued
int shade(Context*, const DScene& sc, const FrameParams& fp, const RayQueue& q, const CandRec* cand, uint32_t begin, uint32_t end, const Sinks& sinks,
          FrameTotals* totals, TravCounters* cnt);

// shadow rays after their walk: exact test of undecided pairs, then accum[pixel] += colour of every unblocked ray;
// visible (may be null): visible[i] = 1 / 0 (test hook)
int resolve_shadow(Context*, const DScene& sc, const ShadowQueue& q, const CandRec* cand, float* accum, uint8_t* visible, FrameTotals* totals,
                   TravCounters* cnt, uint32_t n_hint);

// raycast() records for q[0 .. *q.count) (test hook)
int hit_records(Context*, const DScene& sc, const RayQueue& q, const CandRec* cand, HitRec* hits, uint32_t n_hint);

// ---- device KD-tree build: one level of the whole tree per round of these passes (kdbuild.h; driver: csrc/kdbuild.cpp).
// All pointers are device memory of the context. tb = triangle bounds, SoA [6][nTris].
int kd_bounds(Context*, const double* vertices, const int32_t* triV, uint32_t nTris, double* tb);
int kd_iota(Context*, uint32_t* refTri, uint32_t* refNode, uint32_t n);  // refTri[i] = i, refNode[i] = 0
// start / end histograms ([nNodes][3][2][HXR_KDB_BINS], zeroed by the caller) of the nodes with more than P.binnedAbove references
int kd_bin(Context*, const kdb::Params& P, const kdb::NodeWork* work, const uint32_t* refTri, const uint32_t* refNode, uint32_t nRefs, const double* tb,
           uint32_t nTris, uint32_t* hist);
// hist may be null when no node of the level is binned
int kd_choose(Context*, const kdb::Params& P, const kdb::NodeWork* work, uint32_t nNodes, int depth, const uint32_t* hist, const uint32_t* refTri,
              const double* tb, uint32_t nTris, kdb::Decision* dec);
int kd_classify(Context*, const uint32_t* refTri, const uint32_t* refNode, uint32_t nRefs, const kdb::Decision* dec, const double* tb, uint32_t nTris,
                uint32_t* flagL, uint32_t* flagR);  // also flagL[nRefs] = flagR[nRefs] = 0
int scan_u32(Context*, uint32_t* data, uint32_t n);  // exclusive prefix sum, in place
// childRefs / isSplit / leafRefs [nNodes + 1] (last = 0); *levelMax = the largest child (zeroed by the caller)
int kd_plan(Context*, const kdb::NodeWork* work, uint32_t nNodes, kdb::Decision* dec, const uint32_t* scanL, const uint32_t* scanR, uint32_t* childRefs,
            uint32_t* isSplit, uint32_t* leafRefs, uint32_t* levelMax);
int kd_emit(Context*, const kdb::NodeWork* work, uint32_t nNodes, const kdb::Decision* dec, const uint32_t* childRefs, const uint32_t* isSplit,
            const uint32_t* leafRefs, uint32_t outCount, uint32_t leafBase, kdb::OutNode* out, kdb::NodeWork* next);
int kd_scatter(Context*, const uint32_t* refTri, const uint32_t* refNode, uint32_t nRefs, const kdb::NodeWork* work, const kdb::Decision* dec,
               const uint32_t* scanL, const uint32_t* scanR, const uint32_t* childRefs, const uint32_t* isSplit, const uint32_t* leafRefs,
               uint32_t leafBase, uint32_t* nextTri, uint32_t* nextNode, uint32_t* leafOut);

// needsAA flags -> compacted pixel list (order unspecified) ; *n_out = number of flagged pixels
// rows restricted to y with ((y / HXR_ROW_BAND) % shard_count) == shard_index
int aa_detect(Context*, const float* vfb, int W, int H, int shard_index, int shard_count, uint32_t* list, uint32_t* n_out, uint8_t* mask);
// vfb[p] *= mul for every flagged pixel in list[0..*n)
int scale_listed(Context*, float* vfb, const uint32_t* list, const uint32_t* n, uint32_t cap, float mul);
// buf[i] *= mul, i < n
int scale_all(Context*, float* buf, size_t n, float mul);
// dst[i] += src[i]
int add_into(Context*, float* dst, const float* src, size_t n);
// float RGB frame -> BMP pixel array (bottom-up BGR rows of rowsz bytes, padding zeroed) through the 4097-entry table `lut`
int to_bmp_rows(Context*, const float* rgb, int W, int H, int rowsz, const uint8_t* lut, uint8_t* out);
// float RGB frame -> half RGBA scan lines as an uncompressed EXR stores them (per row: all A, all B, all G, all R halves)
int to_exr_rows(Context*, const float* rgb, int W, int H, uint16_t* out);
// anaglyph mix of the two eyes' images into out (n_pixels RGB pixels)
int stereo_mix(Context*, float* out, const float* left, const float* right, size_t n_pixels);

// ---- several devices in one process (multi.cpp: one Renderer per device, one host thread each) ----------------------
// Can kernels of `a` read `b`'s device memory directly (NVLink / PCIe peer access, enabled here)? The same device: yes.
bool enable_peer(Context* a, Context* b);
// dst = (dst + srcs[0] + ... + srcs[n_src - 1]) * scale over n floats, ONE kernel on a's device that reads the other
// devices' buffers in place over NVLink: the per-GPU partial frames are summed and resolved (1 / spp) in the same pass
#define HXR_MAX_PEERS 15
int reduce_peers(Context* a, float* dst, const float* const* srcs, int n_src, size_t n, float scale);
// An NCCL communicator over the contexts (ncclCommInitAll; libnccl is loaded at run time, nullptr + err if it is missing or
// refuses, e.g. the same device twice). comm_reduce_sum: ncclReduce(sum) of bufs[i] (on ctxs[i]) into bufs[0] on the contexts' streams.
struct Comm;
Comm* comm_create(Context* const* ctxs, int n, char* err, size_t errlen);
void comm_destroy(Comm*);
bool comm_reduce_sum(Comm*, float* const* bufs, size_t n);

}  // namespace dev
}  // namespace hxr
