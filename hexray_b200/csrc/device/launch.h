// launch.h — the narrow interface between the host-side frame driver (renderer.cpp) and the
// device. The product implements it in launch_cuda.cu (sm_100a kernels, CUDA streams/events).
// tests/emu/launch_emu.cpp implements the same interface as plain host loops over the SAME
// per-item functions (pipeline.h) — test infrastructure for the CPU tier, never shipped.
#pragma once
#include "pipeline.h"

namespace hxr {
namespace dev {

// ---- device lifetime / memory -------------------------------------------------------
// returns false and fills err if no usable device (product: no CUDA device => hard failure)
bool init(int device, char* err, size_t errlen);
const char* backend_name();
void* alloc(size_t bytes);             // nullptr on failure
void free_(void* p);
bool upload(void* dst, const void* src, size_t bytes);
bool download(void* dst, const void* src, size_t bytes);       // blocking
bool upload_pinned_async(void* dst, const void* src, size_t bytes);
bool zero(void* p, size_t bytes);
bool copy_d2d(void* dst, const void* src, size_t bytes);
bool sync();
const char* last_error();              // text of the last failed call ("" if none)

// timers on the context's stream (CUDA events in the product)
struct Timer;
Timer* timer_create();
void timer_destroy(Timer*);
void timer_start(Timer*);
void timer_stop(Timer*);
double timer_ms(Timer*);               // blocks until the stop event has happened

bool set_u32(uint32_t* p, uint32_t v);  // async, stream ordered

// per-category device time of the launches below (CUDA events around every launch when enabled)
// PROF_WALK: the k_walk launches alone (they are also inside PROF_TRACE_CLOSEST / PROF_TRACE_SHADOW)
enum ProfCat { PROF_TRACE_CLOSEST = 0, PROF_TRACE_SHADOW = 1, PROF_SHADE = 2, PROF_OTHER = 3, PROF_WALK = 4, PROF_NCAT = 5 };
void prof_enable(bool on);
void prof_reset();
void prof_collect(double ms[PROF_NCAT], uint64_t launches[PROF_NCAT]);  // blocks; launches are counted even when disabled

// ---- kernels -------------------------------------------------------------------------
// Every launcher returns the number of kernel launches it issued.

// primary rays for `n_items` (pixel, sample) pairs:
//   item i -> pixel = pixels ? pixels[i / spp_pass] : first_pixel + i / spp_pass,
//             sample = fp.sample_base + (i % spp_pass) * fp.sample_stride
// written to q[0 .. n_items); *q_count is set to n_items.
int gen_primary(const DScene& sc, const FrameParams& fp, const uint32_t* pixels, uint32_t first_pixel,
                uint32_t n_items, uint32_t spp_pass, RayTask* q, uint32_t* q_count);

// closest hit for q[0 .. *q_count) (count read on the device) -> hits[i]. Three stages:
// setup (inline nodes + queue big-mesh walks) -> walk (persistent KD traversal) -> finalize.
// n_hint: a host-side upper bound of *q_count (sizes the grids; small waves do not pay for full-size launches)
int trace_closest(const DScene& sc, const RayTask* q, const uint32_t* q_count, uint32_t q_cap, HitRec* hits,
                  const TraceScratch& ts, TravCounters* cnt, uint32_t n_hint);

// shade q[begin .. min(end, *q_count)); gi selects pathtrace vs Whitted
int shade(const DScene& sc, const FrameParams& fp, const RayTask* q, const uint32_t* q_count, const HitRec* hits,
          uint32_t begin, uint32_t end, const Sinks& sinks);

// visible() for shadow[0 .. *count): ts.occluded[i] = 1 if blocked; when accum != nullptr the carried colour of
// every unblocked task is added to its pixel. *total += *count (64-bit running total kept on the device).
int trace_shadow(const DScene& sc, const ShadowTask* shadow, const uint32_t* count, uint32_t cap, float* accum,
                 const TraceScratch& ts, TravCounters* cnt, unsigned long long* total, uint32_t n_hint);

// needsAA flags -> compacted pixel list (order unspecified) ; *n_out = number of flagged pixels
// rows restricted to y with ((y / HXR_ROW_BAND) % shard_count) == shard_index
int aa_detect(const float* vfb, int W, int H, int shard_index, int shard_count, uint32_t* list, uint32_t* n_out,
              uint8_t* mask);
// vfb[p] *= mul for every flagged pixel in list[0..*n)
int scale_listed(float* vfb, const uint32_t* list, const uint32_t* n, uint32_t cap, float mul);
// buf[i] *= mul, i < n
int scale_all(float* buf, size_t n, float mul);
// dst[i] += src[i]
int add_into(float* dst, const float* src, size_t n);
// float RGB frame -> BMP pixel array (bottom-up BGR rows of rowsz bytes, padding zeroed) through the 4097-entry table `lut`
int to_bmp_rows(const float* rgb, int W, int H, int rowsz, const uint8_t* lut, uint8_t* out);
// anaglyph mix of the two eyes' images into out (n_pixels RGB pixels)
int stereo_mix(float* out, const float* left, const float* right, size_t n_pixels);

}  // namespace dev
}  // namespace hxr
