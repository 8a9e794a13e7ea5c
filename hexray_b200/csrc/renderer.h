// renderer.h — host-side frame driver of the wavefront renderer (one context = one GPU).
// Replaces the reference's frame driver (render / renderWithoutMonteCarlo / renderWithMonteCarlo,
// src/main.cpp:253-426) and its thread pool (src/threading.cpp): the per-pixel loops become
// kernel launches over ray queues, the bucket cursor becomes the queues' atomic heads.
#pragma once
#include <string>
#include <vector>
#include "device/launch.h"
#include "host/kdtree.h"

namespace hxr {

class Renderer {
public:
    ~Renderer();
    int create(const hxr_config& cfg);
    int uploadScene(const hxr_scene* sc);
    int setCamera(const hxr_camera* cam);
    int render(const hxr_render_params& p, float* hostOut, void* devOut, hxr_stats* stats);
    int resolveDevice(void* d_rgb, int W, int H, int spp);
    int saveFrameBmp(const void* d_rgb, int W, int H, const char* path);
    int traceClosest(const hxr_ray* rays, size_t n, hxr_hit* hits);
    int traceVisible(const double* seg, size_t n, uint8_t* out);
    int traceColor(const hxr_ray* rays, size_t n, float* rgb);
    int accelInfo(int mesh, hxr_accel_info* out) const;
    const std::string& error() const { return m_err; }

private:
    int fail(int code, const std::string& msg) { m_err = msg; return code; }
    bool ensureQueues();
    void freeScene();
    void* keep(void* p) { if (p) m_sceneAllocs.push_back(p); return p; }
    template <class T> T* uploadArray(const T* src, size_t n);
    // run the wavefront until the current queue drains; `gi` selects the integrator
    int drain(const FrameParams& fp, float* accum, uint32_t nPrimary, hxr_stats& st);
    int renderOnce(const hxr_render_params& p, float* hostOut, void* devOut, hxr_stats* stats, uint32_t primaryBatch, bool& overflow);
    uint32_t readCount(const uint32_t* dptr);

    std::string m_err;
    hxr_config m_cfg{};
    bool m_created = false, m_haveScene = false, m_haveCamera = false;
    DScene m_scene{};
    std::vector<void*> m_sceneAllocs;
    std::vector<hxr_accel_info> m_accel;
    int m_maxShadowPerHit = 1, m_maxChildrenPerHit = 1;

    // queues
    uint32_t m_cap = 0, m_shadowCap = 0;
    RayTask* m_q[2] = {nullptr, nullptr};
    HitRec* m_hits = nullptr;
    ShadowTask* m_shadow = nullptr;
    uint32_t* m_counters = nullptr;  // [0],[1] queue counts, [2] shadow count, [3] overflow, [4] work head A, [5] work head B, [6] aa count
    TravCounters* m_trav = nullptr;
    // traversal scratch (device/pipeline.h: TraceScratch)
    RayPre* m_pre = nullptr;
    WalkTask* m_tasks = nullptr;
    PairRec* m_pairs = nullptr;
    double* m_pairGamma = nullptr;
    uint32_t m_pairCap = 0;
    MeshRes* m_res = nullptr;
    uint8_t* m_occluded = nullptr;
    uint32_t m_taskCap = 0;
    int m_nBig = 0;
    TraceScratch scratch(uint32_t headCounter) const;
    uint32_t* m_aaList = nullptr;
    uint8_t* m_aaMask = nullptr;
    float* m_accum = nullptr;
    int m_lastW = 0, m_lastH = 0;    // size of the frame m_accum holds
    uint8_t* m_srgbLut = nullptr;    // the reference's 4097-entry sRGB table on the device
    float* m_eye[2] = {nullptr, nullptr};  // per-eye accumulation of anaglyph frames
    size_t m_eyePixels = 0;
    size_t m_accumPixels = 0;
    size_t m_aaCap = 0;
    bool m_countTraversal = false;
};

}  // namespace hxr
