// renderer.h — host-side frame driver of the wavefront renderer (one Renderer = one GPU; multi.h drives several).
// Replaces the reference's frame driver (render / renderWithoutMonteCarlo / renderWithMonteCarlo,
// src/main.cpp:253-426) and its thread pool (src/threading.cpp): the per-pixel loops become
// kernel launches over ray queues, the bucket cursor becomes the queues' atomic heads.
#pragma once
#include <string>
#include <vector>
#include "device/launch.h"
#include "host/kdtree.h"

namespace hxr {

// the host-side product of hxr_upload_scene that does not depend on the device: per mesh, the KD-tree and the flattened
// triangle records. Built once and shared by every device of a multi-GPU context (the tree is never built twice).
struct MeshTables {
    host::KdTree kd;
    std::vector<TriTest> tt;
    std::vector<TriAttr> ta;
    std::vector<TriAttrUv> tu;
    std::vector<TriF32> tf;
    std::vector<TriPacked> tp;  // empty: the scene walks tri_f32
    const char* kdSource = "built";  // "built" here (host), "device" (built on the GPU), or "cache" / "waited" (host/cache.h)
};
struct SceneTables {
    std::vector<MeshTables> meshes;
    double build_ms = 0;
};
// validates the scene and builds its host tables (KD-trees with all host threads); false + why on an invalid scene
// buildOn: the context whose GPU builds the KD-trees when the configuration asks for the device build (HXR_CFG_DEVICE_KD_BUILD)
bool buildSceneTables(const hxr_scene& s, const hxr_config& cfg, SceneTables& out, std::string& why, dev::Context* buildOn = nullptr);

class Renderer {
public:
    ~Renderer();
    int create(const hxr_config& cfg, int device);
    int uploadScene(const hxr_scene* sc);                          // builds the tables itself
    int uploadScene(const hxr_scene& sc, const SceneTables& tab);  // tables built by the caller (shared across devices)
    int setCamera(const hxr_camera* cam);
    // noOutput: leave the (partial) frame in frame() and copy it nowhere (multi.cpp sums the shards itself)
    int render(const hxr_render_params& p, float* hostOut, void* devOut, hxr_stats* stats, bool noOutput = false);
    // what render() will do with these parameters: Monte-Carlo or Whitted, and the samples per pixel of the WHOLE frame
    void framePlan(const hxr_render_params& p, bool& mc, int& spp) const;
    void frameSize(const hxr_render_params& p, int& W, int& H) const;  // the frame these parameters render (0 = the scene's own size)
    int resolveDevice(void* d_rgb, int W, int H, int spp);
    int saveFrameBmp(const void* d_rgb, int W, int H, const char* path);
    int saveFrameExr(const void* d_rgb, int W, int H, const char* path);
    int traceClosest(const hxr_ray* rays, size_t n, hxr_hit* hits);
    int traceVisible(const double* seg, size_t n, uint8_t* out);
    int traceColor(const hxr_ray* rays, size_t n, float* rgb);
    int accelInfo(int mesh, hxr_accel_info* out) const;
    void setProfiling(bool on) { if (m_dev) dev::prof_enable(m_dev, on); }
    const std::string& error() const { return m_err; }
    dev::Context* device() const { return m_dev; }
    float* frame() const { return m_accum; }  // the frame of the last render (device memory, W*H*3 floats)
    int frameWidth() const { return m_lastW; }
    int frameHeight() const { return m_lastH; }

private:
    int fail(int code, const std::string& msg) { m_err = msg; return code; }
    bool ensureQueues();
    void freeQueues();
    void freeScene();
    void* keep(void* p) { if (p) m_sceneAllocs.push_back(p); return p; }
    template <class T> T* uploadArray(const T* src, size_t n);
    // run the wavefront on the rays in queue 0 until the ray tree is exhausted (no host read-back between bounces)
    void drain(const FrameParams& fp, float* accum, uint32_t nPrimary, hxr_stats& st);
    int renderOnce(const hxr_render_params& p, float* hostOut, void* devOut, hxr_stats* stats, uint32_t primaryBatch, bool& overflow);
    bool m_noOutput = false;
    uint32_t readCount(const uint32_t* dptr);
    RayQueue queue(int i) const;
    dev::WalkBuffers walkBuffers(CandRec* cand, bool shadow) const;
    ShadowQueue shadowQueue() const;

    std::string m_err;
    hxr_config m_cfg{};
    dev::Context* m_dev = nullptr;
    bool m_haveScene = false, m_haveCamera = false;
    DScene m_scene{};
    std::vector<void*> m_sceneAllocs;
    std::vector<hxr_accel_info> m_accel;
    int m_maxShadowPerHit = 1, m_maxChildrenPerHit = 1;

    // queues: two ray queues (ping-pong per bounce), one shadow queue, one candidate buffer shared by both walks
    uint32_t m_cap = 0, m_shadowCap = 0;
    RayGeom* m_qg[2] = {nullptr, nullptr};
    RayAux* m_qa[2] = {nullptr, nullptr};
    RayGeom* m_sg = nullptr;
    ShadowAux* m_sa = nullptr;
    CandRec* m_cand = nullptr;
    MeshEntry* m_entry = nullptr;    // slot-0 entry records of the closest-hit rays of the current level
    MeshEntry* m_sentry = nullptr;   // ... of the shadow rays
    OverflowEntry* m_ovfList = nullptr;  // rays whose candidate record filled up in the current walk
    OverflowEntry* m_ovfListS = nullptr; // ... of the shadow walk, when it runs on its own lane
    bool m_overlap = false;
    CandRec* m_scand = nullptr;      // shadow candidates of levels shaded in chunks (allocated on first use)
    bool m_allocFailed = false;
    HitRec* m_hits = nullptr;        // test hook only, allocated on first use
    uint8_t* m_visible = nullptr;    // test hook only
    uint32_t* m_counters = nullptr;  // see renderer.cpp (C_*)
    dev::FrameTotals* m_totals = nullptr;
    TravCounters* m_trav = nullptr;
    int m_nBig = 0;
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
    bool m_oneLane = false;
};

}  // namespace hxr
