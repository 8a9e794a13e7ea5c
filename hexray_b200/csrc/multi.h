// multi.h — one context over several GPUs of the box, owned by the library.
// Replaces the reference's ThreadPool (src/threading.cpp:54-97: a blocking fork-join of one lambda over N host threads, each
// pulling 64x64 buckets): here the fork-join is over N GPUs, one host thread each, every GPU rendering its shard of the frame
// (Monte-Carlo: sample passes s % N == g of every pixel; Whitted: 16-row bands, see renderer.cpp) into its own accumulation
// buffer; the partial frames are then summed and resolved on the first GPU - by one kernel that reads the peers' buffers in
// place over NVLink (dev::reduce_peers), or by an NCCL reduce (ncclCommInitAll communicator owned by this object) - and only
// the finished frame crosses PCIe. The scene tables (KD-trees, flattened triangles) are built ONCE on the host and uploaded to
// every GPU; nothing is ever exchanged between GPUs while rays are in flight.
#pragma once
#include <memory>
#include <vector>
#include "renderer.h"

namespace hxr {

class MultiRenderer {
public:
    ~MultiRenderer();
    int create(const hxr_config& cfg);
    int uploadScene(const hxr_scene* sc);
    int setCamera(const hxr_camera* cam);
    int render(const hxr_render_params& p, float* hostOut, void* devOut, hxr_stats* stats);
    // progressive Monte-Carlo frames: see include/hxr.h
    int progressiveBegin(const hxr_render_params& p, int nPasses);
    int progressivePass(float* hostOut, hxr_stats* stats);
    int progressiveState(float* sumOut, int* passesDone, int* sppDone);
    int progressiveResume(const hxr_render_params& p, int nPasses, const float* sum, int passesDone, int sppDone);
    int deviceCount() const { return (int)m_r.size(); }
    Renderer& primary() { return *m_r[0]; }
    const std::string& error() const { return m_err.empty() ? m_r[0]->error() : m_err; }
    void setProfiling(bool on) { for (auto& r : m_r) r->setProfiling(on); }
    const char* reduceName() const { return m_reduceName; }

private:
    int fail(int code, const std::string& msg) { m_err = msg; return code; }
    std::vector<std::unique_ptr<Renderer>> m_r;
    std::string m_err;
    hxr_config m_cfg{};
    dev::Comm* m_comm = nullptr;
    bool m_peer = false;       // the first GPU can read every other GPU's memory
    const char* m_reduceName = "none";
    std::vector<float> m_stage;  // host staging of the last-resort reduce
    int reduceShards(float scale, double& reduceMs);  // the GPUs' partial frames -> the first GPU's frame, times scale
    // progressive state: the running sum lives on the first GPU
    hxr_render_params m_pp{};
    int m_passes = 0, m_passNext = 0, m_sppSoFar = 0, m_sppTotal = 0;
    float* m_sum = nullptr;
    float* m_estimate = nullptr;
    size_t m_sumFloats = 0;
    bool ensureSum(size_t n);  // the running sum and the estimate buffer on the first GPU, n floats each
};

}  // namespace hxr
