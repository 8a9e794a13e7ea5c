// multi.cpp — see multi.h.
#include "multi.h"
#include <cstdlib>
#include <cstring>
#include <thread>

namespace hxr {

MultiRenderer::~MultiRenderer()
{
    if (!m_r.empty()) {
        dev::free_(m_r[0]->device(), m_sum);
        dev::free_(m_r[0]->device(), m_estimate);
    }
    dev::comm_destroy(m_comm);
}

int MultiRenderer::create(const hxr_config& cfg)
{
    m_cfg = cfg;
    std::vector<int> devices;
    if (cfg.n_devices > 0) {
        if (cfg.n_devices > HXR_MAX_PEERS + 1) return fail(HXR_ERR_INVALID, "too many devices in hxr_config");
        for (int i = 0; i < cfg.n_devices; i++) devices.push_back(cfg.devices ? cfg.devices[i] : i);
    } else {
        devices.push_back(cfg.device);
    }
    for (int d : devices) {
        std::unique_ptr<Renderer> r(new Renderer);
        const int rc = r->create(cfg, d);
        if (rc != HXR_OK) {
            m_err = r->error();
            m_r.clear();
            return rc;
        }
        m_r.push_back(std::move(r));
    }
    if (m_r.size() > 1) {
        // how the partial frames will meet: peer reads by one kernel on the first GPU (default when every peer is reachable),
        // or an NCCL reduce (HXR_REDUCE=nccl, or when peer access is missing), or - last resort - through the host
        m_peer = true;
        for (size_t i = 1; i < m_r.size(); i++) m_peer = dev::enable_peer(m_r[0]->device(), m_r[i]->device()) && m_peer;
        const char* want = getenv("HXR_REDUCE");
        const bool wantNccl = want && !strcmp(want, "nccl");
        const bool wantHost = want && !strcmp(want, "host");
        if ((wantNccl || !m_peer) && !wantHost) {
            std::vector<dev::Context*> ctxs;
            for (auto& r : m_r) ctxs.push_back(r->device());
            char err[256] = "";
            m_comm = dev::comm_create(ctxs.data(), (int)ctxs.size(), err, sizeof err);
            if (!m_comm && wantNccl) return fail(HXR_ERR_CUDA, std::string("HXR_REDUCE=nccl: ") + err);
        }
        if (wantHost) m_peer = false;
        m_reduceName = m_comm ? "nccl" : (m_peer ? "peer" : "host");
    }
    return HXR_OK;
}

int MultiRenderer::uploadScene(const hxr_scene* sc)
{
    if (m_r.empty()) return fail(HXR_ERR_INVALID, "context not created");
    if (!sc) return fail(HXR_ERR_INVALID, "null scene");
    m_err.clear();
    if (m_r.size() == 1) return m_r[0]->uploadScene(sc);
    // the KD-trees and triangle records once, for every GPU
    SceneTables tab;
    std::string why;
    if (!buildSceneTables(*sc, m_cfg, tab, why, m_r[0]->device())) return fail(HXR_ERR_INVALID, "invalid scene: " + why);
    std::vector<int> rc(m_r.size(), HXR_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < m_r.size(); i++) th.emplace_back([&, i] { rc[i] = m_r[i]->uploadScene(*sc, tab); });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < m_r.size(); i++)
        if (rc[i] != HXR_OK) return fail(rc[i], m_r[i]->error());
    return HXR_OK;
}

int MultiRenderer::setCamera(const hxr_camera* cam)
{
    for (auto& r : m_r) {
        const int rc = r->setCamera(cam);
        if (rc != HXR_OK) return fail(rc, r->error());
    }
    return HXR_OK;
}

int MultiRenderer::reduceShards(float scale, double& reduceMs)
{
    const int N = (int)m_r.size();
    Renderer& r0 = *m_r[0];
    dev::Context* d0 = r0.device();
    const size_t n = (size_t)r0.frameWidth() * r0.frameHeight() * 3;
    dev::Timer* tm = dev::timer_create(d0);
    dev::timer_start(d0, tm);
    bool ok = true;
    if (N == 1) {
        if (scale != 1.0f) dev::scale_all(d0, r0.frame(), n, scale);
    } else if (m_comm) {
        std::vector<float*> bufs;
        for (auto& r : m_r) bufs.push_back(r->frame());
        ok = dev::comm_reduce_sum(m_comm, bufs.data(), n);
        if (ok && scale != 1.0f) dev::scale_all(d0, r0.frame(), n, scale);
    } else if (m_peer) {
        std::vector<const float*> srcs;
        for (int i = 1; i < N; i++) srcs.push_back(m_r[i]->frame());
        dev::reduce_peers(d0, r0.frame(), srcs.data(), N - 1, n, scale);
    } else {
        m_stage.resize(n);
        std::vector<float> sum(n, 0.0f);
        for (int i = 0; i < N && ok; i++) {
            ok = dev::download(m_r[i]->device(), m_stage.data(), m_r[i]->frame(), n * sizeof(float));
            for (size_t k = 0; k < n; k++) sum[k] += m_stage[k];
        }
        for (size_t k = 0; k < n; k++) sum[k] *= scale;
        ok = ok && dev::upload(d0, r0.frame(), sum.data(), n * sizeof(float));
    }
    dev::timer_stop(d0, tm);
    reduceMs = dev::timer_ms(d0, tm);
    dev::timer_destroy(d0, tm);
    for (auto& r : m_r) ok = dev::sync(r->device()) && ok;
    if (!ok || dev::failed(d0)) return fail(HXR_ERR_CUDA, std::string("multi-GPU reduce failed: ") + dev::last_error(d0));
    return HXR_OK;
}

static void mergeStats(std::vector<hxr_stats>& st, double reduceMs, hxr_stats& s)
{
    s = st[0];
    for (size_t i = 1; i < st.size(); i++) {
        s.rays_closest += st[i].rays_closest;
        s.rays_shadow += st[i].rays_shadow;
        s.kd_inner += st[i].kd_inner;
        s.kd_leaves += st[i].kd_leaves;
        s.tri_tests += st[i].tri_tests;
        s.mesh_queries += st[i].mesh_queries;
        s.kernel_launches += st[i].kernel_launches;
        s.cand_overflow += st[i].cand_overflow;
        s.spp_done += st[i].spp_done;
        s.aa_pixels += st[i].aa_pixels;
        s.walk_launches += st[i].walk_launches;
        s.render_ms = std::max(s.render_ms, st[i].render_ms);  // the GPUs run side by side: the frame takes as long as the slowest
        s.walk_ms = std::max(s.walk_ms, st[i].walk_ms);
    }
    s.reduce_ms = reduceMs;
    s.render_ms += reduceMs;
    s.n_devices = (uint32_t)st.size();
}

// ---- progressive frames: pass k of P on GPU g of N renders the sample passes s % (P * N) == k * N + g
int MultiRenderer::progressiveBegin(const hxr_render_params& p, int nPasses)
{
    m_err.clear();
    if (nPasses < 1) return fail(HXR_ERR_INVALID, "progressive: n_passes must be >= 1");
    if (p.shard_count > 1) return fail(HXR_ERR_INVALID, "progressive: the passes are the shards; leave shard_count at 0");
    bool mc;
    int spp;
    m_r[0]->framePlan(p, mc, spp);
    if (!mc) return fail(HXR_ERR_INVALID, "progressive rendering refines Monte-Carlo frames (gi or dof scenes, or mode = HXR_MODE_MONTECARLO)");
    m_pp = p;
    m_pp.mode = HXR_MODE_MONTECARLO;
    m_pp.spp = spp;
    m_passes = nPasses;
    m_passNext = 0;
    m_sppSoFar = 0;
    m_sppTotal = spp;
    return HXR_OK;
}

int MultiRenderer::progressivePass(float* hostOut, hxr_stats* stats)
{
    m_err.clear();
    if (m_passes <= 0) return fail(HXR_ERR_INVALID, "progressive: call hxr_progressive_begin first");
    if (m_passNext >= m_passes) return fail(HXR_ERR_INVALID, "progressive: all passes are done");
    const int N = (int)m_r.size(), P = m_passes, k = m_passNext;
    std::vector<int> rc(N, HXR_OK);
    std::vector<hxr_stats> st(N);
    std::vector<std::thread> th;
    for (int g = 0; g < N; g++)
        th.emplace_back([&, g] {
            hxr_render_params q = m_pp;
            q.shard_index = k * N + g;
            q.shard_count = P * N;
            rc[g] = m_r[g]->render(q, nullptr, nullptr, &st[g], true);
        });
    for (auto& t : th) t.join();
    for (int g = 0; g < N; g++)
        if (rc[g] != HXR_OK) return fail(rc[g], m_r[g]->error());
    double reduceMs = 0;
    const int rrc = reduceShards(1.0f, reduceMs);  // ONE reduce per pass; the sum stays un-normalised
    if (rrc != HXR_OK) return rrc;
    Renderer& r0 = *m_r[0];
    dev::Context* d0 = r0.device();
    const size_t n = (size_t)r0.frameWidth() * r0.frameHeight() * 3;
    if (k == 0 || m_sumFloats != n) {
        if (!ensureSum(n)) return fail(HXR_ERR_CUDA, "progressive: out of device memory");
        dev::zero(d0, m_sum, n * sizeof(float));
    }
    dev::add_into(d0, m_sum, r0.frame(), n);
    for (int g = 0; g < N; g++) m_sppSoFar += (int)st[g].spp_done;
    m_passNext++;
    if (hostOut) {
        dev::copy_d2d(d0, m_estimate, m_sum, n * sizeof(float));
        dev::scale_all(d0, m_estimate, n, 1.0f / (float)std::max(1, m_sppSoFar));
        if (!dev::download(d0, hostOut, m_estimate, n * sizeof(float))) return fail(HXR_ERR_CUDA, dev::last_error(d0));
    } else if (!dev::sync(d0)) {
        return fail(HXR_ERR_CUDA, dev::last_error(d0));
    }
    if (stats) {
        mergeStats(st, reduceMs, *stats);
        stats->spp_done = (uint32_t)m_sppSoFar;
    }
    return HXR_OK;
}

bool MultiRenderer::ensureSum(size_t n)
{
    if (m_sumFloats == n) return true;
    dev::Context* d0 = m_r[0]->device();
    dev::free_(d0, m_sum);
    dev::free_(d0, m_estimate);
    m_sum = (float*)dev::alloc(d0, n * sizeof(float));
    m_estimate = (float*)dev::alloc(d0, n * sizeof(float));
    m_sumFloats = (m_sum && m_estimate) ? n : 0;
    return m_sumFloats != 0;
}

// A frame interrupted after passesDone passes goes on from its checkpoint (hxr_progressive_state of the run that wrote it).
int MultiRenderer::progressiveResume(const hxr_render_params& p, int nPasses, const float* sum, int passesDone, int sppDone)
{
    const int rc = progressiveBegin(p, nPasses);
    if (rc != HXR_OK) return rc;
    const int N = (int)m_r.size();
    // the samples the first passesDone passes of THIS plan cover: s in [0, spp) with s % (P * N) < passesDone * N
    const long long period = (long long)nPasses * N, done = (long long)passesDone * N;
    const long long expect = passesDone >= 0 && passesDone <= nPasses ? (m_sppTotal / period) * done + std::min<long long>(m_sppTotal % period, done) : -1;
    if (!sum || passesDone < 1 || passesDone > nPasses || sppDone != expect) {
        m_passes = 0;  // (no frame is open)
        return fail(HXR_ERR_INVALID, "progressive resume: the checkpoint does not belong to this frame plan (passes, samples per pixel or GPU count differ)");
    }
    int W = 0, H = 0;
    m_r[0]->frameSize(p, W, H);
    const size_t n = (size_t)W * H * 3;
    if (n == 0 || !ensureSum(n)) { m_passes = 0; return fail(HXR_ERR_CUDA, "progressive: out of device memory"); }
    dev::Context* d0 = m_r[0]->device();
    if (!dev::upload(d0, m_sum, sum, n * sizeof(float)) || !dev::sync(d0)) { m_passes = 0; return fail(HXR_ERR_CUDA, dev::last_error(d0)); }
    m_passNext = passesDone;
    m_sppSoFar = sppDone;
    return HXR_OK;
}

int MultiRenderer::progressiveState(float* sumOut, int* passesDone, int* sppDone)
{
    if (passesDone) *passesDone = m_passNext;
    if (sppDone) *sppDone = m_sppSoFar;
    if (sumOut) {
        if (!m_sum || m_passNext == 0) return fail(HXR_ERR_INVALID, "progressive: no pass rendered yet");
        if (!dev::download(m_r[0]->device(), sumOut, m_sum, m_sumFloats * sizeof(float))) return fail(HXR_ERR_CUDA, dev::last_error(m_r[0]->device()));
    }
    return HXR_OK;
}

int MultiRenderer::render(const hxr_render_params& p, float* hostOut, void* devOut, hxr_stats* stats)
{
    m_err.clear();
    const int N = (int)m_r.size();
    if (N == 1) return m_r[0]->render(p, hostOut, devOut, stats);
    if (p.shard_count > 1) return fail(HXR_ERR_INVALID, "a multi-GPU context shards the frame itself: leave hxr_render_params.shard_count at 0");
    if (!hostOut && !devOut) return fail(HXR_ERR_INVALID, "render: no output buffer");
    // fork: every GPU renders its shard into its own frame buffer
    std::vector<int> rc(N, HXR_OK);
    std::vector<hxr_stats> st(N);
    std::vector<std::thread> th;
    for (int i = 0; i < N; i++)
        th.emplace_back([&, i] {
            hxr_render_params q = p;
            q.shard_index = i;
            q.shard_count = N;
            rc[i] = m_r[i]->render(q, nullptr, nullptr, &st[i], true);
        });
    for (auto& t : th) t.join();  // join (each render ends with its stream synchronised)
    for (int i = 0; i < N; i++)
        if (rc[i] != HXR_OK) return fail(rc[i], m_r[i]->error());
    // the frames meet on the first GPU: sum, and for Monte-Carlo frames the division by the sample count, in one pass
    Renderer& r0 = *m_r[0];
    dev::Context* d0 = r0.device();
    const size_t n = (size_t)r0.frameWidth() * r0.frameHeight() * 3;
    bool mc;
    int spp;
    r0.framePlan(p, mc, spp);
    double reduceMs = 0;
    const int rrc = reduceShards(mc ? 1.0f / (float)spp : 1.0f, reduceMs);
    if (rrc != HXR_OK) return rrc;
    bool ok = true;
    if (devOut) ok = dev::copy_d2d(d0, devOut, r0.frame(), n * sizeof(float)) && dev::sync(d0);
    if (hostOut) ok = ok && dev::download(d0, hostOut, r0.frame(), n * sizeof(float));
    if (!ok) return fail(HXR_ERR_CUDA, std::string("result copy failed: ") + dev::last_error(d0));
    if (stats) mergeStats(st, reduceMs, *stats);
    return HXR_OK;
}

}  // namespace hxr
