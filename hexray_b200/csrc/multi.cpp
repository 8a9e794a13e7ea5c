// multi.cpp — see multi.h.
#include "multi.h"
#include <cstdlib>
#include <cstring>
#include <thread>

namespace hxr {

MultiRenderer::~MultiRenderer()
{
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
    if (!buildSceneTables(*sc, m_cfg, tab, why)) return fail(HXR_ERR_INVALID, "invalid scene: " + why);
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
    const int W = r0.frameWidth(), H = r0.frameHeight();
    const size_t n = (size_t)W * H * 3;
    bool mc;
    int spp;
    r0.framePlan(p, mc, spp);
    const float scale = mc ? 1.0f / (float)spp : 1.0f;
    dev::Timer* tm = dev::timer_create(d0);
    dev::timer_start(d0, tm);
    bool ok = true;
    if (m_comm) {
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
    const double reduceMs = dev::timer_ms(d0, tm);
    dev::timer_destroy(d0, tm);
    for (auto& r : m_r) ok = dev::sync(r->device()) && ok;
    if (!ok || dev::failed(d0)) return fail(HXR_ERR_CUDA, std::string("multi-GPU reduce failed: ") + dev::last_error(d0));
    if (devOut) ok = dev::copy_d2d(d0, devOut, r0.frame(), n * sizeof(float)) && dev::sync(d0);
    if (hostOut) ok = ok && dev::download(d0, hostOut, r0.frame(), n * sizeof(float));
    if (!ok) return fail(HXR_ERR_CUDA, std::string("result copy failed: ") + dev::last_error(d0));
    if (stats) {
        hxr_stats s = st[0];
        for (int i = 1; i < N; i++) {
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
        s.n_devices = (uint32_t)N;
        *stats = s;
    }
    return HXR_OK;
}

}  // namespace hxr
