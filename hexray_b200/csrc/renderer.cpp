// renderer.cpp — see renderer.h. Host logic only: validation, KD build, uploads, and the
// wavefront schedule. Everything per-ray runs in the kernels behind device/launch.h.
#include "renderer.h"
#include "kd_device_build.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>
#include "host/bitmap.h"
#include "host/cache.h"
#include "host/exr_codec.h"

namespace hxr {

// device counters (uint32): queue counts, the shadow queue's count, the work-fetch cursors of the two walks, flags
// (HEAD / OVF / FETCH of a walk are consecutive: one clear before the walk is queued)
enum { C_Q0 = 0, C_Q1 = 1, C_SHADOW = 2, C_HEAD_A = 3, C_OVF_A = 4, C_FETCH_A = 5, C_HEAD_B = 6, C_OVF_B = 7, C_FETCH_B = 8, C_OVERFLOW = 9, C_AA = 10,
       C_NCOUNTERS = 16 };

Renderer::~Renderer()
{
    if (!m_dev) return;
    freeScene();
    freeQueues();
    dev::free_(m_dev, m_counters);
    dev::free_(m_dev, m_totals);
    dev::free_(m_dev, m_trav);
    dev::free_(m_dev, m_aaList);
    dev::free_(m_dev, m_aaMask);
    dev::free_(m_dev, m_accum);
    dev::free_(m_dev, m_srgbLut);
    dev::free_(m_dev, m_eye[0]);
    dev::free_(m_dev, m_eye[1]);
    dev::destroy(m_dev);
}

int Renderer::create(const hxr_config& cfg, int device)
{
    m_cfg = cfg;
    char err[256] = "";
    m_dev = dev::create(device, err, sizeof err);
    if (!m_dev) return fail(HXR_ERR_NO_DEVICE, err);
    return HXR_OK;
}

void Renderer::freeScene()
{
    for (void* p : m_sceneAllocs) dev::free_(m_dev, p);
    m_sceneAllocs.clear();
    m_haveScene = false;
    m_accel.clear();
}

template <class T> T* Renderer::uploadArray(const T* src, size_t n)
{
    // a 1-element allocation keeps device pointers non-null for empty tables
    T* d = (T*)keep(dev::alloc(m_dev, std::max<size_t>(n, 1) * sizeof(T)));
    if (d && n && !dev::upload(m_dev, d, src, n * sizeof(T))) return nullptr;
    return d;
}

// ------------------------------------------------------------------------------ scene
static bool validateScene(const hxr_scene& s, std::string& why)
{
    auto bad = [&](const std::string& m) { why = m; return false; };
    if (s.abi_version != HXR_ABI_VERSION) return bad("hxr_scene.abi_version mismatch");
    if (s.n_nodes < 0 || s.n_geometries < 0 || s.n_meshes < 0 || s.n_heightfields < 0 || s.n_shaders < 0 || s.n_layers < 0 ||
        s.n_textures < 0 || s.n_images < 0 || s.n_lights < 0)
        return bad("negative table size");
    auto in = [](int i, int n) { return i >= 0 && i < n; };
    for (int i = 0; i < s.n_images; i++)
        if (s.images[i].width < 0 || s.images[i].height < 0 || ((size_t)s.images[i].width * s.images[i].height > 0 && !s.images[i].rgb))
            return bad("image " + std::to_string(i) + " has no pixels");
    for (int i = 0; i < s.n_textures; i++) {
        const hxr_texture& t = s.textures[i];
        if (t.type < HXR_TEX_CHECKER || t.type > HXR_TEX_BUMPS) return bad("texture type");
        if ((t.type == HXR_TEX_BITMAP || t.type == HXR_TEX_BUMP) && !in(t.image, s.n_images)) return bad("texture image index");
    }
    for (int i = 0; i < s.n_layers; i++)
        if (!in(s.layers[i].shader, s.n_shaders) || (s.layers[i].tex != -1 && !in(s.layers[i].tex, s.n_textures))) return bad("layer reference");
    for (int i = 0; i < s.n_shaders; i++) {
        const hxr_shader& sh = s.shaders[i];
        if (sh.type < HXR_SHADER_LAMBERT || sh.type > HXR_SHADER_CONST) return bad("shader type");
        if (sh.tex != -1 && !in(sh.tex, s.n_textures)) return bad("shader texture index");
        if (sh.type == HXR_SHADER_LAYERED && (sh.n_layers < 0 || sh.first_layer < 0 || sh.first_layer + sh.n_layers > s.n_layers)) return bad("layer range");
        if (sh.type == HXR_SHADER_REFLECTION && sh.i0 < 1) return bad("Reflection numSamples");
    }
    // layered shaders must form a DAG of bounded size (the device evaluates them with a fixed stack)
    std::function<long(int, int)> leaves = [&](int si, int depth) -> long {
        if (depth > 8) return 1L << 20;
        const hxr_shader& sh = s.shaders[si];
        if (sh.type != HXR_SHADER_LAYERED) return 1;
        long n = 0;
        for (int l = 0; l < sh.n_layers; l++) n += leaves(s.layers[sh.first_layer + l].shader, depth + 1);
        return n;
    };
    for (int i = 0; i < s.n_shaders; i++)
        if (leaves(i, 0) > HXR_SHADE_STACK - 4) return bad("Layered shader too deep/wide (or cyclic)");
    for (int i = 0; i < s.n_geometries; i++) {
        const hxr_geometry& g = s.geometries[i];
        switch (g.type) {
            case HXR_GEOM_PLANE: case HXR_GEOM_SPHERE: case HXR_GEOM_CUBE: break;
            case HXR_GEOM_CSG:
                if (g.a < 0 || g.a > 2 || !in(g.b, s.n_geometries) || !in(g.c, s.n_geometries)) return bad("CSG reference");
                break;
            case HXR_GEOM_MESH: if (!in(g.a, s.n_meshes)) return bad("mesh index"); break;
            case HXR_GEOM_HEIGHTFIELD: if (!in(g.a, s.n_heightfields)) return bad("heightfield index"); break;
            default: return bad("geometry type");
        }
    }
    for (int i = 0; i < s.n_meshes; i++) {
        const hxr_mesh& m = s.meshes[i];
        if (m.n_vertices < 1 || m.n_normals < 1 || m.n_uvs < 1 || m.n_triangles < 0 || !m.vertices || !m.normals || !m.uvs || (m.n_triangles && !m.triangles))
            return bad("mesh arrays");
        for (int t = 0; t < m.n_triangles; t++)
            for (int k = 0; k < 3; k++)
                if (!in(m.triangles[t].v[k], m.n_vertices) || !in(m.triangles[t].n[k], m.n_normals) || !in(m.triangles[t].t[k], m.n_uvs))
                    return bad("mesh " + std::to_string(i) + ": triangle index out of range");
    }
    for (int i = 0; i < s.n_heightfields; i++) {
        const hxr_heightfield& h = s.heightfields[i];
        if (h.width < 1 || h.height < 1 || !h.heights || !h.max_h || !h.normals) return bad("heightfield arrays");
        if (h.use_optimization && (!h.high_map || h.max_k < 0 || h.max_k > 16)) return bad("heightfield high map");
    }
    for (int i = 0; i < s.n_nodes; i++) {
        const hxr_node& n = s.nodes[i];
        if (!in(n.geom, s.n_geometries) || !in(n.shader, s.n_shaders)) return bad("node reference");
        if (n.bump_tex != -1 && !in(n.bump_tex, s.n_textures)) return bad("node bump texture");
    }
    for (int i = 0; i < s.n_lights; i++) {
        const hxr_light& l = s.lights[i];
        if (l.type != HXR_LIGHT_POINT && l.type != HXR_LIGHT_RECT) return bad("light type");
        if (l.type == HXR_LIGHT_RECT && (l.xsubd < 1 || l.ysubd < 1 || l.xsubd > 1000 || l.ysubd > 1000)) return bad("RectLight subdivisions");
    }
    if (s.has_environment)
        for (int i = 0; i < 6; i++)
            if (!in(s.env_images[i], s.n_images)) return bad("environment image index");
    return true;
}


// ---- the device-independent part of a scene upload: KD-trees and flattened triangle records
bool buildSceneTables(const hxr_scene& s, const hxr_config& cfg, SceneTables& out, std::string& why, dev::Context* buildOn)
{
    if (!validateScene(s, why)) return false;
    out.meshes.clear();
    out.meshes.resize(s.n_meshes);
    out.build_ms = 0;
    size_t filterBytes = 0;
    for (int i = 0; i < s.n_meshes; i++) filterBytes += (size_t)s.meshes[i].n_triangles * sizeof(TriF32);
    // The walk's optional 32-byte form of the triangles (TriPacked: one sector per test instead of 1.5-2, ~35 more instructions
    // per test). Round 1's walk waited on memory and gained 2.7 % from it on terrain-10M; since round 2 the walk is bound by
    // instruction issue (ncu: 70-78 % issue-active) and the 48-byte TriF32 records are 2.5 % FASTER there (profiles/README.md),
    // as they always were on L2-resident scenes. So: off unless HXR_TRI_PACK=1 asks for it (kept for scenes that outgrow HBM
    // bandwidth and for the layout-invariance test). A mesh with an edge that does not fit the format keeps the scene on TriF32.
    const bool wantPack = getenv("HXR_TRI_PACK") ? atoi(getenv("HXR_TRI_PACK")) != 0 : false;
    (void)filterBytes;
    for (int i = 0; i < s.n_meshes; i++) {
        const hxr_mesh& m = s.meshes[i];
        MeshTables& M = out.meshes[i];
        // through the cache (host/cache.h): a big tree built before - or being built right now by another process of this
        // box, e.g. the other ranks of a torchrun job - is loaded instead of built again
        const char* how = "built";
        const char* envBuild = getenv("HXR_KD_BUILD");
        const bool onDevice = buildOn && m.n_triangles > 1 && (envBuild ? !strcmp(envBuild, "device") : (cfg.flags & HXR_CFG_DEVICE_KD_BUILD) != 0);
        if (onDevice) {
            // the tree is built by the GPU, level by level (kd_device_build.cpp); nothing is read from or written to the cache
            std::string err;
            if (!buildKdTreeOnDevice(buildOn, m, host::KdBuildParams(), M.kd, err)) { why = err; return false; }
            how = "device";
        } else {
            host::cachedKdTree(m, host::KdBuildParams(), M.kd, &how);
        }
        M.kdSource = how;
        if (!strcmp(how, "built") || !strcmp(how, "device")) out.build_ms += M.kd.buildMs;
        const size_t nt = (size_t)m.n_triangles;
        M.tt.resize(nt);
        M.ta.resize(nt);
        M.tu.resize(nt);
        M.tf.resize(nt);
        auto fill = [&](size_t t0, size_t t1) {
            for (size_t t = t0; t < t1; t++) {
                const hxr_triangle& T = m.triangles[t];
                for (int k = 0; k < 3; k++) {
                    M.tt[t].A[k] = m.vertices[3 * (size_t)T.v[0] + k];
                    M.tt[t].AB[k] = T.ab[k];
                    M.tt[t].AC[k] = T.ac[k];
                    M.tt[t].N[k] = T.ab_cross_ac[k];
                    M.tf[t].A[k] = (float)M.tt[t].A[k];
                    M.tf[t].AB[k] = (float)M.tt[t].AB[k];
                    M.tf[t].AC[k] = (float)M.tt[t].AC[k];
                    M.tf[t].N[k] = (float)M.tt[t].N[k];
                    M.ta[t].gnormal[k] = T.gnormal[k];
                    for (int c = 0; c < 3; c++) M.ta[t].nrm[k][c] = m.normals[3 * (size_t)T.n[k] + c];
                    for (int c = 0; c < 2; c++) M.tu[t].uv[k][c] = m.uvs[3 * (size_t)T.t[k] + c];
                    M.tu[t].dNdx[k] = T.dndx[k];
                    M.tu[t].dNdy[k] = T.dndy[k];
                }
            }
        };
        // big meshes: the record fill is a pure streaming pass, split over the host threads
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const size_t nth = nt > (1u << 18) ? std::min<size_t>(hw, 32) : 1;
        if (nth <= 1) fill(0, nt);
        else {
            std::vector<std::thread> th;
            for (size_t k = 0; k < nth; k++) th.emplace_back(fill, nt * k / nth, nt * (k + 1) / nth);
            for (auto& x : th) x.join();
        }
        M.tp.clear();
        if (wantPack) {
            M.tp.resize(nt);
            for (size_t t = 0; t < nt; t++)
                if (!pack_tri(M.tt[t], M.tp[t])) { M.tp.clear(); break; }
        }
    }
    return true;
}

int Renderer::uploadScene(const hxr_scene* sp)
{
    if (!m_dev) return fail(HXR_ERR_INVALID, "context not created");
    if (!sp) return fail(HXR_ERR_INVALID, "null scene");
    SceneTables tab;
    std::string why;
    if (!buildSceneTables(*sp, m_cfg, tab, why, m_dev)) return fail(HXR_ERR_INVALID, "invalid scene: " + why);
    return uploadScene(*sp, tab);
}

int Renderer::uploadScene(const hxr_scene& s, const SceneTables& tab)
{
    if (!m_dev) return fail(HXR_ERR_INVALID, "context not created");
    if ((int)tab.meshes.size() != s.n_meshes) return fail(HXR_ERR_INVALID, "scene tables do not match the scene");
    freeScene();
    auto oom = [&]() { freeScene(); return fail(HXR_ERR_CUDA, std::string("scene upload failed: ") + dev::last_error(m_dev)); };

    std::vector<DMesh> dm(s.n_meshes);
    m_accel.resize(s.n_meshes);
    for (int i = 0; i < s.n_meshes; i++) {
        const hxr_mesh& m = s.meshes[i];
        const MeshTables& M = tab.meshes[i];
        DMesh& d = dm[i];
        memset(&d, 0, sizeof d);
        if (!M.tp.empty()) {
            d.tri_pk = uploadArray(M.tp.data(), M.tp.size());
            if (!d.tri_pk) return oom();
        }
        d.blocks = uploadArray(M.kd.blocks.data(), M.kd.blocks.size());
        d.leaf_tris = uploadArray(M.kd.leafTris.data(), M.kd.leafTris.size());
        d.tri_test = uploadArray(M.tt.data(), M.tt.size());
        d.tri_f32 = uploadArray(M.tf.data(), M.tf.size());
        d.tri_attr = uploadArray(M.ta.data(), M.ta.size());
        d.tri_attr_uv = uploadArray(M.tu.data(), M.tu.size());
        if (!d.blocks || !d.leaf_tris || !d.tri_test || !d.tri_f32 || !d.tri_attr || !d.tri_attr_uv) return oom();
        double amax = 0;
        for (int k = 0; k < 3; k++) {
            d.bbmin[k] = m.bbox_min[k];
            d.bbmax[k] = m.bbox_max[k];
            amax = std::max(amax, std::max(std::fabs(m.bbox_min[k]), std::fabs(m.bbox_max[k])));
        }
        d.abs_max = std::nextafter((float)amax, INFINITY);
        for (int k = 0; k < 3; k++) {
            d.fbmin[k] = std::nextafter((float)(m.bbox_min[k] - 1e-6), -INFINITY);
            d.fbmax[k] = std::nextafter((float)(m.bbox_max[k] + 1e-6), INFINITY);
        }
        d.faceted = m.faceted;
        d.backface = m.backface_culling;
        d.n_tris = m.n_triangles;
        const int smallMesh = getenv("HXR_SMALL_MESH") ? atoi(getenv("HXR_SMALL_MESH")) : HXR_SMALL_MESH;
        d.brute = (m.n_triangles <= smallMesh || (m_cfg.flags & HXR_CFG_BRUTE_FORCE_MESHES)) ? 1 : 0;
        if (d.brute && m.n_triangles <= HXR_SMALL_MESH && !getenv("HXR_QUAD_SLAB")) d.brute = 3;  // tiny: no box gate either
        hxr_accel_info& ai = m_accel[i];
        ai.nodes = M.kd.blocks.size();
        ai.leaves = M.kd.leaves;
        ai.tri_refs = M.kd.leafTris.size();
        ai.bytes_nodes = M.kd.blocks.size() * sizeof(KdBlock);
        ai.bytes_tris = M.kd.leafTris.size() * sizeof(uint32_t) + (M.tp.empty() ? M.tf.size() * sizeof(TriF32) : M.tp.size() * sizeof(TriPacked)) + M.tt.size() * sizeof(TriTest);
        ai.max_depth = M.kd.maxDepth;
        ai.n_triangles = (uint32_t)m.n_triangles;
        ai.build_ms = M.kd.buildMs;
        ai.from_cache = strcmp(M.kdSource, "built") != 0 && strcmp(M.kdSource, "device") != 0;
        ai.device_build = !strcmp(M.kdSource, "device");
        ai.device_ms = M.kd.deviceMs;
    }
    std::vector<DHeightfield> dh(s.n_heightfields);
    for (int i = 0; i < s.n_heightfields; i++) {
        const hxr_heightfield& h = s.heightfields[i];
        const size_t n = (size_t)h.width * h.height;
        DHeightfield& d = dh[i];
        memset(&d, 0, sizeof d);
        d.heights = uploadArray(h.heights, n);
        d.max_h = uploadArray(h.max_h, n);
        d.normals = uploadArray(h.normals, n * 3);
        d.high_map = h.high_map ? uploadArray(h.high_map, n * 16) : nullptr;
        if (!d.heights || !d.max_h || !d.normals || (h.high_map && !d.high_map)) return oom();
        for (int k = 0; k < 3; k++) { d.bbmin[k] = h.bbox_min[k]; d.bbmax[k] = h.bbox_max[k]; }
        d.W = h.width;
        d.H = h.height;
        d.use_opt = h.use_optimization && h.high_map;
        d.max_k = h.max_k;
    }
    std::vector<DImage> di(s.n_images);
    for (int i = 0; i < s.n_images; i++) {
        const size_t n = (size_t)s.images[i].width * s.images[i].height * 3;
        di[i].w = s.images[i].width;
        di[i].h = s.images[i].height;
        di[i].rgb = n ? uploadArray(s.images[i].rgb, n) : nullptr;
        if (n && !di[i].rgb) return oom();
    }
    memset(&m_scene, 0, sizeof m_scene);
    {
        // the device copy of the node table carries one derived flag: an untransformed node (HXR_NODE_IDENT)
        std::vector<hxr_node> nodes(s.nodes, s.nodes + s.n_nodes);
        for (hxr_node& nd : nodes) {
            bool ident = !getenv("HXR_NO_IDENT");
            for (int k = 0; k < 9 && ident; k++) {
                const double want = (k % 4 == 0) ? 1.0 : 0.0;
                ident = nd.T.m[k] == want && nd.T.inv[k] == want && nd.T.inv_t[k] == want;
            }
            nd.pad = ident ? HXR_NODE_IDENT : 0;
        }
        m_scene.nodes = uploadArray(nodes.data(), nodes.size());
    }
    m_scene.geoms = uploadArray(s.geometries, s.n_geometries);
    m_scene.meshes = uploadArray(dm.data(), dm.size());
    m_scene.hfs = uploadArray(dh.data(), dh.size());
    m_scene.shaders = uploadArray(s.shaders, s.n_shaders);
    m_scene.layers = uploadArray(s.layers, s.n_layers);
    m_scene.textures = uploadArray(s.textures, s.n_textures);
    m_scene.images = uploadArray(di.data(), di.size());
    m_scene.lights = uploadArray(s.lights, s.n_lights);
    {
        // world boxes of the rect lights (unit square in local xz, src/lights.cpp:53-73), padded; point lights cannot be hit
        std::vector<float> lb((size_t)std::max(1, s.n_lights) * 6, 0.0f);
        for (int i = 0; i < s.n_lights; i++) {
            float* b = &lb[(size_t)i * 6];
            if (s.lights[i].type != HXR_LIGHT_RECT) { for (int k = 0; k < 3; k++) { b[k] = INFINITY; b[3 + k] = -INFINITY; } continue; }
            const hxr_transform& T = s.lights[i].T;
            double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
            for (int c = 0; c < 4; c++) {
                const double p[3] = {(c & 1) ? 0.5 : -0.5, 0.0, (c & 2) ? 0.5 : -0.5};
                for (int k = 0; k < 3; k++) {
                    const double w = p[0] * T.m[k] + p[1] * T.m[3 + k] + p[2] * T.m[6 + k] + T.offset[k];
                    mn[k] = std::min(mn[k], w);
                    mx[k] = std::max(mx[k], w);
                }
            }
            for (int k = 0; k < 3; k++) {
                const double pad = 1e-4 * (1.0 + std::max(std::fabs(mn[k]), std::fabs(mx[k])));
                const bool ok = std::isfinite(mn[k]) && std::isfinite(mx[k]) && std::fabs(mn[k]) < 1e30 && std::fabs(mx[k]) < 1e30;
                b[k] = ok ? std::nextafter((float)(mn[k] - pad), -INFINITY) : -INFINITY;
                b[3 + k] = ok ? std::nextafter((float)(mx[k] + pad), INFINITY) : INFINITY;
            }
        }
        m_scene.light_box = uploadArray(lb.data(), lb.size());
        if (!m_scene.light_box) return oom();
    }
    if (!m_scene.nodes || !m_scene.geoms || !m_scene.meshes || !m_scene.hfs || !m_scene.shaders || !m_scene.layers ||
        !m_scene.textures || !m_scene.images || !m_scene.lights)
        return oom();
    {
        // nodes whose geometry is a mesh too big for the inline brute-force test get a result slot
        // (at most HXR_MAX_WALKED_NODES of them: a candidate record names the slot in 8 bits; further mesh nodes are
        // intersected inline with the double walk)
        std::vector<int32_t> slot(std::max(1, s.n_nodes), -1);
        std::vector<int32_t> bigNodes;
        m_nBig = 0;
        for (int i = 0; i < s.n_nodes; i++) {
            const hxr_geometry& g = s.geometries[s.nodes[i].geom];
            if (g.type == HXR_GEOM_MESH && !dm[g.a].brute && m_nBig < HXR_MAX_WALKED_NODES) {
                slot[i] = m_nBig++;
                bigNodes.push_back(i);
            }
        }
        m_scene.node_slot = uploadArray(slot.data(), slot.size());
        m_scene.big_nodes = uploadArray(bigNodes.data(), bigNodes.size());
        if (!m_scene.node_slot || !m_scene.big_nodes) return oom();
        // nodes on which nothing reads a hit's u, v, dNdx, dNdy: no bump map and no texture anywhere in the shader tree
        std::function<bool(int, int)> textured = [&](int si, int depth) -> bool {
            if (si < 0 || si >= s.n_shaders || depth > 16) return true;  // unknown: assume it reads them
            const hxr_shader& sh = s.shaders[si];
            if (sh.tex >= 0) return true;
            if (sh.type == HXR_SHADER_LAYERED)
                for (int l = sh.first_layer; l < sh.first_layer + sh.n_layers; l++)
                    if (s.layers[l].tex >= 0 || textured(s.layers[l].shader, depth + 1)) return true;
            return false;
        };
        std::vector<int32_t> lean(std::max(1, s.n_nodes), 0);
        for (int i = 0; i < s.n_nodes; i++) lean[i] = (s.nodes[i].bump_tex < 0 && !textured(s.nodes[i].shader, 0) && !getenv("HXR_FULL_ATTR")) ? 1 : 0;
        m_scene.node_lean = uploadArray(lean.data(), lean.size());
        if (!m_scene.node_lean) return oom();
        m_scene.n_big = m_nBig;
        // world boxes of the inline nodes' geometry: the 8 corners of the object-space box through the node transform
        std::vector<double> box((size_t)std::max(1, s.n_nodes) * 6);
        for (int i = 0; i < s.n_nodes; i++) {
            double* b = &box[(size_t)i * 6];
            for (int k = 0; k < 3; k++) { b[k] = -1e300; b[3 + k] = 1e300; }
            const hxr_geometry& g = s.geometries[s.nodes[i].geom];
            double lo[3], hi[3];
            bool bounded = true;
            switch (g.type) {
                case HXR_GEOM_PLANE: lo[0] = lo[2] = -g.p[1]; hi[0] = hi[2] = g.p[1]; lo[1] = hi[1] = g.p[0]; break;
                case HXR_GEOM_SPHERE: for (int k = 0; k < 3; k++) { lo[k] = g.p[k] - g.p[3]; hi[k] = g.p[k] + g.p[3]; } break;
                case HXR_GEOM_CUBE: for (int k = 0; k < 3; k++) { lo[k] = g.p[k] - g.p[3]; hi[k] = g.p[k] + g.p[3]; } break;
                case HXR_GEOM_MESH: for (int k = 0; k < 3; k++) { lo[k] = s.meshes[g.a].bbox_min[k]; hi[k] = s.meshes[g.a].bbox_max[k]; } break;
                default: bounded = false; break;  // CSG, heightfield: always tested
            }
            for (int k = 0; k < 3 && bounded; k++) bounded = std::isfinite(lo[k]) && std::isfinite(hi[k]) && std::fabs(lo[k]) < 1e30 && std::fabs(hi[k]) < 1e30;
            if (!bounded) continue;
            const hxr_transform& T = s.nodes[i].T;
            double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
            for (int c = 0; c < 8; c++) {
                const double p[3] = {(c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2]};
                for (int k = 0; k < 3; k++) {
                    const double w = p[0] * T.m[k] + p[1] * T.m[3 + k] + p[2] * T.m[6 + k] + T.offset[k];
                    mn[k] = std::min(mn[k], w);
                    mx[k] = std::max(mx[k], w);
                }
            }
            for (int k = 0; k < 3; k++) {
                // margin: the intersectors' own 1e-6 tolerances (scaled by the transform) plus rounding
                const double mag = std::max(std::fabs(mn[k]), std::fabs(mx[k]));
                double scale = 0;
                for (int r = 0; r < 3; r++) scale += std::fabs(T.m[3 * r + k]);
                const double pad = 1e-5 * (1.0 + scale) + 1e-9 * mag;
                b[k] = mn[k] - pad;
                b[3 + k] = mx[k] + pad;
            }
        }
        {
            // float copies of the boxes, rounded outward: the pre-test of the walk kernel (walked nodes) and of the inline node loops
            std::vector<int32_t> inlineNodes;
            for (int i = 0; i < s.n_nodes; i++)
                if (slot[i] < 0) inlineNodes.push_back(i);
            auto floatBoxes = [&](const std::vector<int32_t>& nodes) {
                std::vector<float> fb(std::max<size_t>(1, nodes.size()) * 6, 0.0f);
                for (size_t sl = 0; sl < nodes.size(); sl++)
                    for (int k = 0; k < 3; k++) {
                        const double lo = box[(size_t)nodes[sl] * 6 + k], hi = box[(size_t)nodes[sl] * 6 + 3 + k];
                        fb[sl * 6 + k] = lo < -3e38 ? -INFINITY : std::nextafter((float)lo, -INFINITY);
                        fb[sl * 6 + 3 + k] = hi > 3e38 ? INFINITY : std::nextafter((float)hi, INFINITY);
                    }
                return fb;
            };
            const std::vector<float> bb = floatBoxes(bigNodes), ib = floatBoxes(inlineNodes);
            m_scene.big_box = uploadArray(bb.data(), bb.size());
            m_scene.inline_box = uploadArray(ib.data(), ib.size());
            m_scene.inline_nodes = uploadArray(inlineNodes.data(), inlineNodes.size());
            if (!m_scene.big_box || !m_scene.inline_box || !m_scene.inline_nodes) return oom();
            m_scene.n_inline = (int)inlineNodes.size();
        }
        m_scene.use_node_box = getenv("HXR_NO_NODE_BOX") ? 0 : 1;
        m_scene.simple_inline = 1;
        for (int i = 0; i < s.n_nodes; i++) {
            if (slot[i] >= 0) continue;
            const hxr_geometry& g = s.geometries[s.nodes[i].geom];
            const bool simple = g.type == HXR_GEOM_PLANE || g.type == HXR_GEOM_SPHERE || g.type == HXR_GEOM_CUBE ||
                                (g.type == HXR_GEOM_MESH && dm[g.a].brute);
            if (!simple) m_scene.simple_inline = 0;
        }
    }
    m_scene.walk_packed = s.n_meshes > 0 ? 1 : 0;
    for (int i = 0; i < s.n_meshes; i++)
        if (!dm[i].tri_pk) m_scene.walk_packed = 0;
    m_scene.n_nodes = s.n_nodes;
    m_scene.n_lights = s.n_lights;
    m_scene.has_env = s.has_environment;
    for (int i = 0; i < 6; i++) m_scene.env_images[i] = s.has_environment ? s.env_images[i] : 0;
    m_scene.settings = s.settings;

    // worst-case fan-out of one Whitted shading item, used to size shade launches
    long lightSamples = 0;
    for (int i = 0; i < s.n_lights; i++) lightSamples += s.lights[i].type == HXR_LIGHT_RECT ? (long)s.lights[i].xsubd * s.lights[i].ysubd : 1;
    std::function<void(int, long&, long&)> fan = [&](int si, long& sh, long& ch) {
        const hxr_shader& S = s.shaders[si];
        switch (S.type) {
            case HXR_SHADER_LAMBERT: case HXR_SHADER_PHONG: sh += lightSamples; break;
            case HXR_SHADER_REFLECTION: ch += S.f0 < 1.0f ? S.i0 : 1; break;
            case HXR_SHADER_REFRACTION: ch += 1; break;
            case HXR_SHADER_LAYERED:
                for (int l = 0; l < S.n_layers; l++) fan(s.layers[S.first_layer + l].shader, sh, ch);
                break;
            default: break;
        }
    };
    m_maxShadowPerHit = 1;
    m_maxChildrenPerHit = 1;
    for (int i = 0; i < s.n_shaders; i++) {
        long sh = 0, ch = 0;
        fan(i, sh, ch);
        m_maxShadowPerHit = (int)std::max<long>(m_maxShadowPerHit, sh);
        m_maxChildrenPerHit = (int)std::max<long>(m_maxChildrenPerHit, ch);
    }
    {
        // the scene struct itself, in device memory (DScene::self)
        DScene* d = (DScene*)keep(dev::alloc(m_dev, sizeof(DScene)));
        if (!d) return oom();
        m_scene.self = d;
        if (!dev::upload(m_dev, d, &m_scene, sizeof(DScene))) return oom();
    }
    m_haveScene = true;
    return HXR_OK;
}

int Renderer::setCamera(const hxr_camera* cam)
{
    if (!cam) return fail(HXR_ERR_INVALID, "null camera");
    m_scene.cam = *cam;
    m_haveCamera = true;
    return HXR_OK;
}

int Renderer::accelInfo(int mesh, hxr_accel_info* out) const
{
    if (!out || mesh < 0 || mesh >= (int)m_accel.size()) return HXR_ERR_INVALID;
    *out = m_accel[mesh];
    return HXR_OK;
}

// ------------------------------------------------------------------------------ queues
void Renderer::freeQueues()
{
    for (int i = 0; i < 2; i++) {
        dev::free_(m_dev, m_qg[i]); m_qg[i] = nullptr;
        dev::free_(m_dev, m_qa[i]); m_qa[i] = nullptr;
    }
    dev::free_(m_dev, m_sg); m_sg = nullptr;
    dev::free_(m_dev, m_sa); m_sa = nullptr;
    dev::free_(m_dev, m_cand); m_cand = nullptr;
    dev::free_(m_dev, m_scand); m_scand = nullptr;
    dev::free_(m_dev, m_entry); m_entry = nullptr;
    dev::free_(m_dev, m_sentry); m_sentry = nullptr;
    dev::free_(m_dev, m_ovfList); m_ovfList = nullptr;
    dev::free_(m_dev, m_ovfListS); m_ovfListS = nullptr;
    dev::free_(m_dev, m_hits); m_hits = nullptr;
    dev::free_(m_dev, m_visible); m_visible = nullptr;
    m_cap = m_shadowCap = 0;
}

// Per ray in flight: 2 x (64 + 24) B of ray queues, 64 + 16 B of shadow queue, 16 B of candidate record (one buffer serves
// the closest-hit walk and, after its rays are shaded, the shadow walk). Nothing here scales with the number of meshes.
bool Renderer::ensureQueues()
{
    const uint32_t cap = m_cfg.queue_capacity ? (uint32_t)std::min<uint64_t>(m_cfg.queue_capacity, 1u << 30) : (8u << 20);
    // path tracing issues at most one shadow ray per hit; a Whitted hit issues one per light sample and shader layer
    // (a shade launch is cut into chunks whose shadow rays fit)
    const uint64_t wantShadow = m_scene.settings.gi ? cap : std::max<uint64_t>((uint64_t)cap * 2, (uint64_t)m_maxShadowPerHit * 4096);
    const uint32_t shadowCap = (uint32_t)std::min<uint64_t>(1u << 30, std::max<uint64_t>(wantShadow, (uint64_t)m_maxShadowPerHit));
    if (m_qg[0] && cap == m_cap && shadowCap == m_shadowCap) return true;
    freeQueues();
    m_cap = cap;
    m_shadowCap = shadowCap;
    for (int i = 0; i < 2; i++) {
        m_qg[i] = (RayGeom*)dev::alloc(m_dev, (size_t)cap * sizeof(RayGeom));
        m_qa[i] = (RayAux*)dev::alloc(m_dev, (size_t)cap * sizeof(RayAux));
    }
    m_sg = (RayGeom*)dev::alloc(m_dev, (size_t)shadowCap * sizeof(RayGeom));
    m_sa = (ShadowAux*)dev::alloc(m_dev, (size_t)shadowCap * sizeof(ShadowAux));
    m_cand = (CandRec*)dev::alloc(m_dev, (size_t)std::max(cap, shadowCap) * sizeof(CandRec));
    m_entry = (MeshEntry*)dev::alloc(m_dev, (size_t)cap * sizeof(MeshEntry));
    m_sentry = (MeshEntry*)dev::alloc(m_dev, (size_t)shadowCap * sizeof(MeshEntry));
    m_ovfList = (OverflowEntry*)dev::alloc(m_dev, (size_t)std::max(cap, shadowCap) * sizeof(OverflowEntry));
    // the two lanes of drain(): the shadow chain needs its own candidate records and overflow list (HXR_NO_OVERLAP=1: one lane)
    m_overlap = !getenv("HXR_NO_OVERLAP");
    if (m_overlap) {
        m_scand = (CandRec*)dev::alloc(m_dev, (size_t)shadowCap * sizeof(CandRec));
        m_ovfListS = (OverflowEntry*)dev::alloc(m_dev, (size_t)shadowCap * sizeof(OverflowEntry));
        if (!m_scand || !m_ovfListS) m_overlap = false;  // (short of memory: one lane, shared buffers)
    }
    if (!m_counters) m_counters = (uint32_t*)dev::alloc(m_dev, C_NCOUNTERS * sizeof(uint32_t));
    if (!m_totals) m_totals = (dev::FrameTotals*)dev::alloc(m_dev, sizeof(dev::FrameTotals));
    if (!m_trav) m_trav = (TravCounters*)dev::alloc(m_dev, sizeof(TravCounters));
    if (!m_qg[0] || !m_qg[1] || !m_qa[0] || !m_qa[1] || !m_sg || !m_sa || !m_cand || !m_entry || !m_sentry || !m_ovfList || !m_counters || !m_totals || !m_trav) {
        m_err = "queue allocation failed (out of device memory: lower hxr_config.queue_capacity)";
        freeQueues();
        return false;
    }
    dev::zero(m_dev, m_counters, C_NCOUNTERS * sizeof(uint32_t));
    dev::zero(m_dev, m_totals, sizeof(dev::FrameTotals));
    dev::zero(m_dev, m_trav, sizeof(TravCounters));
    return true;
}

dev::WalkBuffers Renderer::walkBuffers(CandRec* cand, bool shadow) const
{
    dev::WalkBuffers wb;
    wb.cand = cand;
    wb.head = m_counters + (shadow ? C_HEAD_B : C_HEAD_A);
    wb.ovf_list = shadow && m_ovfListS ? m_ovfListS : m_ovfList;
    wb.ovf_count = m_counters + (shadow ? C_OVF_B : C_OVF_A);
    wb.fetch = m_counters + (shadow ? C_FETCH_B : C_FETCH_A);
    return wb;
}
RayQueue Renderer::queue(int i) const
{
    RayQueue q;
    q.geom = m_qg[i];
    q.aux = m_qa[i];
    q.entry = m_entry;
    q.count = m_counters + (i ? C_Q1 : C_Q0);
    q.cap = m_cap;
    return q;
}
ShadowQueue Renderer::shadowQueue() const
{
    ShadowQueue q;
    q.geom = m_sg;
    q.aux = m_sa;
    q.entry = m_sentry;
    q.count = m_counters + C_SHADOW;
    q.cap = m_shadowCap;
    return q;
}

uint32_t Renderer::readCount(const uint32_t* dptr)
{
    uint32_t v = 0;
    dev::download(m_dev, &v, dptr, sizeof v);
    return v;
}

// One bounce = [setup(closest)] -> walk(closest) -> shade -> setup(shadow) -> walk(shadow) -> resolve(shadow). For big waves the
// counts stay on the device: the host enqueues max_depth + 1 bounces (a ray's depth grows by one per bounce and the guard stops
// it at max_depth) without reading anything back; levels that turn out empty cost a few near-empty launches. Small frames (at
// most kSmallWave primary rays: every 1080p Whitted frame) are bound by launch latency instead, and there one 4-byte read-back
// per level pays for itself: levels nobody reaches are not launched (simple.hexray 0.97 -> 0.74 ms, kdtree_test 3.02 -> 1.94 ms). Two bounds per level: `bound`, a true upper bound of
// the level's population (what the shade launches must cover), and `grid`, a realistic one that only sizes grids (every kernel
// loops over the device-side count; ray trees die much faster than maxChildrenPerHit ^ level grows).
// Two lanes: the shadow chain of bounce L (aux lane) shares no buffer with the closest-hit chain of bounce L + 1 (main lane),
// so they are queued side by side; the main lane waits for the aux lane only before the next shade launch, which refills the
// shadow queue. Big waves fill the GPU either way; small frames, whose launches are latency-bound, overlap.
void Renderer::drain(const FrameParams& fp, float* accum, uint32_t nPrimary, hxr_stats& st)
{
    int cur = 0;
    const uint32_t perHit = fp.gi ? 1u : (uint32_t)m_maxShadowPerHit;
    const uint32_t chunk = std::max<uint32_t>(1, m_shadowCap / perHit);
    const uint64_t fan = fp.gi ? 1u : (uint64_t)std::max(1, m_maxChildrenPerHit);
    TravCounters* cnt = m_countTraversal ? m_trav : nullptr;
    const bool overlap = m_overlap && !m_oneLane && m_scand && m_ovfListS;
    bool auxBusy = false;
    // Frames of at most this many primary rays are bound by launch latency, not by throughput (a 1080p Whitted frame is 2 Mi rays)
    static const uint32_t kSmallWave = getenv("HXR_SMALL_WAVE") ? (uint32_t)atol(getenv("HXR_SMALL_WAVE")) : (1u << 22);
    const bool smallWave = nPrimary <= kSmallWave;
    uint64_t bound = nPrimary;
    for (int level = 0; level <= fp.max_depth && bound > 0; level++) {
        const uint32_t n = (uint32_t)std::min<uint64_t>(bound, m_cap);
        const uint32_t grid = (uint32_t)std::min<uint64_t>(n, (uint64_t)nPrimary * 2);
        const RayQueue q = queue(cur);
        dev::zero(m_dev, m_counters + (cur ? C_Q0 : C_Q1), sizeof(uint32_t));
        dev::zero(m_dev, m_counters + C_HEAD_A, 3 * sizeof(uint32_t));  // the closest-hit walk's cursor, its overflow-list count and k_finish's cursor
        if (level > 0) st.kernel_launches += dev::setup_closest(m_dev, m_scene, q, m_cand, cnt, grid);  // (level 0 arrives set up)
        st.kernel_launches += dev::walk(m_dev, m_scene, false, q.geom, q.entry, q.count, q.cap, walkBuffers(m_cand, false), m_totals, cnt, grid);
        Sinks sk;
        sk.next = queue(1 - cur);
        sk.shadow = shadowQueue();
        sk.accum = accum;
        sk.overflow = m_counters + C_OVERFLOW;
        // The shadow walk has its own candidate buffer when it may run beside the next closest-hit walk, or when the level is
        // shaded in chunks (Whitted hits with many light samples: later chunks still need their closest-hit records).
        CandRec* scand = overlap ? m_scand : m_cand;
        if (!overlap && n > chunk && m_scene.n_big) {
            if (!m_scand) m_scand = (CandRec*)dev::alloc(m_dev, (size_t)m_shadowCap * sizeof(CandRec));
            if (!m_scand) { m_err = "shadow candidate buffer allocation failed"; m_allocFailed = true; return; }
            scand = m_scand;
        }
        const bool side = overlap && n <= chunk;  // one shade launch covers the level: its shadow chain goes to the aux lane
        for (uint32_t b = 0; b < n; b += chunk) {
            const uint32_t e = (uint32_t)std::min<uint64_t>(n, (uint64_t)b + chunk);
            if (auxBusy) { dev::join(m_dev); auxBusy = false; }  // the previous shadow chain is done with the shadow queue
            dev::zero(m_dev, m_counters + C_SHADOW, sizeof(uint32_t));
            st.kernel_launches += dev::shade(m_dev, m_scene, fp, q, m_cand, b, e, sk, m_totals, cnt);
            const uint32_t ns = (uint32_t)std::min<uint64_t>(m_shadowCap, (uint64_t)std::min<uint64_t>(e - b, grid) * perHit);
            if (side) { dev::fork(m_dev); dev::lane(m_dev, 1); }
            dev::zero(m_dev, m_counters + C_HEAD_B, 3 * sizeof(uint32_t));
            st.kernel_launches += dev::setup_shadow(m_dev, m_scene, sk.shadow, scand, accum, cnt, ns);
            st.kernel_launches += dev::walk(m_dev, m_scene, true, m_sg, m_sentry, m_counters + C_SHADOW, m_shadowCap, walkBuffers(scand, true), m_totals, cnt, ns);
            st.kernel_launches += dev::resolve_shadow(m_dev, m_scene, sk.shadow, scand, accum, nullptr, m_totals, cnt, ns);
            if (side) { dev::lane(m_dev, 0); auxBusy = true; }
        }
        if (smallWave && level < fp.max_depth) {
            // small frames: the next level's population is worth one read-back (the main lane has just been given this level's
            // shade launch; the shadow chain keeps running on the aux lane) - levels nobody reaches are not launched at all
            bound = std::min<uint64_t>(readCount(m_counters + (cur ? C_Q0 : C_Q1)), m_cap);
        } else {
            bound = std::min<uint64_t>(bound * fan, m_cap);
        }
        cur = 1 - cur;
    }
    if (auxBusy) dev::join(m_dev);
}

// ------------------------------------------------------------------------------ frames
void Renderer::framePlan(const hxr_render_params& p, bool& mc, int& spp) const
{
    // what render() would pick (src/main.cpp:421-425)
    int raysPerPixel = 0;
    if (m_scene.cam.dof) raysPerPixel = m_scene.cam.num_samples;
    if (m_scene.settings.gi) raysPerPixel = std::max(raysPerPixel, m_scene.settings.num_paths);
    mc = raysPerPixel > 0;
    if (p.mode == HXR_MODE_WHITTED) mc = false;
    if (p.mode == HXR_MODE_MONTECARLO) mc = true;
    spp = p.spp > 0 ? p.spp : std::max(raysPerPixel, 1);
}

void Renderer::frameSize(const hxr_render_params& p, int& W, int& H) const
{
    W = p.width > 0 ? p.width : m_scene.settings.frame_width;
    H = p.height > 0 ? p.height : m_scene.settings.frame_height;
}

int Renderer::render(const hxr_render_params& p, float* hostOut, void* devOut, hxr_stats* stats, bool noOutput)
{
    if (!m_haveScene || !m_haveCamera) return fail(HXR_ERR_INVALID, "render: scene and camera must be set first");
    if (!hostOut && !devOut && !noOutput) return fail(HXR_ERR_INVALID, "render: no output buffer");
    m_noOutput = noOutput;
    if (!ensureQueues()) return HXR_ERR_CUDA;
    uint32_t batch = m_maxChildrenPerHit > 1 ? m_cap / 4 : m_cap;
    for (int attempt = 0; attempt < 6; attempt++) {
        bool overflow = false;
        int rc = renderOnce(p, hostOut, devOut, stats, std::max<uint32_t>(batch, 1024), overflow);
        if (rc != HXR_OK) return rc;
        if (!overflow) return HXR_OK;
        batch /= 4;  // a ray tree outgrew the queue: redo the frame with smaller primary batches
    }
    return fail(HXR_ERR_OVERFLOW, "ray queue overflow even at the smallest primary batch; raise hxr_config.queue_capacity");
}

int Renderer::renderOnce(const hxr_render_params& p, float* hostOut, void* devOut, hxr_stats* stats, uint32_t primaryBatch, bool& overflow)
{
    hxr_stats st;
    memset(&st, 0, sizeof st);
    int W, H;
    frameSize(p, W, H);
    if (W <= 0 || H <= 0 || (uint64_t)W * H > (1ull << 31)) return fail(HXR_ERR_INVALID, "bad frame size");
    const size_t nPix = (size_t)W * H;
    bool mc;
    int spp;
    framePlan(p, mc, spp);
    const int shardCount = p.shard_count > 1 ? p.shard_count : 1;
    const int shardIndex = shardCount > 1 ? p.shard_index : 0;
    if (shardIndex < 0 || shardIndex >= shardCount) return fail(HXR_ERR_INVALID, "bad shard index");
    const bool wantAA = p.want_aa >= 0 ? p.want_aa != 0 : m_scene.settings.want_aa != 0;
    const hxr_settings savedSettings = m_scene.settings;
    if (p.max_depth >= 0) m_scene.settings.max_trace_depth = p.max_depth;
    m_countTraversal = (p.flags & HXR_RENDER_COUNT_TRAVERSAL) != 0;
    m_oneLane = (p.flags & HXR_RENDER_ONE_LANE) != 0;
    dev::clear_error(m_dev);
    m_allocFailed = false;

    FrameParams fp;
    memset(&fp, 0, sizeof fp);
    fp.W = W;
    fp.H = H;
    fp.max_depth = m_scene.settings.max_trace_depth;
    fp.gi = mc && m_scene.settings.gi;
    fp.montecarlo = mc;
    fp.seed = p.seed;

    if (m_accumPixels < nPix) {
        dev::free_(m_dev, m_accum);
        m_accum = (float*)dev::alloc(m_dev, nPix * 3 * sizeof(float));
        m_accumPixels = m_accum ? nPix : 0;
        if (!m_accum) { m_scene.settings = savedSettings; return fail(HXR_ERR_CUDA, "framebuffer allocation failed"); }
    }
    // stereo anaglyph (src/main.cpp:234-248): every sample is traced once per eye; the eyes accumulate separately
    // and are mixed (a linear map) into the frame
    const int eyes = m_scene.cam.stereo_separation != 0.0 ? 2 : 1;
    float* eyeBuf[2] = {m_accum, nullptr};
    if (eyes == 2) {
        if (m_eyePixels < nPix) {
            for (int e = 0; e < 2; e++) { dev::free_(m_dev, m_eye[e]); m_eye[e] = (float*)dev::alloc(m_dev, nPix * 3 * sizeof(float)); }
            m_eyePixels = (m_eye[0] && m_eye[1]) ? nPix : 0;
            if (!m_eyePixels) { m_scene.settings = savedSettings; return fail(HXR_ERR_CUDA, "stereo buffer allocation failed"); }
        }
        eyeBuf[0] = m_eye[0];
        eyeBuf[1] = m_eye[1];
    }
    auto setEye = [&](int e) {
        fp.stereo_offset = eyes == 2 ? (e == 0 ? -0.5f : 0.5f) : 0.0f;
        fp.stream0 = 1u + (uint32_t)e;
    };
    setEye(0);
    if (mc && m_scene.cam.dof && m_scene.cam.auto_focus) {
        // autofocus: distance of the closest NODE along the centre ray (src/main.cpp:350-362)
        hxr_ray r;
        Ray cr;
        {
            hxr_camera c = m_scene.cam;
            c.dof = 0;
            cr = camera_ray(c, W, H, W * 0.5, H * 0.5, 0, 0, 0);
        }
        r.start[0] = cr.o.x; r.start[1] = cr.o.y; r.start[2] = cr.o.z;
        r.dir[0] = cr.d.x; r.dir[1] = cr.d.y; r.dir[2] = cr.d.z;
        r.depth = 0; r.flags = 0;
        hxr_hit h;
        // note: like raycast, this also sees lights; a light in front of the geometry hides it
        if (traceClosest(&r, 1, &h) == HXR_OK && h.status == 0) m_scene.cam.focal_plane_dist = h.dist;
    }
    dev::prof_reset(m_dev);
    dev::Timer* tm = dev::timer_create(m_dev);
    dev::timer_start(m_dev, tm);
    dev::zero(m_dev, m_accum, nPix * 3 * sizeof(float));
    for (int e = 0; e < eyes && eyes == 2; e++) dev::zero(m_dev, eyeBuf[e], nPix * 3 * sizeof(float));
    dev::zero(m_dev, m_counters + C_OVERFLOW, sizeof(uint32_t));
    dev::zero(m_dev, m_totals, sizeof(dev::FrameTotals));
    dev::zero(m_dev, m_trav, sizeof(TravCounters));

    if (mc) {
        const uint32_t nMine = (uint32_t)((spp - shardIndex + shardCount - 1) / shardCount);
        fp.sample_stride = (uint32_t)shardCount;
        if (nPix <= primaryBatch) {
            const uint32_t sppPass = std::max<uint32_t>(1, (uint32_t)(primaryBatch / nPix));
            for (uint32_t k0 = 0; k0 < nMine; k0 += sppPass) {
                const uint32_t k = std::min(sppPass, nMine - k0);
                fp.sample_base = (uint32_t)shardIndex + k0 * (uint32_t)shardCount;
                const uint32_t items = (uint32_t)(nPix * k);
                for (int e = 0; e < eyes; e++) {
                    setEye(e);
                    st.kernel_launches += dev::gen_primary(m_dev, m_scene, fp, nullptr, nullptr, 0, items, k, queue(0), m_cand);
                    drain(fp, eyeBuf[e], items, st);
                }
            }
        } else {
            for (uint32_t k0 = 0; k0 < nMine; k0++) {
                fp.sample_base = (uint32_t)shardIndex + k0 * (uint32_t)shardCount;
                for (size_t first = 0; first < nPix; first += primaryBatch) {
                    const uint32_t items = (uint32_t)std::min<size_t>(primaryBatch, nPix - first);
                    for (int e = 0; e < eyes; e++) {
                        setEye(e);
                        st.kernel_launches += dev::gen_primary(m_dev, m_scene, fp, nullptr, nullptr, (uint32_t)first, items, 1, queue(0), m_cand);
                        drain(fp, eyeBuf[e], items, st);
                    }
                }
            }
        }
        st.spp_done = nMine;
        if (eyes == 2) st.kernel_launches += dev::stereo_mix(m_dev, m_accum, eyeBuf[0], eyeBuf[1], nPix);
        if (shardCount == 1) st.kernel_launches += dev::scale_all(m_dev, m_accum, nPix * 3, 1.0f / (float)spp);
    } else {
        // pass 1: one ray through every pixel corner. Row shards own rows y with (y/16) % count == index
        // and also render a one-row halo around each owned band so that pass 2 sees all 8 neighbours.
        fp.sample_base = 0;
        fp.sample_stride = 1;
        auto owned = [&](int y) { return shardCount == 1 || ((y / HXR_ROW_BAND) % shardCount) == shardIndex; };
        auto needed = [&](int y) { return owned(y) || (y > 0 && owned(y - 1)) || (y + 1 < H && owned(y + 1)); };
        int y = 0;
        while (y < H) {
            if (!needed(y)) { y++; continue; }
            int y1 = y;
            while (y1 < H && needed(y1) && (size_t)(y1 - y + 1) * W <= std::max<size_t>(primaryBatch, W)) y1++;
            size_t first = (size_t)y * W, count = (size_t)(y1 - y) * W;
            for (size_t off = 0; off < count; off += primaryBatch) {
                const uint32_t items = (uint32_t)std::min<size_t>(primaryBatch, count - off);
                for (int e = 0; e < eyes; e++) {
                    setEye(e);
                    st.kernel_launches += dev::gen_primary(m_dev, m_scene, fp, nullptr, nullptr, (uint32_t)(first + off), items, 1, queue(0), m_cand);
                    drain(fp, eyeBuf[e], items, st);
                }
            }
            y = y1;
        }
        if (eyes == 2) st.kernel_launches += dev::stereo_mix(m_dev, m_accum, eyeBuf[0], eyeBuf[1], nPix);  // what detectAApixels looks at
        if (wantAA) {
            if (m_aaCap < nPix) {
                dev::free_(m_dev, m_aaList);
                dev::free_(m_dev, m_aaMask);
                m_aaList = (uint32_t*)dev::alloc(m_dev, nPix * sizeof(uint32_t));
                m_aaMask = (uint8_t*)dev::alloc(m_dev, nPix);
                m_aaCap = (m_aaList && m_aaMask) ? nPix : 0;
                if (!m_aaCap) { m_scene.settings = savedSettings; dev::timer_destroy(m_dev, tm); return fail(HXR_ERR_CUDA, "AA buffer allocation failed"); }
            }
            dev::zero(m_dev, m_counters + C_AA, sizeof(uint32_t));
            st.kernel_launches += dev::aa_detect(m_dev, m_accum, W, H, shardIndex, shardCount, m_aaList, m_counters + C_AA, m_aaMask);
            const uint32_t nAA = readCount(m_counters + C_AA);  // the frame's one host read-back: sizes the second pass
            st.aa_pixels = nAA;
            fp.sample_base = 1;
            const uint32_t pixPerBatch = std::max<uint32_t>(1, primaryBatch / 4);
            for (uint32_t first = 0; first < nAA; first += pixPerBatch) {
                const uint32_t np = std::min(pixPerBatch, nAA - first);
                for (int e = 0; e < eyes; e++) {
                    setEye(e);
                    st.kernel_launches += dev::gen_primary(m_dev, m_scene, fp, m_aaList + first, nullptr, 0, np * 4, 4, queue(0), m_cand);
                    drain(fp, eyeBuf[e], np * 4, st);
                }
            }
            if (eyes == 2) st.kernel_launches += dev::stereo_mix(m_dev, m_accum, eyeBuf[0], eyeBuf[1], nPix);
            st.kernel_launches += dev::scale_listed(m_dev, m_accum, m_aaList, m_counters + C_AA, (uint32_t)nPix, 1.0f / 5);
        }
        if (shardCount > 1) {
            // drop the halo rows: the caller sums the shards
            for (int yy = 0; yy < H; yy++)
                if (!owned(yy) && needed(yy)) dev::zero(m_dev, m_accum + (size_t)yy * W * 3, (size_t)W * 3 * sizeof(float));
        }
    }
    dev::timer_stop(m_dev, tm);
    const bool ov = readCount(m_counters + C_OVERFLOW) != 0;
    st.render_ms = dev::timer_ms(m_dev, tm);
    dev::timer_destroy(m_dev, tm);
    m_scene.settings = savedSettings;
    if (m_allocFailed) return HXR_ERR_CUDA;
    if (dev::failed(m_dev)) return fail(HXR_ERR_CUDA, std::string("render failed: ") + dev::last_error(m_dev));
    if (ov) { overflow = true; return HXR_OK; }

    m_lastW = W;
    m_lastH = H;
    bool ok = true;
    if (devOut) ok = ok && dev::copy_d2d(m_dev, devOut, m_accum, nPix * 3 * sizeof(float));
    if (hostOut) ok = ok && dev::download(m_dev, hostOut, m_accum, nPix * 3 * sizeof(float));
    if (!ok) return fail(HXR_ERR_CUDA, std::string("result copy failed: ") + dev::last_error(m_dev));
    {
        dev::FrameTotals ft;
        TravCounters tc;
        dev::download(m_dev, &ft, m_totals, sizeof ft);
        dev::download(m_dev, &tc, m_trav, sizeof tc);
        st.rays_closest = ft.rays_closest;
        st.rays_shadow = ft.rays_shadow;
        st.cand_overflow = ft.cand_overflow;
        st.kd_inner = tc.kd_inner;
        st.kd_leaves = tc.kd_leaves;
        st.tri_tests = tc.tri_tests;
        st.mesh_queries = tc.mesh_queries;
        double ms[dev::PROF_NCAT];
        uint64_t ln[dev::PROF_NCAT];
        dev::prof_collect(m_dev, ms, ln);
        st.trace_closest_ms = ms[dev::PROF_WALK_CLOSEST];
        st.trace_shadow_ms = ms[dev::PROF_WALK_SHADOW] + ms[dev::PROF_SHADOW_RESOLVE];
        st.shade_ms = ms[dev::PROF_SHADE];
        st.other_ms = ms[dev::PROF_OTHER] + ms[dev::PROF_GEN];
        st.setup_ms = ms[dev::PROF_SETUP];
        st.finish_ms = ms[dev::PROF_FINISH];
        st.trace_closest_launches = ln[dev::PROF_WALK_CLOSEST];
        st.trace_shadow_launches = ln[dev::PROF_WALK_SHADOW] + ln[dev::PROF_SHADOW_RESOLVE];
        st.walk_ms = ms[dev::PROF_WALK_CLOSEST] + ms[dev::PROF_WALK_SHADOW];
        st.walk_launches = ln[dev::PROF_WALK_CLOSEST] + ln[dev::PROF_WALK_SHADOW];
        st.shadow_resolve_ms = ms[dev::PROF_SHADOW_RESOLVE];
        st.gen_ms = ms[dev::PROF_GEN];
    }
    if (stats) *stats = st;
    if (!dev::sync(m_dev)) return fail(HXR_ERR_CUDA, dev::last_error(m_dev));
    return HXR_OK;
}

int Renderer::resolveDevice(void* d_rgb, int W, int H, int spp)
{
    if (!d_rgb || W <= 0 || H <= 0 || spp <= 0) return fail(HXR_ERR_INVALID, "resolve: bad arguments");
    dev::scale_all(m_dev, (float*)d_rgb, (size_t)W * H * 3, 1.0f / (float)spp);
    if (!dev::sync(m_dev)) return fail(HXR_ERR_CUDA, dev::last_error(m_dev));
    return HXR_OK;
}

// The screenshot (takeScreenshot -> Bitmap::saveBMP, src/sdl.cpp:103-116, src/bitmap.cpp:202-240) straight from device
// memory: the 8-bit conversion runs on the GPU and a quarter of the float frame's bytes cross PCIe.
int Renderer::saveFrameBmp(const void* d_rgb, int W, int H, const char* path)
{
    if (!path) return fail(HXR_ERR_INVALID, "save_frame: null path");
    if (!d_rgb) { d_rgb = m_accum; W = m_lastW; H = m_lastH; }
    if (!d_rgb || W <= 0 || H <= 0) return fail(HXR_ERR_INVALID, "save_frame: no frame");
    if (!m_srgbLut) {
        uint8_t lut[4097];
        for (int i = 0; i <= 4096; i++) lut[i] = (uint8_t)host::convertTo8bit_sRGB(i / 4096.0f);
        m_srgbLut = (uint8_t*)dev::alloc(m_dev, sizeof lut);
        if (!m_srgbLut || !dev::upload(m_dev, m_srgbLut, lut, sizeof lut)) return fail(HXR_ERR_CUDA, dev::last_error(m_dev));
    }
    int rowsz = W * 3;
    if (rowsz % 4) rowsz += 4 - (rowsz % 4);
    const size_t bytes = (size_t)rowsz * H;
    uint8_t* dout = (uint8_t*)dev::alloc(m_dev, bytes);
    if (!dout) return fail(HXR_ERR_CUDA, "save_frame: out of device memory");
    std::vector<uint8_t> rows(bytes);
    dev::to_bmp_rows(m_dev, (const float*)d_rgb, W, H, rowsz, m_srgbLut, dout);
    const bool ok = dev::download(m_dev, rows.data(), dout, bytes);
    dev::free_(m_dev, dout);
    if (!ok) return fail(HXR_ERR_CUDA, dev::last_error(m_dev));
    if (!host::writeBmpFile(path, W, H, rowsz, rows.data())) return fail(HXR_ERR_IO, std::string("cannot write ") + path);
    return HXR_OK;
}

// The EXR screenshot (Bitmap::saveEXR, src/bitmap.cpp:270-288: HALF RGBA, alpha 1): float -> half on the GPU, two thirds of
// the float frame's bytes cross PCIe; the file is byte-identical to hxr_save_image(".exr") of the same frame.
int Renderer::saveFrameExr(const void* d_rgb, int W, int H, const char* path)
{
    if (!path) return fail(HXR_ERR_INVALID, "save_frame: null path");
    if (!d_rgb) { d_rgb = m_accum; W = m_lastW; H = m_lastH; }
    if (!d_rgb || W <= 0 || H <= 0) return fail(HXR_ERR_INVALID, "save_frame: no frame");
    const size_t halves = (size_t)W * H * 4;
    uint16_t* dout = (uint16_t*)dev::alloc(m_dev, halves * sizeof(uint16_t));
    if (!dout) return fail(HXR_ERR_CUDA, "save_frame: out of device memory");
    std::vector<uint16_t> rows(halves);
    dev::to_exr_rows(m_dev, (const float*)d_rgb, W, H, dout);
    const bool ok = dev::download(m_dev, rows.data(), dout, halves * sizeof(uint16_t));
    dev::free_(m_dev, dout);
    if (!ok) return fail(HXR_ERR_CUDA, dev::last_error(m_dev));
    if (!exr::save_half_rows(path, W, H, rows.data())) return fail(HXR_ERR_IO, std::string("cannot write ") + path);
    return HXR_OK;
}

// ------------------------------------------------------------------------------ test hooks
// explicit rays as raw queue records (their inline part is decided on the device by setup_closest)
static void rawRays(const hxr_ray* rays, uint32_t m, std::vector<RayGeom>& geoms, std::vector<RayAux>& auxs)
{
    geoms.resize(m);
    auxs.resize(m);
    for (uint32_t i = 0; i < m; i++) {
        Ray ray;
        ray.o = ld3(rays[i].start);
        ray.d = ld3(rays[i].dir);
        ray.depth = rays[i].depth;
        ray.flags = rays[i].flags;
        geoms[i] = raw_geom(ray);
        auxs[i].w[0] = auxs[i].w[1] = auxs[i].w[2] = 1.0f;
        auxs[i].pixel = i;
        auxs[i].sample = 0;
        auxs[i].stream = 1;
    }
}

int Renderer::traceClosest(const hxr_ray* rays, size_t n, hxr_hit* hits)
{
    if (!m_haveScene) return fail(HXR_ERR_INVALID, "trace: no scene");
    if (n && (!rays || !hits)) return fail(HXR_ERR_INVALID, "trace: null buffer");
    if (!ensureQueues()) return HXR_ERR_CUDA;
    if (!m_hits) m_hits = (HitRec*)dev::alloc(m_dev, (size_t)m_cap * sizeof(HitRec));
    if (!m_hits) return fail(HXR_ERR_CUDA, "trace_closest: out of device memory");
    dev::clear_error(m_dev);
    std::vector<HitRec> recs;
    std::vector<RayGeom> geoms;
    std::vector<RayAux> auxs;
    for (size_t first = 0; first < n; first += m_cap) {
        const uint32_t m = (uint32_t)std::min<size_t>(m_cap, n - first);
        recs.resize(m);
        rawRays(rays + first, m, geoms, auxs);
        const RayQueue q = queue(0);
        dev::upload(m_dev, q.geom, geoms.data(), (size_t)m * sizeof(RayGeom));
        dev::upload(m_dev, q.count, &m, sizeof m);
        DScene full = m_scene;
        full.full_attr = 1;  // the hook reports u, v, dNdx, dNdy of every hit, whatever the node's shader reads
        dev::setup_closest(m_dev, full, q, m_cand, nullptr, m);
        dev::zero(m_dev, m_counters + C_SHADOW, (C_FETCH_B - C_SHADOW + 1) * sizeof(uint32_t));
        dev::walk(m_dev, full, false, q.geom, q.entry, q.count, q.cap, walkBuffers(m_cand, false), nullptr, nullptr, m);
        dev::hit_records(m_dev, full, q, m_cand, m_hits, m);
        if (!dev::download(m_dev, recs.data(), m_hits, (size_t)m * sizeof(HitRec)) || dev::failed(m_dev)) return fail(HXR_ERR_CUDA, dev::last_error(m_dev));
        for (uint32_t i = 0; i < m; i++) {
            const HitRec& h = recs[i];
            hxr_hit& o = hits[first + i];
            memset(&o, 0, sizeof o);
            o.status = h.node >= 0 ? 0 : 1;
            o.node = h.node;
            if (h.node >= 0) {
                o.dist = h.dist;
                o.u = h.u;
                o.v = h.v;
                for (int k = 0; k < 3; k++) { o.ip[k] = h.ip[k]; o.norm[k] = h.norm[k]; o.dndx[k] = h.dNdx[k]; o.dndy[k] = h.dNdy[k]; }
            } else {
                for (int k = 0; k < 3; k++) o.color[k] = h.color[k];
            }
        }
    }
    return HXR_OK;
}

int Renderer::traceVisible(const double* seg, size_t n, uint8_t* out)
{
    if (!m_haveScene) return fail(HXR_ERR_INVALID, "trace: no scene");
    if (n && (!seg || !out)) return fail(HXR_ERR_INVALID, "trace: null buffer");
    if (!ensureQueues()) return HXR_ERR_CUDA;
    if (!m_visible) m_visible = (uint8_t*)dev::alloc(m_dev, (size_t)m_shadowCap);
    if (!m_visible) return fail(HXR_ERR_CUDA, "trace_visible: out of device memory");
    dev::clear_error(m_dev);
    const uint32_t batchCap = m_shadowCap;
    std::vector<RayGeom> geoms;
    for (size_t first = 0; first < n; first += batchCap) {
        const uint32_t m = (uint32_t)std::min<size_t>(batchCap, n - first);
        geoms.resize(m);
        for (uint32_t i = 0; i < m; i++) {
            double D;
            const double* sg = seg + (first + i) * 6;
            const Ray ray = shadow_ray(ld3(sg), ld3(sg + 3), D);
            geoms[i] = shadow_geom(ray, D, false);
        }
        const ShadowQueue q = shadowQueue();
        dev::upload(m_dev, q.geom, geoms.data(), (size_t)m * sizeof(RayGeom));
        dev::upload(m_dev, q.count, &m, sizeof m);
        dev::setup_shadow(m_dev, m_scene, q, m_cand, nullptr, nullptr, m);
        dev::zero(m_dev, m_counters + C_HEAD_A, (C_FETCH_B - C_HEAD_A + 1) * sizeof(uint32_t));
        dev::walk(m_dev, m_scene, true, q.geom, q.entry, q.count, q.cap, walkBuffers(m_cand, true), nullptr, nullptr, m);
        dev::resolve_shadow(m_dev, m_scene, q, m_cand, nullptr, m_visible, nullptr, nullptr, m);
        if (!dev::download(m_dev, out + first, m_visible, m) || dev::failed(m_dev)) return fail(HXR_ERR_CUDA, dev::last_error(m_dev));
    }
    return HXR_OK;
}

int Renderer::traceColor(const hxr_ray* rays, size_t n, float* rgb)
{
    if (!m_haveScene) return fail(HXR_ERR_INVALID, "trace: no scene");
    if (n && (!rays || !rgb)) return fail(HXR_ERR_INVALID, "trace: null buffer");
    if (!ensureQueues()) return HXR_ERR_CUDA;
    const uint32_t batch = std::max<uint32_t>(1024, m_cap / 8);
    float* acc = (float*)dev::alloc(m_dev, (size_t)batch * 3 * sizeof(float));
    if (!acc) return fail(HXR_ERR_CUDA, "trace_color: out of device memory");
    dev::clear_error(m_dev);
    FrameParams fp;
    memset(&fp, 0, sizeof fp);
    fp.W = (int)batch;
    fp.H = 1;
    fp.max_depth = m_scene.settings.max_trace_depth;
    fp.sample_stride = 1;
    hxr_stats st;
    memset(&st, 0, sizeof st);
    std::vector<RayGeom> geoms;
    std::vector<RayAux> auxs;
    int rc = HXR_OK;
    m_countTraversal = false;
    m_allocFailed = false;
    dev::zero(m_dev, m_counters + C_OVERFLOW, sizeof(uint32_t));
    for (size_t first = 0; first < n && rc == HXR_OK; first += batch) {
        const uint32_t m = (uint32_t)std::min<size_t>(batch, n - first);
        dev::zero(m_dev, acc, (size_t)batch * 3 * sizeof(float));
        rawRays(rays + first, m, geoms, auxs);  // pixel = index in the batch, weight 1
        const RayQueue q = queue(0);
        dev::upload(m_dev, q.geom, geoms.data(), (size_t)m * sizeof(RayGeom));
        dev::upload(m_dev, q.aux, auxs.data(), (size_t)m * sizeof(RayAux));
        dev::upload(m_dev, q.count, &m, sizeof m);
        dev::setup_closest(m_dev, m_scene, q, m_cand, nullptr, m);
        drain(fp, acc, m, st);
        if (m_allocFailed) rc = HXR_ERR_CUDA;
        else if (readCount(m_counters + C_OVERFLOW)) rc = fail(HXR_ERR_OVERFLOW, "trace_color: queue overflow");
        else if (!dev::download(m_dev, rgb + first * 3, acc, (size_t)m * 3 * sizeof(float)) || dev::failed(m_dev)) rc = fail(HXR_ERR_CUDA, dev::last_error(m_dev));
    }
    dev::free_(m_dev, acc);
    return rc;
}

}  // namespace hxr
