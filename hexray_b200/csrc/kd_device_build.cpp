// kd_device_build.cpp — driver of the device KD-tree build (device/kdbuild.h has the algorithm, device/launch.h the passes).
// The tree grows one level per round; the host only reads back four numbers per level (how many nodes split, how many
// references the next level holds, how many went into leaves, the largest child) to size the next round's buffers.
// The finished binary tree comes back to the host once and is packed into the walk's blocks by the same code as a host-built
// tree (host/kdtree.cpp: packKdTree), so everything downstream - records, upload, walk - is unchanged.
#include "kd_device_build.h"
#include <algorithm>
#include <chrono>
#include <cstring>

namespace hxr {

namespace {

// device buffers of one build, freed together; `grow` keeps the content
struct Pool {
    dev::Context* c;
    std::vector<void*> all;
    bool ok = true;
    explicit Pool(dev::Context* ctx) : c(ctx) {}
    ~Pool() { for (void* p : all) dev::free_(c, p); }
    template <class T> T* get(size_t n)
    {
        void* p = dev::alloc(c, std::max<size_t>(n, 1) * sizeof(T));
        if (!p) { ok = false; return nullptr; }
        all.push_back(p);
        return (T*)p;
    }
    void drop(void* p)
    {
        if (!p) return;
        all.erase(std::remove(all.begin(), all.end(), p), all.end());
        dev::free_(c, p);
    }
    // make *buf hold at least `need` elements (`keep` of them carried over)
    template <class T> bool ensure(T*& buf, size_t& cap, size_t need, size_t keep = 0)
    {
        if (need <= cap && buf) return true;
        const size_t ncap = std::max(need + need / 4 + 1024, cap);
        T* nb = get<T>(ncap);
        if (!nb) return false;
        if (keep && buf) dev::copy_d2d(c, nb, buf, keep * sizeof(T));
        if (buf) { dev::sync(c); drop(buf); }
        buf = nb;
        cap = ncap;
        return true;
    }
};

}  // namespace

bool buildKdTreeOnDevice(dev::Context* c, const hxr_mesh& mesh, const host::KdBuildParams& params, host::KdTree& out, std::string& err)
{
    const auto t0 = std::chrono::steady_clock::now();
    host::KdBuildParams hp = params;
    host::resolveKdParams(mesh, hp);
    kdb::Params P;
    P.traversalCost = hp.traversalCost;
    P.intersectCost = hp.intersectCost;
    P.emptyBonus = hp.emptyBonus;
    P.maxLeafSize = hp.maxLeafSize;
    P.maxDepth = hp.maxDepth;
    P.binnedAbove = std::min(hp.binnedAbove, HXR_KDB_EXACT_MAX);
    const uint32_t nTris = (uint32_t)std::max(0, mesh.n_triangles);
    out = host::KdTree();
    if (nTris == 0) {
        host::packKdTree(out, mesh);
        return true;
    }
    Pool pool(c);
    auto fail = [&](const char* what) {
        err = std::string("device KD build: ") + what + (dev::failed(c) ? std::string(": ") + dev::last_error(c) : std::string());
        return false;
    };

    // the mesh: vertices and the triangles' vertex indices
    double* dVerts = pool.get<double>((size_t)mesh.n_vertices * 3);
    int32_t* dTriV = pool.get<int32_t>((size_t)nTris * 3);
    double* dTb = pool.get<double>((size_t)nTris * 6);
    if (!pool.ok) return fail("out of device memory");
    {
        std::vector<int32_t> idx((size_t)nTris * 3);
        for (uint32_t t = 0; t < nTris; t++)
            for (int k = 0; k < 3; k++) idx[(size_t)t * 3 + k] = mesh.triangles[t].v[k];
        dev::upload(c, dVerts, mesh.vertices, (size_t)mesh.n_vertices * 3 * sizeof(double));
        dev::upload(c, dTriV, idx.data(), idx.size() * sizeof(int32_t));
    }
    dev::kd_bounds(c, dVerts, dTriV, nTris, dTb);

    // level state
    uint32_t *refTri[2] = {nullptr, nullptr}, *refNode[2] = {nullptr, nullptr}, *flagL = nullptr, *flagR = nullptr;
    size_t capTri[2] = {0, 0}, capNode[2] = {0, 0}, capFL = 0, capFR = 0;
    kdb::NodeWork* work[2] = {nullptr, nullptr};
    size_t capWork[2] = {0, 0};
    kdb::Decision* dec = nullptr;
    size_t capDec = 0;
    uint32_t *childRefs = nullptr, *isSplit = nullptr, *leafRefs = nullptr, *hist = nullptr;
    size_t capCR = 0, capIS = 0, capLR = 0, capHist = 0;
    kdb::OutNode* outNodes = nullptr;
    size_t capOut = 0;
    uint32_t* leafOut = nullptr;
    size_t capLeaf = 0;
    uint32_t* dLevelMax = pool.get<uint32_t>(1);

    int cur = 0;
    uint32_t nNodes = 1, nRefs = nTris, outCount = 1, leafCount = 0, levelMax = nTris;
    if (!pool.ensure(refTri[0], capTri[0], (size_t)nTris * 3) || !pool.ensure(refNode[0], capNode[0], (size_t)nTris * 3) ||
        !pool.ensure(work[0], capWork[0], 1024) || !pool.ensure(outNodes, capOut, (size_t)nTris + 1024) ||
        !pool.ensure(leafOut, capLeaf, (size_t)nTris * 4 + 1024))
        return fail("out of device memory");
    dev::kd_iota(c, refTri[0], refNode[0], nTris);
    {
        kdb::NodeWork root;
        host::kdRootBox(mesh, root.mn, root.mx);
        root.start = 0; root.count = nTris; root.out = 0; root.bad = 0;
        dev::upload(c, work[0], &root, sizeof root);
    }
    uint32_t maxDepth = 0;
    uint64_t leaves = 0;
    for (int depth = 0; nNodes > 0; depth++) {
        if (depth > HXR_KD_MAX_DEPTH + 1) return fail("the tree does not terminate");
        const int nxt = 1 - cur;
        if (!pool.ensure(dec, capDec, nNodes) || !pool.ensure(childRefs, capCR, (size_t)nNodes + 1) || !pool.ensure(isSplit, capIS, (size_t)nNodes + 1) ||
            !pool.ensure(leafRefs, capLR, (size_t)nNodes + 1) || !pool.ensure(flagL, capFL, (size_t)nRefs + 1) || !pool.ensure(flagR, capFR, (size_t)nRefs + 1))
            return fail("out of device memory");
        const bool binned = (int)levelMax > P.binnedAbove;
        if (binned) {
            const size_t words = (size_t)nNodes * 3 * 2 * HXR_KDB_BINS;
            if (!pool.ensure(hist, capHist, words)) return fail("out of device memory");
            dev::zero(c, hist, words * sizeof(uint32_t));
            dev::kd_bin(c, P, work[cur], refTri[cur], refNode[cur], nRefs, dTb, nTris, hist);
        }
        dev::kd_choose(c, P, work[cur], nNodes, depth, binned ? hist : nullptr, refTri[cur], dTb, nTris, dec);
        dev::kd_classify(c, refTri[cur], refNode[cur], nRefs, dec, dTb, nTris, flagL, flagR);
        dev::scan_u32(c, flagL, nRefs + 1);
        dev::scan_u32(c, flagR, nRefs + 1);
        dev::zero(c, dLevelMax, sizeof(uint32_t));
        dev::kd_plan(c, work[cur], nNodes, dec, flagL, flagR, childRefs, isSplit, leafRefs, dLevelMax);
        dev::scan_u32(c, childRefs, nNodes + 1);
        dev::scan_u32(c, isSplit, nNodes + 1);
        dev::scan_u32(c, leafRefs, nNodes + 1);
        uint32_t totChild = 0, nSplit = 0, totLeaf = 0;
        dev::download(c, &totChild, childRefs + nNodes, sizeof(uint32_t));
        dev::download(c, &nSplit, isSplit + nNodes, sizeof(uint32_t));
        dev::download(c, &totLeaf, leafRefs + nNodes, sizeof(uint32_t));
        dev::download(c, &levelMax, dLevelMax, sizeof(uint32_t));
        if (dev::failed(c)) return fail("a pass failed");
        if ((uint64_t)outCount + 2ull * nSplit > 0x7FFFFFFFull || (uint64_t)leafCount + totLeaf > 0x7FFFFFFFull) return fail("the tree outgrows 31-bit references");
        if (!pool.ensure(outNodes, capOut, (size_t)outCount + 2 * (size_t)nSplit, outCount) || !pool.ensure(leafOut, capLeaf, (size_t)leafCount + totLeaf, leafCount) ||
            !pool.ensure(refTri[nxt], capTri[nxt], totChild) || !pool.ensure(refNode[nxt], capNode[nxt], totChild) ||
            !pool.ensure(work[nxt], capWork[nxt], 2 * (size_t)nSplit))
            return fail("out of device memory");
        dev::kd_emit(c, work[cur], nNodes, dec, childRefs, isSplit, leafRefs, outCount, leafCount, outNodes, work[nxt]);
        dev::kd_scatter(c, refTri[cur], refNode[cur], nRefs, work[cur], dec, flagL, flagR, childRefs, isSplit, leafRefs, leafCount, refTri[nxt], refNode[nxt],
                        leafOut);
        if (nNodes > nSplit) maxDepth = (uint32_t)depth;
        leaves += nNodes - nSplit;
        outCount += 2 * nSplit;
        leafCount += totLeaf;
        nNodes = 2 * nSplit;
        nRefs = totChild;
        cur = nxt;
    }
    // the binary tree, back on the host
    static_assert(sizeof(kdb::OutNode) == sizeof(KdNode), "OutNode mirrors KdNode");
    out.nodes.resize(outCount);
    out.leafTris.resize(leafCount);
    dev::download(c, out.nodes.data(), outNodes, (size_t)outCount * sizeof(KdNode));
    if (leafCount) dev::download(c, out.leafTris.data(), leafOut, (size_t)leafCount * sizeof(uint32_t));
    if (dev::failed(c)) return fail("reading the tree back failed");
    out.maxDepth = maxDepth;
    out.leaves = leaves;
    const auto t1 = std::chrono::steady_clock::now();
    host::packKdTree(out, mesh);
    const auto t2 = std::chrono::steady_clock::now();
    out.buildMs = std::chrono::duration<double, std::milli>(t2 - t0).count();
    out.deviceMs = std::chrono::duration<double, std::milli>(t1 - t0).count();
    return true;
}

}  // namespace hxr
