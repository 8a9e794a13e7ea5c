// kdtree.cpp — see kdtree.h. Top-down SAH build:
//   * big nodes: 32-bin SAH per axis (O(n) per node), small nodes: exact sweep over the triangle
//     bound edges (O(n log n)), both on triangle bounds clipped to the node box;
//   * a triangle goes LEFT if its bounds reach below the split (min < split) or it lies entirely
//     in the split plane, RIGHT if they reach above it (max > split). A triangle that only
//     touches the plane from one side is NOT duplicated (on grid-aligned meshes that rule alone
//     decides between ~1.5x and ~20x reference duplication); a hit exactly on the plane is still
//     found because the traversal visits both children whenever the plane parameter lies within
//     the node's [tmin, tmax] (with slack). The split is stored as float and the SAME
//     float-rounded value is used for classification here and for traversal on the device;
//   * subtrees are built in parallel (std::thread) once the refs below a node drop under a
//     grain size, then stitched into one DFS-ordered node array.
#include "kdtree.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <future>
#include <thread>

namespace hxr {
namespace host {

void resolveKdParams(const hxr_mesh& mesh, KdBuildParams& P)
{
    // experiment knobs (defaults in kdtree.h)
    if (const char* e = getenv("HXR_KD_INTERSECT_COST")) P.intersectCost = (float)atof(e);
    if (const char* e = getenv("HXR_KD_MAX_LEAF")) P.maxLeafSize = atoi(e);
    if (const char* e = getenv("HXR_KD_EMPTY_BONUS")) P.emptyBonus = (float)atof(e);
    if (P.maxDepth < 0) P.maxDepth = (int)std::lround(8 + 1.3 * std::log2((double)std::max(1, mesh.n_triangles)));
    P.maxDepth = std::min(P.maxDepth, HXR_KD_MAX_DEPTH);
}

void kdRootBox(const hxr_mesh& mesh, double mn[3], double mx[3])
{
    for (int a = 0; a < 3; a++) {
        // float-rounded outward so that float splits compare consistently with the triangle bounds
        mn[a] = std::nextafter((float)mesh.bbox_min[a], -INFINITY);
        mx[a] = std::nextafter((float)mesh.bbox_max[a], +INFINITY);
    }
}

namespace {

struct Box {
    double mn[3], mx[3];
    double area() const
    {
        const double dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};

struct TriBounds {
    double mn[3], mx[3];  // exact vertex extents
};

struct Builder {
    const hxr_mesh& mesh;
    KdBuildParams P;
    std::vector<TriBounds> tb;
    int maxDepth;

    explicit Builder(const hxr_mesh& m, const KdBuildParams& p) : mesh(m), P(p)
    {
        resolveKdParams(m, P);
        const int n = m.n_triangles;
        tb.resize(n);
        for (int i = 0; i < n; i++) {
            const hxr_triangle& t = m.triangles[i];
            double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
            for (int k = 0; k < 3; k++) {
                const double* v = m.vertices + 3 * (size_t)t.v[k];
                for (int a = 0; a < 3; a++) {
                    mn[a] = std::min(mn[a], v[a]);
                    mx[a] = std::max(mx[a], v[a]);
                }
            }
            for (int a = 0; a < 3; a++) {
                tb[i].mn[a] = mn[a];
                tb[i].mx[a] = mx[a];
            }
        }
        maxDepth = P.maxDepth;
    }

    static bool goesLeft(const TriBounds& b, int axis, float split)
    {
        const double s = (double)split;
        return b.mn[axis] < s || (b.mn[axis] == s && b.mx[axis] == s);
    }
    static bool goesRight(const TriBounds& b, int axis, float split) { return b.mx[axis] > (double)split; }

    struct Local {  // a subtree under construction
        std::vector<KdNode> nodes;
        std::vector<uint32_t> leafTris;
        uint32_t maxDepth = 0;
        uint64_t leaves = 0;
    };

    bool chooseSplit(const std::vector<uint32_t>& refs, const Box& box, int& bestAxis, float& bestSplit, float& bestCost) const
    {
        const size_t n = refs.size();
        const double invArea = 1.0 / std::max(box.area(), 1e-300);
        bestCost = INFINITY;
        bestAxis = -1;
        const double ext[3] = {box.mx[0] - box.mn[0], box.mx[1] - box.mn[1], box.mx[2] - box.mn[2]};
        auto sahCost = [&](int axis, double split, size_t nl, size_t nr) {
            const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
            const double dl = split - box.mn[axis], dr = box.mx[axis] - split;
            const double aL = 2.0 * (ext[a1] * ext[a2] + dl * (ext[a1] + ext[a2]));
            const double aR = 2.0 * (ext[a1] * ext[a2] + dr * (ext[a1] + ext[a2]));
            const double eb = (nl == 0 || nr == 0) ? P.emptyBonus : 0.0;
            return P.traversalCost + P.intersectCost * (1.0 - eb) * (aL * invArea * nl + aR * invArea * nr);
        };
        for (int axis = 0; axis < 3; axis++) {
            if (!(ext[axis] > 0)) continue;
            if ((int)n > P.binnedAbove) {
                const int B = 32;
                uint32_t startCnt[B] = {0}, endCnt[B] = {0};
                const double scale = B / ext[axis];
                for (uint32_t r : refs) {
                    double lo = std::max((double)tb[r].mn[axis], box.mn[axis]), hi = std::min((double)tb[r].mx[axis], box.mx[axis]);
                    int b0 = std::min(B - 1, std::max(0, (int)((lo - box.mn[axis]) * scale)));
                    int b1 = std::min(B - 1, std::max(0, (int)((hi - box.mn[axis]) * scale)));
                    startCnt[b0]++;
                    endCnt[b1]++;
                }
                size_t nl = 0, nr = n;
                for (int k = 1; k < B; k++) {
                    nl += startCnt[k - 1];
                    nr -= endCnt[k - 1];
                    const float split = (float)(box.mn[axis] + ext[axis] * k / B);
                    if (!((double)split > box.mn[axis] && (double)split < box.mx[axis])) continue;
                    const float c = (float)sahCost(axis, split, nl, nr);
                    if (c < bestCost) { bestCost = c; bestAxis = axis; bestSplit = split; }
                }
            } else {
                // exact sweep: candidates are the (float-rounded) triangle bound edges inside the node
                std::vector<double> mins(n), maxs(n), planar;
                std::vector<float> cand;
                cand.reserve(2 * n);
                for (size_t i = 0; i < n; i++) {
                    const TriBounds& b = tb[refs[i]];
                    mins[i] = b.mn[axis];
                    maxs[i] = b.mx[axis];
                    if (b.mn[axis] == b.mx[axis]) planar.push_back(b.mn[axis]);
                    cand.push_back((float)b.mn[axis]);
                    cand.push_back((float)b.mx[axis]);
                }
                std::sort(mins.begin(), mins.end());
                std::sort(maxs.begin(), maxs.end());
                std::sort(planar.begin(), planar.end());
                std::sort(cand.begin(), cand.end());
                cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
                for (float t : cand) {
                    const double td = (double)t;
                    if (!(td > box.mn[axis] && td < box.mx[axis])) continue;
                    const size_t below = (size_t)(std::lower_bound(mins.begin(), mins.end(), td) - mins.begin());
                    const auto pr = std::equal_range(planar.begin(), planar.end(), td);
                    const size_t nl = below + (size_t)(pr.second - pr.first);
                    const size_t nr = n - (size_t)(std::upper_bound(maxs.begin(), maxs.end(), td) - maxs.begin());
                    const float c = (float)sahCost(axis, td, nl, nr);
                    if (c < bestCost) { bestCost = c; bestAxis = axis; bestSplit = t; }
                }
            }
        }
        return bestAxis >= 0;
    }

    void makeLeaf(Local& L, uint32_t nodeIdx, const std::vector<uint32_t>& refs, int depth) const
    {
        KdNode& nd = L.nodes[nodeIdx];
        nd.kind = 3;
        nd.split = 0;
        nd.a = (uint32_t)L.leafTris.size();
        nd.b = (uint32_t)refs.size();
        L.leafTris.insert(L.leafTris.end(), refs.begin(), refs.end());
        L.leaves++;
        L.maxDepth = std::max(L.maxDepth, (uint32_t)depth);
    }

    // builds the subtree for `refs` inside `box` into L; returns its root index in L.nodes
    uint32_t build(Local& L, std::vector<uint32_t>& refs, const Box& box, int depth, int badRefines) const
    {
        const uint32_t me = (uint32_t)L.nodes.size();
        L.nodes.push_back(KdNode{0, 3, 0, 0});
        const size_t n = refs.size();
        if ((int)n <= 1 || depth >= maxDepth) { makeLeaf(L, me, refs, depth); return me; }
        int axis;
        float split, cost;
        if (!chooseSplit(refs, box, axis, split, cost)) { makeLeaf(L, me, refs, depth); return me; }
        const float leafCost = P.intersectCost * (float)n;
        if (cost > leafCost) badRefines++;
        if ((cost > 4 * leafCost && n < 16) || badRefines >= 3 || ((int)n <= P.maxLeafSize && cost >= leafCost)) {
            makeLeaf(L, me, refs, depth);
            return me;
        }
        std::vector<uint32_t> left, right;
        left.reserve(n);
        right.reserve(n);
        for (uint32_t r : refs) {
            if (goesLeft(tb[r], axis, split)) left.push_back(r);
            if (goesRight(tb[r], axis, split)) right.push_back(r);
        }
        if (left.size() == n && right.size() == n) { makeLeaf(L, me, refs, depth); return me; }
        std::vector<uint32_t>().swap(refs);  // release the parent's list before recursing
        Box lb = box, rb = box;
        lb.mx[axis] = split;
        rb.mn[axis] = split;
        const uint32_t lc = build(L, left, lb, depth + 1, badRefines);
        const uint32_t rc = build(L, right, rb, depth + 1, badRefines);
        KdNode& nd = L.nodes[me];
        nd.kind = (uint32_t)axis;
        nd.split = split;
        nd.a = lc;
        nd.b = rc;
        return me;
    }
};

// append `sub` to `dst`, fixing child and leaf offsets; returns the new index of sub's root
uint32_t stitch(Builder::Local& dst, const Builder::Local& sub)
{
    const uint32_t nodeBase = (uint32_t)dst.nodes.size(), triBase = (uint32_t)dst.leafTris.size();
    for (KdNode nd : sub.nodes) {
        if (nd.kind < 3) { nd.a += nodeBase; nd.b += nodeBase; }
        else nd.a += triBase;
        dst.nodes.push_back(nd);
    }
    dst.leafTris.insert(dst.leafTris.end(), sub.leafTris.begin(), sub.leafTris.end());
    dst.leaves += sub.leaves;
    dst.maxDepth = std::max(dst.maxDepth, sub.maxDepth);
    return nodeBase;
}

// ---- binary tree -> 32-byte blocks (two levels per block), DFS pre-order
// device leaf lists: [count, index 0, index 1, ...] per leaf, in block (DFS) order
struct LeafLists {
    const std::vector<uint32_t>& raw;  // the build's lists (KdNode::a = first, b = count)
    std::vector<uint32_t> out;
};

uint32_t leafRef(LeafLists& ll, const KdNode& n)
{
    if (n.b == 0) return HXR_KD_EMPTY;
    const size_t off = ll.out.size();
    ll.out.push_back(n.b);
    ll.out.insert(ll.out.end(), ll.raw.begin() + n.a, ll.raw.begin() + n.a + n.b);
    return HXR_KD_LEAF | (uint32_t)off;
}

uint32_t makeBlock(KdTree& t, LeafLists& ll, uint32_t nodeIdx)
{
    const uint32_t me = (uint32_t)t.blocks.size();
    t.blocks.push_back(KdBlock{});
    const KdNode n = t.nodes[nodeIdx];
    KdBlock b{};
    b.split[0] = n.split;
    b.meta = n.kind;
    for (int c = 0; c < 2; c++) {
        const KdNode child = t.nodes[c ? n.b : n.a];
        if (child.kind == 3) {
            b.meta |= 3u << (2 + 2 * c);
            b.split[1 + c] = 0;
            b.ref[2 * c] = leafRef(ll, child);
            b.ref[2 * c + 1] = HXR_KD_EMPTY;
        } else {
            b.meta |= child.kind << (2 + 2 * c);
            b.split[1 + c] = child.split;
            for (int k = 0; k < 2; k++) {
                const uint32_t gi = k ? child.b : child.a;
                const KdNode g = t.nodes[gi];
                b.ref[2 * c + k] = g.kind == 3 ? leafRef(ll, g) : makeBlock(t, ll, gi);
            }
        }
    }
    t.blocks[me] = b;
    return me;
}

void makeBlocks(KdTree& t, const hxr_mesh& mesh)
{
    t.blocks.clear();
    LeafLists ll{t.leafTris, {}};
    ll.out.reserve(t.leafTris.size() + t.leaves + 4);
    if (t.nodes[0].kind == 3) {
        // the whole mesh is one leaf: a block whose plane lies beyond the mesh puts everything on its left
        KdBlock b{};
        b.split[0] = std::nextafter((float)mesh.bbox_max[0], INFINITY) + 1.0f + 1e-3f * std::fabs((float)mesh.bbox_max[0]);
        b.meta = 0u | (3u << 2) | (3u << 4);
        b.ref[0] = leafRef(ll, t.nodes[0]);
        b.ref[1] = b.ref[2] = b.ref[3] = HXR_KD_EMPTY;
        t.blocks.push_back(b);
    } else {
        t.blocks.reserve(t.nodes.size() / 3 + 16);
        makeBlock(t, ll, 0);
    }
    ll.out.resize(ll.out.size() + 32, 0);  // readers may fetch (and then ignore) a few entries past a list
    t.leafTris.swap(ll.out);
}

}  // namespace

void packKdTree(KdTree& tree, const hxr_mesh& mesh)
{
    if (tree.nodes.empty()) tree.nodes.push_back(KdNode{0, 3, 0, 0});
    makeBlocks(tree, mesh);
}

void buildKdTree(const hxr_mesh& mesh, const KdBuildParams& params, KdTree& out)
{
    const auto t0 = std::chrono::steady_clock::now();
    Builder B(mesh, params);
    const int n = mesh.n_triangles;
    Box root;
    kdRootBox(mesh, root.mn, root.mx);
    std::vector<uint32_t> all(n);
    for (int i = 0; i < n; i++) all[i] = (uint32_t)i;

    Builder::Local top;
    int threads = params.threads > 0 ? params.threads : (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (n < 200000 || threads == 1) {
        B.build(top, all, root, 0, 0);
    } else {
        // split the top of the tree sequentially until the pending subtrees are small, then build
        // those subtrees concurrently and stitch them in.
        struct Pending { std::vector<uint32_t> refs; Box box; int depth; uint32_t parent; int side; };
        std::vector<Pending> pending;
        const size_t grain = std::max<size_t>(50000, (size_t)n / (size_t)(threads * 8));
        struct Work { std::vector<uint32_t> refs; Box box; int depth; uint32_t parent; int side; };
        std::vector<Work> stack;
        stack.push_back(Work{std::move(all), root, 0, UINT32_MAX, 0});
        uint32_t rootIdx = UINT32_MAX;
        while (!stack.empty()) {
            Work w = std::move(stack.back());
            stack.pop_back();
            int axis;
            float split, cost;
            if (w.refs.size() <= grain || w.depth >= 12 || !B.chooseSplit(w.refs, w.box, axis, split, cost)) {
                pending.push_back(Pending{std::move(w.refs), w.box, w.depth, w.parent, w.side});
                continue;
            }
            const uint32_t me = (uint32_t)top.nodes.size();
            top.nodes.push_back(KdNode{split, (uint32_t)axis, 0, 0});
            if (w.parent == UINT32_MAX) rootIdx = me;
            else (w.side ? top.nodes[w.parent].b : top.nodes[w.parent].a) = me;
            Work l, r;
            l.box = r.box = w.box;
            l.box.mx[axis] = split;
            r.box.mn[axis] = split;
            l.depth = r.depth = w.depth + 1;
            l.parent = r.parent = me;
            l.side = 0;
            r.side = 1;
            for (uint32_t t : w.refs) {
                if (Builder::goesLeft(B.tb[t], axis, split)) l.refs.push_back(t);
                if (Builder::goesRight(B.tb[t], axis, split)) r.refs.push_back(t);
            }
            std::vector<uint32_t>().swap(w.refs);
            stack.push_back(std::move(r));
            stack.push_back(std::move(l));
        }
        std::vector<Builder::Local> subs(pending.size());
        std::atomic<size_t> next{0};
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&] {
                for (size_t i = next++; i < pending.size(); i = next++) B.build(subs[i], pending[i].refs, pending[i].box, pending[i].depth, 0);
            });
        for (auto& th : pool) th.join();
        for (size_t i = 0; i < pending.size(); i++) {
            const uint32_t r = stitch(top, subs[i]);
            if (pending[i].parent == UINT32_MAX) rootIdx = r;
            else (pending[i].side ? top.nodes[pending[i].parent].b : top.nodes[pending[i].parent].a) = r;
            Builder::Local().nodes.swap(subs[i].nodes);
            std::vector<uint32_t>().swap(subs[i].leafTris);
        }
        if (rootIdx != 0) {
            // the traversal starts at node 0: swap the root into place if the first emitted node is not it
            // (cannot happen: the first node pushed is always the root or the only pending subtree's root)
        }
    }
    out.nodes.swap(top.nodes);
    out.leafTris.swap(top.leafTris);
    out.maxDepth = top.maxDepth;
    out.leaves = top.leaves;
    if (out.nodes.empty()) out.nodes.push_back(KdNode{0, 3, 0, 0});
    makeBlocks(out, mesh);
    out.buildMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace host
}  // namespace hxr
