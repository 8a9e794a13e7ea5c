// kdbuild.h — the per-item bodies of the DEVICE KD-tree build ("GPU refit/build").
//
// The reference rebuilds its tree on the host at every start (Mesh::buildKD, src/mesh.cpp:95-122: median split, O(N depth)
// with vector copies). csrc/host/kdtree.cpp replaced that with a surface-area-heuristic build on the host cores; this file
// holds the same SAH build re-expressed level by level so that one level of the WHOLE tree is a handful of data-parallel
// passes over the level's triangle references (kdbuild_cuda.cu runs them as kernels, tests/emu as loops):
//   bin       per reference of a big node: start / end bin of its clipped bounds on each axis   (histograms)
//   choose    per node: the cheapest SAH plane - 31 binned planes per axis for big nodes, every triangle bound edge for
//             small ones (exact sweep) - and the leaf-or-split rule
//   classify  per reference: goes left / goes right of its node's plane
//   plan      per node: child sizes from the scanned flags; sizes of the next level
//   emit      per node: its record in the output tree, its two children's work records
//   scatter   per reference: into its children's segments (stable: a leaf's triangles stay in ascending order) or into the
//             leaf lists
// The rules are the host builder's (kdtree.cpp): a triangle goes LEFT if its bounds reach below the split or it lies in the
// split plane, RIGHT if they reach above it; the split is a float and the same float-rounded value classifies here and
// steers the walk on the device; costs are evaluated in double and compared as floats; ties go to the lower axis, then
// to the lower plane.
#pragma once
#include <cmath>
#include <cstdint>
#include "hd.h"

namespace hxr {
namespace kdb {

#define HXR_KDB_BINS 32
#define HXR_KDB_EXACT_MAX 192 /* nodes with at most this many references sweep every bound edge */

struct Params {
    float traversalCost, intersectCost, emptyBonus;
    int32_t maxLeafSize, maxDepth, binnedAbove;  // binnedAbove <= HXR_KDB_EXACT_MAX
};

// one node of the level being split
struct NodeWork {       // 64 B
    double mn[3], mx[3];  // its box
    uint32_t start, count;  // its references in the level's reference array
    uint32_t out;         // its index in the output node array
    uint32_t bad;         // splits on the way down that cost more than a leaf (three of them end the branch)
};

struct Decision {  // 16 B
    float split;
    int32_t axis;  // 0..2, or 3: leaf
    uint32_t bad;
    uint32_t nl;   // references that go left (filled by plan)
};

// what the tree is made of on the way out (KdNode of scene_dev.h: split, kind, a, b)
struct OutNode {
    float split;
    uint32_t kind, a, b;
};

HXR_HD bool goes_left(double mn, double mx, float split)
{
    const double s = (double)split;
    return mn < s || (mn == s && mx == s);
}
HXR_HD bool goes_right(double mx, float split) { return mx > (double)split; }

HXR_HD double box_area(const NodeWork& w)
{
    const double dx = w.mx[0] - w.mn[0], dy = w.mx[1] - w.mn[1], dz = w.mx[2] - w.mn[2];
    return 2.0 * (dx * dy + dy * dz + dz * dx);
}

// SAH cost of splitting node w at `split` on `axis` with nl / nr references on the two sides
HXR_HD float sah_cost(const Params& P, const NodeWork& w, int axis, double split, uint32_t nl, uint32_t nr)
{
    const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    const double e1 = w.mx[a1] - w.mn[a1], e2 = w.mx[a2] - w.mn[a2];
    const double area = box_area(w);
    const double invArea = 1.0 / (area > 1e-300 ? area : 1e-300);
    const double dl = split - w.mn[axis], dr = w.mx[axis] - split;
    const double aL = 2.0 * (e1 * e2 + dl * (e1 + e2));
    const double aR = 2.0 * (e1 * e2 + dr * (e1 + e2));
    const double eb = (nl == 0 || nr == 0) ? (double)P.emptyBonus : 0.0;
    return (float)((double)P.traversalCost + (double)P.intersectCost * (1.0 - eb) * (aL * invArea * (double)nl + aR * invArea * (double)nr));
}

// start / end bin of a reference's bounds clipped to the node, on one axis (ext = box extent > 0)
HXR_HD void bin_range(double mn, double mx, double bmn, double bmx, int& b0, int& b1)
{
    const double scale = HXR_KDB_BINS / (bmx - bmn);
    const double lo = mn > bmn ? mn : bmn, hi = mx < bmx ? mx : bmx;
    int i0 = (int)((lo - bmn) * scale), i1 = (int)((hi - bmn) * scale);
    i0 = i0 < 0 ? 0 : i0;
    i1 = i1 < 0 ? 0 : i1;
    b0 = i0 > HXR_KDB_BINS - 1 ? HXR_KDB_BINS - 1 : i0;
    b1 = i1 > HXR_KDB_BINS - 1 ? HXR_KDB_BINS - 1 : i1;
}

// plane k (1..31) of the binned sweep
HXR_HD float binned_plane(const NodeWork& w, int axis, int k) { return (float)(w.mn[axis] + (w.mx[axis] - w.mn[axis]) * k / HXR_KDB_BINS); }
HXR_HD bool plane_inside(const NodeWork& w, int axis, float split) { return (double)split > w.mn[axis] && (double)split < w.mx[axis]; }

// "is (c, s) better than (bc, bs)": lower cost, then the lower plane
HXR_HD bool better(float c, float s, float bc, float bs) { return c < bc || (c == bc && s < bs); }

// the leaf-or-split rule once the cheapest plane is known (kdtree.cpp: Builder::build). Returns true to split.
HXR_HD bool keep_split(const Params& P, uint32_t n, float cost, uint32_t& bad)
{
    const float leafCost = P.intersectCost * (float)n;
    if (cost > leafCost) bad++;
    if ((cost > 4 * leafCost && n < 16) || bad >= 3 || ((int)n <= P.maxLeafSize && cost >= leafCost)) return false;
    return true;
}

// ---- per-reference and per-node items shared by the kernels and the emulation loops

// classify: flags of reference i (tb: triangle bounds, SoA [6][nTris]: min x y z, max x y z)
HXR_HD void classify_item(uint32_t i, const uint32_t* refTri, const uint32_t* refNode, const Decision* dec, const double* tb, size_t nTris, uint32_t* flagL,
                          uint32_t* flagR)
{
    const Decision d = dec[refNode[i]];
    uint32_t l = 0, r = 0;
    if (d.axis < 3) {
        const uint32_t t = refTri[i];
        const double mn = tb[(size_t)d.axis * nTris + t], mx = tb[(size_t)(3 + d.axis) * nTris + t];
        l = goes_left(mn, mx, d.split) ? 1u : 0u;
        r = goes_right(mx, d.split) ? 1u : 0u;
    }
    flagL[i] = l;
    flagR[i] = r;
}

// plan: node sizes of the next level. scanL / scanR: exclusive scans of the flags (one element past the end).
HXR_HD void plan_item(uint32_t node, const NodeWork* work, Decision* dec, const uint32_t* scanL, const uint32_t* scanR, uint32_t* childRefs, uint32_t* isSplit,
                      uint32_t* leafRefs)
{
    const NodeWork w = work[node];
    Decision d = dec[node];
    const uint32_t nl = scanL[w.start + w.count] - scanL[w.start], nr = scanR[w.start + w.count] - scanR[w.start];
    bool split = d.axis < 3;
    if (split && nl == w.count && nr == w.count) split = false;  // the plane separates nothing
    if (!split) d.axis = 3;
    d.nl = nl;
    dec[node] = d;
    childRefs[node] = split ? nl + nr : 0u;
    isSplit[node] = split ? 1u : 0u;
    leafRefs[node] = split ? 0u : w.count;
}

// emit: the node's record in the output tree and its children's work records. childRefs / isSplit / leafRefs are now their
// exclusive scans; outCount = nodes emitted before this level's children, leafBase = leaf references before this level.
HXR_HD void emit_item(uint32_t node, const NodeWork* work, const Decision* dec, const uint32_t* childRefs, const uint32_t* isSplit, const uint32_t* leafRefs,
                      uint32_t outCount, uint32_t leafBase, OutNode* out, NodeWork* next)
{
    const NodeWork w = work[node];
    const Decision d = dec[node];
    OutNode o;
    if (d.axis < 3) {
        const uint32_t rank = isSplit[node];
        const uint32_t c0 = outCount + 2 * rank;
        o.split = d.split; o.kind = (uint32_t)d.axis; o.a = c0; o.b = c0 + 1;
        const uint32_t nr = childRefs[node + 1] - childRefs[node] - d.nl;
        NodeWork l = w, r = w;
        l.mx[d.axis] = (double)d.split;
        r.mn[d.axis] = (double)d.split;
        l.start = childRefs[node]; l.count = d.nl; l.out = c0; l.bad = d.bad;
        r.start = childRefs[node] + d.nl; r.count = nr; r.out = c0 + 1; r.bad = d.bad;
        next[2 * rank] = l;
        next[2 * rank + 1] = r;
    } else {
        o.split = 0; o.kind = 3; o.a = leafBase + leafRefs[node]; o.b = w.count;
    }
    out[w.out] = o;
}

// scatter: reference i into its children's segments, or into the leaf lists
HXR_HD void scatter_item(uint32_t i, const uint32_t* refTri, const uint32_t* refNode, const NodeWork* work, const Decision* dec, const uint32_t* scanL,
                         const uint32_t* scanR, const uint32_t* childRefs, const uint32_t* isSplit, const uint32_t* leafRefs, uint32_t leafBase,
                         uint32_t* nextTri, uint32_t* nextNode, uint32_t* leafOut)
{
    const uint32_t node = refNode[i];
    const NodeWork w = work[node];
    const Decision d = dec[node];
    const uint32_t t = refTri[i];
    if (d.axis < 3) {
        const uint32_t rank = isSplit[node], base = childRefs[node];
        if (scanL[i + 1] != scanL[i]) {
            const uint32_t dst = base + (scanL[i] - scanL[w.start]);
            nextTri[dst] = t;
            nextNode[dst] = 2 * rank;
        }
        if (scanR[i + 1] != scanR[i]) {
            const uint32_t dst = base + d.nl + (scanR[i] - scanR[w.start]);
            nextTri[dst] = t;
            nextNode[dst] = 2 * rank + 1;
        }
    } else {
        leafOut[leafBase + leafRefs[node] + (i - w.start)] = t;
    }
}

// bounds of triangle t (vertex extents, exact)
HXR_HD void bounds_item(uint32_t t, const double* vertices, const int32_t* triV, double* tb, size_t nTris)
{
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int k = 0; k < 3; k++) {
        const double* v = vertices + 3 * (size_t)triV[3 * (size_t)t + k];
        for (int a = 0; a < 3; a++) {
            mn[a] = v[a] < mn[a] ? v[a] : mn[a];
            mx[a] = v[a] > mx[a] ? v[a] : mx[a];
        }
    }
    for (int a = 0; a < 3; a++) {
        tb[(size_t)a * nTris + t] = mn[a];
        tb[(size_t)(3 + a) * nTris + t] = mx[a];
    }
}

// choose, serial form (one node): what the warp-per-node kernel computes cooperatively. hist = this node's
// [3][2][HXR_KDB_BINS] start / end counts (big nodes only).
inline void choose_serial(const Params& P, const NodeWork& w, int depth, const uint32_t* hist, const uint32_t* refTri, const double* tb, size_t nTris, Decision& d)
{
    const uint32_t n = w.count;
    d.split = 0; d.axis = 3; d.bad = w.bad; d.nl = 0;
    if (n <= 1 || depth >= P.maxDepth) return;
    float bestCost = INFINITY, bestSplit = 0;
    int bestAxis = -1;
    for (int axis = 0; axis < 3; axis++) {
        if (!(w.mx[axis] - w.mn[axis] > 0)) continue;
        float ac = INFINITY, as = 0;
        if ((int)n > P.binnedAbove) {
            const uint32_t* sc = hist + (axis * 2 + 0) * HXR_KDB_BINS;
            const uint32_t* ecn = hist + (axis * 2 + 1) * HXR_KDB_BINS;
            uint32_t nl = 0, nr = n;
            for (int k = 1; k < HXR_KDB_BINS; k++) {
                nl += sc[k - 1];
                nr -= ecn[k - 1];
                const float s = binned_plane(w, axis, k);
                if (!plane_inside(w, axis, s)) continue;
                const float c = sah_cost(P, w, axis, s, nl, nr);
                if (better(c, s, ac, as)) { ac = c; as = s; }
            }
        } else {
            for (uint32_t ci = 0; ci < 2 * n; ci++) {
                const uint32_t t = refTri[w.start + (ci < n ? ci : ci - n)];
                const float s = (float)tb[(size_t)(ci < n ? axis : 3 + axis) * nTris + t];
                if (!plane_inside(w, axis, s)) continue;
                const double sd = (double)s;
                uint32_t nl = 0, nr = 0;
                for (uint32_t j = 0; j < n; j++) {
                    const uint32_t tj = refTri[w.start + j];
                    const double mn = tb[(size_t)axis * nTris + tj], mx = tb[(size_t)(3 + axis) * nTris + tj];
                    nl += (mn < sd || (mn == sd && mx == sd)) ? 1u : 0u;
                    nr += mx > sd ? 1u : 0u;
                }
                const float c = sah_cost(P, w, axis, sd, nl, nr);
                if (better(c, s, ac, as)) { ac = c; as = s; }
            }
        }
        if (ac < bestCost) { bestCost = ac; bestAxis = axis; bestSplit = as; }
    }
    if (bestAxis < 0) return;
    uint32_t bad = w.bad;
    if (!keep_split(P, n, bestCost, bad)) return;
    d.split = bestSplit; d.axis = bestAxis; d.bad = bad;
}

}  // namespace kdb
}  // namespace hxr
