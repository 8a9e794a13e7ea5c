// isect.h — ray/geometry intersectors (device functions).
//
// Each function restates one reference intersector with the reference's own arithmetic
// (double geometry, same epsilons, same acceptance tests) so that hit records agree with
// the oracle to rounding:
//   plane_intersect        src/geometry.cpp:31-50
//   sphere_intersect       src/geometry.cpp:52-85
//   cube_intersect         src/geometry.cpp:87-147
//   csg_intersect          src/geometry.cpp:149-194 (findAllIntersections + CSGBase::intersect)
//   mesh_intersect         src/mesh.cpp:178-263 (triangle test) — traversal is OUR KD-tree, see below
//   heightfield_intersect  src/heightfield.cpp:106-171 (+ bbox.h:145-174 closestIntersection)
//   node_intersect         src/geometry.cpp:196-208 + src/matrix.cpp:156-187
//   rect_light_intersect   src/lights.cpp:53-73
//
// Mesh traversal: the reference walks a median-split tree recursively and keeps a leaf hit only
// if it lies inside the leaf box (mesh.cpp:221-245); the net result is the closest triangle hit
// under the triangle test's own arithmetic. We get the same result from a SAH KD-tree (two levels
// per 32-byte block) walked front to back with an explicit stack: every triangle whose bounds
// overlap a leaf is referenced by it, hits are accepted wherever they fall, and a pending subtree
// is skipped only when its whole parameter range lies beyond the best hit. Two forms walk it:
// mesh_closest (double planes, one ray: CSG children, CPU tests) and the conservative FP32 walk
// (plane_cross / block_step / tri_filter below: what the GPU kernel runs).
#pragma once
#include <cmath>
#include <cstring>
#include "scene_dev.h"

namespace hxr {

HXR_HD bool bbox_inside(const double* mn, const double* mx, const d3& v)  // src/bbox.h:80-85
{
    return (mn[0] - 1e-6 <= v.x && v.x <= mx[0] + 1e-6 &&
            mn[1] - 1e-6 <= v.y && v.y <= mx[1] + 1e-6 &&
            mn[2] - 1e-6 <= v.z && v.z <= mx[2] + 1e-6);
}

// src/bbox.h:145-174
HXR_HD double bbox_closest_intersection(const double* mn, const double* mx, const Ray& r)
{
    if (bbox_inside(mn, mx, r.o)) return 0;
    double minDist = HXR_INF;
    for (int dim = 0; dim < 3; dim++) {
        double dd = comp(r.d, dim), so = comp(r.o, dim);
        if ((dd < 0 && so < mn[dim]) || (dd > 0 && so > mx[dim])) return HXR_INF;
        if (fabs(dd) < 1e-9) continue;
        double mul = 1 / dd;
        int u = (dim == 0) ? 1 : 0;
        int v = (dim == 2) ? 1 : 2;
        for (int j = 0; j < 2; j++) {
            double dist = ((j ? mx[dim] : mn[dim]) - so) * mul;
            if (dist < 0) continue;
            double x = comp(r.o, u) + comp(r.d, u) * dist;
            if (mn[u] <= x && x <= mx[u]) {
                double y = comp(r.o, v) + comp(r.d, v) * dist;
                if (mn[v] <= y && y <= mx[v]) minDist = dist < minDist ? dist : minDist;
            }
        }
    }
    return minDist;
}

HXR_HD bool plane_intersect(const hxr_geometry& g, int gi, const Ray& ray, Hit& info)
{
    const double y = g.p[0], limit = g.p[1];
    if (ray.o.y > y && ray.d.y >= 0) return false;
    if (ray.o.y < y && ray.d.y <= 0) return false;
    double going = ray.d.y;
    double toGo = y - ray.o.y;
    double m = toGo / going;
    info.dist = m;
    info.ip = ray.o + ray.d * m;
    if (fabs(info.ip.x) > limit || fabs(info.ip.z) > limit) return false;
    info.norm = mk3(0, (ray.o.y > y) ? 1 : -1, 0);
    info.u = info.ip.x;
    info.v = info.ip.z;
    info.dNdx = mk3(1, 0, 0);
    info.dNdy = mk3(0, 0, 1);
    info.geom = gi;
    return true;
}

HXR_HD bool sphere_intersect(const hxr_geometry& g, int gi, const Ray& ray, Hit& info)
{
    const d3 O = ld3(g.p);
    const double R = g.p[3], uvscaling = g.p[4];
    double A = length_sqr(ray.d);
    d3 H = ray.o - O;
    double B = 2 * dot(ray.d, H);
    double C = length_sqr(H) - R * R;
    double D = B * B - 4 * A * C;
    if (D < 0) return false;
    double sqrtD = sqrt(D);
    double p1 = (-B - sqrtD) / (2 * A);
    double p2 = (-B + sqrtD) / (2 * A);
    double p;
    if (p2 < 0) return false;
    if (p1 < 0) p = p2; else p = p1;
    info.dist = p;
    info.ip = ray.o + ray.d * p;
    info.norm = normalize_m(info.ip - O);
    info.v = asin(info.norm.y);
    info.u = atan2(info.norm.z, info.norm.x);
    info.v = -(info.v / HXR_PI + 0.5f);
    info.u = info.u / (2 * HXR_PI) + 0.5f;
    if (uvscaling != 1) {
        info.u *= uvscaling;
        info.v *= uvscaling;
    }
    // the reference leaves dNdx/dNdy unwritten for spheres; keep them defined here
    info.dNdx = mk3(0, 0, 0);
    info.dNdy = mk3(0, 0, 0);
    info.geom = gi;
    return true;
}

HXR_HD bool cube_in_bounds(double x, double center, double halfSide)
{
    return (x > center - halfSide - 1e-6 && x < center + halfSide + 1e-6);
}

// one of the six sides; `axis` selects which uv pair the side writes (geometry.cpp:130-132)
HXR_HD int cube_side(const d3& O, double hs, const d3& norm, double startCoord, double dir, double target,
                     const Ray& ray, Hit& info, int axis, int gi)
{
    if (fabs(dir) < 1e-9) return 0;
    if (startCoord < target && dir < 0) return 0;
    if (startCoord > target && dir > 0) return 0;
    double p = (target - startCoord) / dir;
    if (p < info.dist) {
        d3 ip = ray.o + ray.d * p;
        if (!cube_in_bounds(ip.x, O.x, hs) || !cube_in_bounds(ip.y, O.y, hs) || !cube_in_bounds(ip.z, O.z, hs)) return 0;
        info.dist = p;
        info.ip = ip;
        info.norm = norm;
        if (axis == 0) { info.u = ip.y; info.v = ip.z; }
        else if (axis == 1) { info.u = ip.x; info.v = ip.z; }
        else { info.u = ip.x; info.v = ip.y; }
        info.geom = gi;
        return 1;
    }
    return 0;
}

HXR_HD bool cube_intersect(const hxr_geometry& g, int gi, const Ray& ray, Hit& info)
{
    const d3 O = ld3(g.p);
    const double hs = g.p[3];
    int n = 0;
    info.dist = HXR_INF;
    info.dNdx = mk3(0, 0, 0);  // unwritten in the reference
    info.dNdy = mk3(0, 0, 0);
    n += cube_side(O, hs, mk3(-1, 0, 0), ray.o.x, ray.d.x, O.x - hs, ray, info, 0, gi);
    n += cube_side(O, hs, mk3(+1, 0, 0), ray.o.x, ray.d.x, O.x + hs, ray, info, 0, gi);
    n += cube_side(O, hs, mk3(0, -1, 0), ray.o.y, ray.d.y, O.y - hs, ray, info, 1, gi);
    n += cube_side(O, hs, mk3(0, +1, 0), ray.o.y, ray.d.y, O.y + hs, ray, info, 1, gi);
    n += cube_side(O, hs, mk3(0, 0, -1), ray.o.z, ray.d.z, O.z - hs, ray, info, 2, gi);
    n += cube_side(O, hs, mk3(0, 0, +1), ray.o.z, ray.d.z, O.z + hs, ray, info, 2, gi);
    return n > 0;
}

// src/mesh.cpp:131-170 (used by the heightfield cells)
HXR_HD bool intersect_triangle_fast(const Ray& ray, const d3& A, const d3& B, const d3& C, double& dist)
{
    d3 AB = B - A;
    d3 AC = C - A;
    d3 D = -ray.d;
    d3 H = ray.o - A;
    d3 ABcrossAC = cross(AB, AC);
    double Dcr = dot(ABcrossAC, D);
    if (fabs(Dcr) < 1e-12) return false;
    double lambda2 = dot(cross(H, AC), D) / Dcr;
    double lambda3 = dot(cross(AB, H), D) / Dcr;
    double gamma = dot(ABcrossAC, H) / Dcr;
    if (gamma < 0 || gamma > dist) return false;
    if (lambda2 < 0 || lambda2 > 1 || lambda3 < 0 || lambda3 > 1 || lambda2 + lambda3 > 1) return false;
    dist = gamma;
    return true;
}

// ---------------------------------------------------------------- triangle mesh
// The mesh query is split into pieces so that the same arithmetic serves (i) the simple per-ray
// function used for CSG children and the host emulation, and (ii) the persistent traversal kernel
// (launch_cuda.cu), which walks the same tree with conservative FP32 plane arithmetic:
//   mesh_slab          ray parameters [t0, t1] against the (slightly inflated) mesh box
//   tri_test           the reference's triangle test (src/mesh.cpp:178-196), verbatim arithmetic
//   mesh_closest       front-to-back KD walk, one binary node at a time, double plane arithmetic
//   mesh_fill_hit      IntersectionInfo of the winning triangle (src/mesh.cpp:197-218)
//
// Which hit wins: the smallest gamma; among exactly equal gammas the HIGHEST triangle index. That is what the
// reference's loops produce (ascending index order, `gamma > info.dist` rejects, so an equal hit overwrites:
// src/mesh.cpp:189, 255-262) and it makes the result independent of the order in which leaves are visited.

struct MeshBest {  // winner so far inside one mesh, in object space
    double gamma, l2, l3;
    int tri;
};

HXR_HD bool mesh_slab(const DMesh& M, const Ray& ray, double gamma_limit, double& t0, double& t1)
{
    t0 = 0;
    t1 = gamma_limit;
    const double o[3] = {ray.o.x, ray.o.y, ray.o.z};
    const double d[3] = {ray.d.x, ray.d.y, ray.d.z};
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int a = 0; a < 3; a++) {
        const double lo = M.bbmin[a] - 1e-6, hi = M.bbmax[a] + 1e-6;
        if (d[a] == 0) {
            if (o[a] < lo || o[a] > hi) return false;
        } else {
            const double inv = 1.0 / d[a];
            double ta = (lo - o[a]) * inv, tb = (hi - o[a]) * inv;
            if (ta > tb) { const double s = ta; ta = tb; tb = s; }
            t0 = ta > t0 ? ta : t0;
            t1 = tb < t1 ? tb : t1;
        }
    }
    return t0 <= t1;
}

// loads one 96-byte triangle record with 128-bit loads
struct TriRec {
    d3 A, AB, AC, N;
};
HXR_HD TriRec load_tri(const TriTest* p)
{
    TriRec r;
#if defined(__CUDA_ARCH__)
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4), f = __ldg(q + 5);
    r.A = mk3(a.x, a.y, b.x);
    r.AB = mk3(b.y, c.x, c.y);
    r.AC = mk3(d.x, d.y, e.x);
    r.N = mk3(e.y, f.x, f.y);
#else
    r.A = ld3(p->A); r.AB = ld3(p->AB); r.AC = ld3(p->AC); r.N = ld3(p->N);
#endif
    return r;
}

// The reference's triangle test (src/mesh.cpp:178-196) against the best hit so far (bestGamma, bestTri): true if
// triangle `ti` is hit at a parameter in [0, bestGamma] (at exactly bestGamma only if ti > bestTri).
HXR_HD bool tri_core(const TriTest* tris, bool backface, const Ray& ray, uint32_t ti, double bestGamma, int bestTri, double& gamma,
                     double& lambda2, double& lambda3)
{
    const TriRec t = load_tri(tris + ti);
    if (backface && dot(ray.d, t.N) > 0) return false;
    const d3 nd = -ray.d;
    const d3 H = ray.o - t.A;
    const double Dcr = dot(t.N, nd);
    if (fabs(Dcr) < 1e-12) return false;
    const double rDcr = 1 / Dcr;
    gamma = dot(t.N, H) * rDcr;
    if (gamma < 0 || gamma > bestGamma) return false;
    if (gamma == bestGamma && (int)ti < bestTri) return false;
    lambda2 = dot(cross(H, t.AC), nd) * rDcr;
    if (lambda2 < 0 || lambda2 > 1) return false;
    lambda3 = dot(cross(t.AB, H), nd) * rDcr;
    if (lambda3 < 0 || lambda3 > 1) return false;
    const double lambda1 = 1 - (lambda2 + lambda3);
    if (lambda1 < 0 || lambda1 > 1) return false;
    return true;
}

// returns true (and updates best) if triangle `ti` beats the best hit so far
HXR_HD bool tri_test(const TriTest* tris, bool backface, const Ray& ray, uint32_t ti, MeshBest& best)
{
    double gamma, l2, l3;
    if (!tri_core(tris, backface, ray, ti, best.gamma, best.tri, gamma, l2, l3)) return false;
    best.gamma = gamma;
    best.tri = (int)ti;
    best.l2 = l2;
    best.l3 = l3;
    return true;
}

// gamma and barycentrics of a triangle already known to be hit (the expressions of tri_test, no range checks)
HXR_HD void tri_eval(const TriTest* tris, const Ray& ray, uint32_t ti, MeshBest& best)
{
    const TriRec t = load_tri(tris + ti);
    const d3 nd = -ray.d;
    const d3 H = ray.o - t.A;
    const double rDcr = 1 / dot(t.N, nd);
    best.gamma = dot(t.N, H) * rDcr;
    best.l2 = dot(cross(H, t.AC), nd) * rDcr;
    best.l3 = dot(cross(t.AB, H), nd) * rDcr;
    best.tri = (int)ti;
}

HXR_HD KdBlock load_block(const KdBlock* p)
{
#if defined(__CUDA_ARCH__)
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p)), v = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    KdBlock b;
    b.split[0] = __uint_as_float(u.x); b.split[1] = __uint_as_float(u.y); b.split[2] = __uint_as_float(u.z);
    b.meta = u.w;
    b.ref[0] = v.x; b.ref[1] = v.y; b.ref[2] = v.z; b.ref[3] = v.w;
    return b;
#else
    return *p;
#endif
}

// Front-to-back walk, one binary node per step. Cursor: HXR_KD_LEAF | first entry, or (block << 2) | sub.
template <bool COUNT>
HXR_HD bool mesh_closest(const DMesh& M, const Ray& ray, double gamma_limit, MeshBest& best, TravCounters* cnt)
{
    double tmin, tmax;
    if (!mesh_slab(M, ray, gamma_limit, tmin, tmax)) return false;
    if (COUNT) cnt->mesh_queries++;
    best.gamma = gamma_limit;
    best.tri = -1;
    best.l2 = best.l3 = 0;
    uint32_t stRef[HXR_KD_STACK];
    double stMin[HXR_KD_STACK], stMax[HXR_KD_STACK];
    int sp = 0;
    uint32_t cur = 0;
    for (;;) {
        if (cur & HXR_KD_LEAF) {
            if (cur != HXR_KD_EMPTY) {
                if (COUNT) cnt->kd_leaves++;
                const uint32_t* list = M.leaf_tris + (cur & ~HXR_KD_LEAF);
                if (COUNT) cnt->tri_tests += list[0];
                for (uint32_t k = 1; k <= list[0]; k++) tri_test(M.tri_test, M.backface != 0, ray, list[k], best);
            }
            // next pending segment that can still hold a hit at or before the best one
            bool found = false;
            while (sp > 0) {
                sp--;
                if (stMin[sp] - 1e-7 * (1.0 + fabs(stMin[sp])) > best.gamma) continue;
                cur = stRef[sp]; tmin = stMin[sp]; tmax = stMax[sp];
                found = true;
                break;
            }
            if (!found) break;
            continue;
        }
        const uint32_t blk = cur >> 2, sub = cur & 3u;
        const KdBlock B = load_block(M.blocks + blk);
        if (COUNT && sub == 0) cnt->kd_inner++;
        const int axis = (int)((B.meta >> (2 * sub)) & 3u);
        const double split = (double)B.split[sub];
        uint32_t cl, cr;
        if (sub == 0) {
            cl = ((B.meta >> 2) & 3u) == 3u ? B.ref[0] : ((blk << 2) | 1u);
            cr = ((B.meta >> 4) & 3u) == 3u ? B.ref[2] : ((blk << 2) | 2u);
        } else {
            cl = B.ref[2 * (sub - 1)];
            cr = B.ref[2 * (sub - 1) + 1];
            if (!(cl & HXR_KD_LEAF)) cl <<= 2;
            if (!(cr & HXR_KD_LEAF)) cr <<= 2;
        }
        // "first" is the side the ray is on BEFORE it crosses the plane (left = coordinates <= split), whatever
        // side of the plane the origin lies on: an origin outside the node's box says nothing about the segment
        const double oa = comp(ray.o, axis), da = comp(ray.d, axis);
        if (da == 0) {
            if (oa == split && sp < HXR_KD_STACK) {  // travelling inside the split plane: both sides
                stRef[sp] = cr; stMin[sp] = tmin; stMax[sp] = tmax; sp++;
                cur = cl;
            } else {
                cur = oa < split ? cl : cr;
            }
            continue;
        }
        const uint32_t firstC = da > 0 ? cl : cr, secondC = da > 0 ? cr : cl;
        const double tpl = (split - oa) / da;
        const double slack = 1e-9 * (1.0 + fabs(tpl));
        if (tpl > tmax + slack) {  // the plane is crossed after this segment
            cur = firstC;
        } else if (tpl < tmin - slack) {  // ... or before it (also: behind the origin)
            cur = secondC;
        } else {
            if (sp < HXR_KD_STACK) { stRef[sp] = secondC; stMin[sp] = tpl; stMax[sp] = tmax; sp++; }
            cur = firstC;
            tmax = tpl;
        }
    }
    return best.tri >= 0;
}

// ---- the conservative FP32 walk (what the traversal kernel runs; the host form below is used by the CPU tests)
// The tree is walked with FP32 plane arithmetic made CONSERVATIVE: every plane parameter carries an error bound
// and the two children get overlapping parameter ranges, so a leaf is visited whenever the exact ray could touch
// it, while every triangle is still tested with the reference's double arithmetic (tri_test). The walk only
// decides WHICH triangles are tested; the winner is the same as for mesh_closest and for brute force.
HXR_HD float f32_below(double x)  // largest float <= x (round toward -inf)
{
#if defined(__CUDA_ARCH__)
    return __double2float_rd(x);
#else
    float f = (float)x;
    return (double)f > x ? nextafterf(f, -INFINITY) : f;
#endif
}
HXR_HD float f32_above(double x)  // smallest float >= x (round toward +inf)
{
#if defined(__CUDA_ARCH__)
    return __double2float_ru(x);
#else
    float f = (float)x;
    return (double)f < x ? nextafterf(f, INFINITY) : f;
#endif
}

struct WalkRay {  // the object-space ray as the walk sees it
    float ox, oy, oz;  // origin, rounded to nearest
    float ix, iy, iz;  // 1 / direction (0 on parallel axes)
    uint32_t par;      // bit a: |d[a]| < 1e-18, treated as parallel to the planes of axis a
    HXR_HD float o(uint32_t axis) const { return axis == 0 ? ox : (axis == 1 ? oy : oz); }
    HXR_HD float inv(uint32_t axis) const { return axis == 0 ? ix : (axis == 1 ? iy : iz); }
};
HXR_HD WalkRay walk_ray_f(float ox, float oy, float oz, float dx, float dy, float dz)
{
    WalkRay w;
    w.ox = ox; w.oy = oy; w.oz = oz;
    w.par = (fabsf(dx) < 1e-18f ? 1u : 0u) | (fabsf(dy) < 1e-18f ? 2u : 0u) | (fabsf(dz) < 1e-18f ? 4u : 0u);
    w.ix = (w.par & 1u) ? 0.0f : 1.0f / dx;
    w.iy = (w.par & 2u) ? 0.0f : 1.0f / dy;
    w.iz = (w.par & 4u) ? 0.0f : 1.0f / dz;
    return w;
}
HXR_HD WalkRay walk_ray(const Ray& t) { return walk_ray_f((float)t.o.x, (float)t.o.y, (float)t.o.z, (float)t.d.x, (float)t.d.y, (float)t.d.z); }

struct PlaneX {  // conservative range [tlo, thi] of the parameter at which the ray crosses a split plane
    float tlo, thi;
    bool leftFirst;  // the ray is on the left (coordinate <= split) before the crossing
};
// RAY provides o(axis), inv(axis) and par (WalkRay above; the kernel reads them from shared memory by axis).
// PAR = false skips the parallel-axis case (the caller has checked par == 0).
template <bool PAR, class RAY>
HXR_HD PlaneX plane_cross(float s, uint32_t axis, const RAY& w, float tmaxSeg)
{
    const float o = w.o(axis);
    const float inv = w.inv(axis);
    PlaneX r;
    // tpl = (s - o) / d in float: relative error <= 2^-22 (o and d rounded from double, one subtraction, one
    // reciprocal, one product) plus |o| 2^-24 |inv| from the rounding of o; both bounds doubled.
    const float tpl = (s - o) * inv;
    const float e = fmaf(fabsf(tpl), 4.76837158e-7f, fabsf(inv * o) * 1.1920929e-7f);
    r.tlo = tpl - e;
    r.thi = tpl + e;
    r.leftFirst = inv > 0;
    if (PAR && ((w.par >> axis) & 1u)) {
        // parallel to the plane for every parameter that matters: pick sides by position
        const float tol = fmaf(fabsf(o), 2.38418579e-7f, 1e-18f * tmaxSeg) + 1e-30f;
        r.leftFirst = true;
        r.thi = (o <= s + tol) ? INFINITY : -INFINITY;
        r.tlo = (o >= s - tol) ? -INFINITY : INFINITY;
    }
    return r;
}

struct WalkEnt {  // a subtree (block index or leaf reference) and the parameter range in which the ray can be inside it
    uint32_t ref;
    float lo, hi;
};
HXR_HD bool ent_valid(const WalkEnt& e) { return e.lo <= e.hi && e.ref != HXR_KD_EMPTY; }

// One block = a node and both its children: up to four grandchildren e0..e3, FRONT TO BACK (test ent_valid on each).
template <bool PAR, class RAY>
HXR_HD void block_step_t(const KdBlock& B, const RAY& w, float tmin, float tmax, float tbest, WalkEnt& e0, WalkEnt& e1, WalkEnt& e2, WalkEnt& e3)
{
    const uint32_t a0 = B.meta & 3u, aL = (B.meta >> 2) & 3u, aR = (B.meta >> 4) & 3u;
    const float tE = fminf(tmax, tbest);
    const PlaneX p0 = plane_cross<PAR>(B.split[0], a0, w, tmax);
    const float nHi = fminf(tE, p0.thi), fLo = fmaxf(tmin, p0.tlo);
    const float lLo = p0.leftFirst ? tmin : fLo, lHi = p0.leftFirst ? nHi : tE;
    const float rLo = p0.leftFirst ? fLo : tmin, rHi = p0.leftFirst ? tE : nHi;
    WalkEnt l0, l1, r0, r1;
    {
        const PlaneX pl = plane_cross<PAR>(B.split[1], aL, w, tmax);
        const bool leaf = aL == 3u;
        const bool lf = pl.leftFirst || leaf;
        l0.ref = lf ? B.ref[0] : B.ref[1];
        l1.ref = leaf ? HXR_KD_EMPTY : (lf ? B.ref[1] : B.ref[0]);
        l0.lo = lLo; l0.hi = leaf ? lHi : fminf(lHi, pl.thi);
        l1.lo = fmaxf(lLo, pl.tlo); l1.hi = lHi;
    }
    {
        const PlaneX pr = plane_cross<PAR>(B.split[2], aR, w, tmax);
        const bool leaf = aR == 3u;
        const bool lf = pr.leftFirst || leaf;
        r0.ref = lf ? B.ref[2] : B.ref[3];
        r1.ref = leaf ? HXR_KD_EMPTY : (lf ? B.ref[3] : B.ref[2]);
        r0.lo = rLo; r0.hi = leaf ? rHi : fminf(rHi, pr.thi);
        r1.lo = fmaxf(rLo, pr.tlo); r1.hi = rHi;
    }
    e0 = p0.leftFirst ? l0 : r0; e1 = p0.leftFirst ? l1 : r1; e2 = p0.leftFirst ? r0 : l0; e3 = p0.leftFirst ? r1 : l1;
}
template <class RAY>
HXR_HD void block_step(const KdBlock& B, const RAY& w, float tmin, float tmax, float tbest, WalkEnt& e0, WalkEnt& e1, WalkEnt& e2, WalkEnt& e3)
{
    if (w.par) block_step_t<true>(B, w, tmin, tmax, tbest, e0, e1, e2, e3);  // rare: a direction component is (almost) zero
    else block_step_t<false>(B, w, tmin, tmax, tbest, e0, e1, e2, e3);
}

// ---- the FP32 triangle filter
// The walk does not run the exact triangle test; it runs the SAME formulas in float on a float copy of the triangle
// (TriF32) together with a running bound on the rounding error of every quantity, and sorts each (ray, triangle)
// pair into: MISS (the exact test certainly rejects it, or its hit lies certainly beyond the best known one),
// CERTAIN (the exact test certainly accepts it, at a parameter <= ghi) or MAYBE. CERTAIN and MAYBE pairs go to the
// exact test (pipeline.h: confirm_*); CERTAIN ones also shorten the walk (tbest). Error model, eps = 2^-24:
//   every component of H = o - A carries |dH| <= err (rounding of o, of A and of the subtraction; task.err),
//   so |d((H x AC).nd)| <= 2 err |AC|_1 + 16 eps |H|_1 |AC|_1, likewise for (AB x H).nd and N.H, and
//   |d(N.nd)| <= 8 eps |N|_1 (|nd| = 1; the constants are the operation counts, doubled).
#define HXR_TF_MISS 0
#define HXR_TF_MAYBE 1
#define HXR_TF_CERTAIN 2

//   Packed triangles (TriPacked, PK = true): AB and AC are exact multiples of 2^e, off the true edges by at most
//   qab = 2^(eAB-1), qac = 2^(eAC-1) per component, and N is rebuilt as fl(AB' x AC'):
//     |d((H x AC).nd)| grows by 2 qac |H|_1, |d((AB x H).nd)| by 2 qab |H|_1,
//     sum_i |dN_i| <= 2 qab |AC|_1 + 2 qac |AB|_1 + 4 eps |AB|_1 |AC|_1 =: dN  (edge error + rounding of the cross product),
//     |d(N.nd)| grows by dN and |d(N.H)| by dN |H|_1.   (2.5 instead of 2 below: slack for the second-order terms.)
template <bool PK>
HXR_HD int tri_filter_core(float ax, float ay, float az, float abx, float aby, float abz, float acx, float acy, float acz, float nx, float ny,
                           float nz, float qab, float qac, bool backface, float ox, float oy, float oz, float dx, float dy, float dz, float err,
                           float tbest, float& ghi)
{
    const float eps = 5.9604645e-8f;
    const float hx = ox - ax, hy = oy - ay, hz = oz - az;
    const float ex = -dx, ey = -dy, ez = -dz;
    const float Dcr = nx * ex + ny * ey + nz * ez;
    const float Ng = nx * hx + ny * hy + nz * hz;
    const float c2 = (hy * acz - hz * acy) * ex + (hz * acx - hx * acz) * ey + (hx * acy - hy * acx) * ez;
    const float c3 = (aby * hz - abz * hy) * ex + (abz * hx - abx * hz) * ey + (abx * hy - aby * hx) * ez;
    const float H1 = fabsf(hx) + fabsf(hy) + fabsf(hz);
    const float AC1 = fabsf(acx) + fabsf(acy) + fabsf(acz), AB1 = fabsf(abx) + fabsf(aby) + fabsf(abz), N1 = fabsf(nx) + fabsf(ny) + fabsf(nz);
    const float eh = 2.0f * err + 16.0f * eps * H1;
    float E2 = AC1 * eh, E3 = AB1 * eh, EG = N1 * eh, ED = 8.0f * eps * N1;
    if (PK) {
        const float dN = 2.5f * (qab * AC1 + qac * AB1) + 4.0f * eps * (AB1 * AC1);
        E2 += 2.5f * qac * H1;
        E3 += 2.5f * qab * H1;
        EG += dN * H1;
        ED += dN;
    }
    if (backface && Dcr < -ED) return HXR_TF_MISS;  // dot(d, N) = -Dcr is certainly positive: culled
    const float aD = fabsf(Dcr);
    if (!(aD > ED)) return HXR_TF_MAYBE;  // grazing (or a degenerate triangle): not even the sign is known
    const bool neg = Dcr < 0;
    const float a2 = neg ? -c2 : c2, a3 = neg ? -c3 : c3, aG = neg ? -Ng : Ng;
    const float a1 = aD - a2 - a3;
    const float E1 = ED + E2 + E3 + 4.0f * eps * (aD + fabsf(a2) + fabsf(a3));
    if (a2 < -E2 || a3 < -E3 || a1 < -E1 || aG < -EG) return HXR_TF_MISS;
    if (aG - EG > tbest * (aD + ED)) return HXR_TF_MISS;  // certainly beyond the best known hit
    const float den = aD - ED;
    if (a2 >= E2 && a3 >= E3 && a1 >= E1 && aG >= EG && den > 1e-11f) {
        ghi = (aG + EG) / den * (1.0f + 4.0f * eps);
        return HXR_TF_CERTAIN;
    }
    return HXR_TF_MAYBE;
}

HXR_HD int tri_filter(const TriF32* p, bool backface, float ox, float oy, float oz, float dx, float dy, float dz, float err, float tbest,
                      float& ghi)
{
#if defined(__CUDA_ARCH__)
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(p)), q1 = __ldg(reinterpret_cast<const float4*>(p) + 1),
                 q2 = __ldg(reinterpret_cast<const float4*>(p) + 2);
    const float ax = q0.x, ay = q0.y, az = q0.z, abx = q0.w, aby = q1.x, abz = q1.y, acx = q1.z, acy = q1.w, acz = q2.x, nx = q2.y, ny = q2.z,
                nz = q2.w;
#else
    const float ax = p->A[0], ay = p->A[1], az = p->A[2], abx = p->AB[0], aby = p->AB[1], abz = p->AB[2], acx = p->AC[0], acy = p->AC[1],
                acz = p->AC[2], nx = p->N[0], ny = p->N[1], nz = p->N[2];
#endif
    return tri_filter_core<false>(ax, ay, az, abx, aby, abz, acx, acy, acz, nx, ny, nz, 0.0f, 0.0f, backface, ox, oy, oz, dx, dy, dz, err, tbest, ghi);
}

HXR_HD float f32_from_bits(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
HXR_HD float sext24f(uint32_t x) { return (float)((int32_t)(x << 8) >> 8); }  // low 24 bits as a signed integer, exactly, in float

// the same filter on a 32-byte packed triangle: ONE sector per test instead of two
HXR_HD int tri_filter_packed(const TriPacked* p, bool backface, float ox, float oy, float oz, float dx, float dy, float dz, float err, float tbest,
                             float& ghi)
{
#if defined(__CUDA_ARCH__)
    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(p)), q1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    const float ax = __uint_as_float(q0.x), ay = __uint_as_float(q0.y), az = __uint_as_float(q0.z);
    const uint32_t w0 = q0.w, w1 = q1.x, w2 = q1.y, w3 = q1.z, w4 = q1.w;
#else
    const float ax = p->A[0], ay = p->A[1], az = p->A[2];
    const uint32_t w0 = p->w[0], w1 = p->w[1], w2 = p->w[2], w3 = p->w[3], w4 = p->w[4];
#endif
    const float sab = f32_from_bits(((w0 & 0xFFu) + 1u) << 23), sac = f32_from_bits((((w0 >> 8) & 0xFFu) + 1u) << 23);  // 2^e: e + 127 = byte + 1
    const float abx = sext24f((w0 >> 16) | (w1 << 16)) * sab, aby = (float)((int32_t)w1 >> 8) * sab, abz = sext24f(w2) * sab;
    const float acx = sext24f((w2 >> 24) | (w3 << 8)) * sac, acy = sext24f((w3 >> 16) | (w4 << 16)) * sac, acz = (float)((int32_t)w4 >> 8) * sac;
    const float nx = aby * acz - abz * acy, ny = abz * acx - abx * acz, nz = abx * acy - aby * acx;
    return tri_filter_core<true>(ax, ay, az, abx, aby, abz, acx, acy, acz, nx, ny, nz, 0.5f * sab, 0.5f * sac, backface, ox, oy, oz, dx, dy, dz, err,
                                 tbest, ghi);
}

#if !defined(__CUDA_ARCH__)
// host: pack one vector as m_i * 2^e, e in [-126, 127], |m_i| < 2^23; false if it does not fit (not finite or >= 2^150)
inline bool pack_vec24(const double v[3], int32_t m[3], uint32_t& ebyte)
{
    const double mx = std::fmax(std::fabs(v[0]), std::fmax(std::fabs(v[1]), std::fabs(v[2])));
    if (!(mx < 1e300)) return false;
    int e = -126;
    if (mx > 0) {
        int ex;
        std::frexp(mx, &ex);  // mx < 2^ex
        e = ex - 23;          // mx / 2^e in [2^22, 2^23)
        if (e < -126) e = -126;
        if (std::nearbyint(std::ldexp(mx, -e)) >= 8388608.0) e++;
        if (e > 127) return false;
    }
    for (int k = 0; k < 3; k++) m[k] = (int32_t)std::nearbyint(std::ldexp(v[k], -e));
    ebyte = (uint32_t)(e + 126);
    return true;
}
inline bool pack_tri(const TriTest& t, TriPacked& o)
{
    int32_t a[3], c[3];
    uint32_t ea, ec;
    if (!pack_vec24(t.AB, a, ea) || !pack_vec24(t.AC, c, ec)) return false;
    const uint32_t M = 0xFFFFFFu;
    const uint32_t m0 = (uint32_t)a[0] & M, m1 = (uint32_t)a[1] & M, m2 = (uint32_t)a[2] & M, m3 = (uint32_t)c[0] & M, m4 = (uint32_t)c[1] & M,
                   m5 = (uint32_t)c[2] & M;
    for (int k = 0; k < 3; k++) o.A[k] = (float)t.A[k];
    o.w[0] = ea | (ec << 8) | ((m0 & 0xFFFFu) << 16);
    o.w[1] = (m0 >> 16) | (m1 << 8);
    o.w[2] = m2 | ((m3 & 0xFFu) << 24);
    o.w[3] = (m3 >> 8) | ((m4 & 0xFFFFu) << 16);
    o.w[4] = (m4 >> 16) | (m5 << 8);
    return true;
}
#endif

// all triangles in index order: exactly the reference's brute-force path (src/mesh.cpp:255-262);
// used for meshes so small that a tree walk costs more than it saves
HXR_HD bool mesh_bruteforce(const DMesh& M, const Ray& ray, double gamma_limit, MeshBest& best)
{
    // The box gate (src/mesh.cpp:249) never changes the answer - a hit lies inside the box - it only saves work; for a quad
    // it costs more issue slots than the two triangle tests it guards, and it splits the warp
    double t0, t1;
    if (M.brute != 3 && !mesh_slab(M, ray, gamma_limit, t0, t1)) return false;
    best.gamma = gamma_limit;
    best.tri = -1;
    best.l2 = best.l3 = 0;
    // (the float filter of the walk does not pay here: measured 7 % slower on cornell_box than the exact test on all <= 24 triangles)
    for (int i = 0; i < M.n_tris; i++) tri_test(M.tri_test, M.backface != 0, ray, (uint32_t)i, best);
    return best.tri >= 0;
}

// lean: the caller knows that nobody reads u, v, dNdx, dNdy of this hit (DScene::node_lean): skip their gather
HXR_HD void mesh_fill_hit(const DMesh& M, int gi, const Ray& ray, const MeshBest& best, Hit& info, bool lean = false)
{
    const TriAttr& ta = M.tri_attr[best.tri];
    info.dist = best.gamma;
    info.ip = ray.o + best.gamma * ray.d;
    if (lean) {
        info.u = info.v = 0;
        info.dNdx = info.dNdy = mk3(0, 0, 0);
    } else {
        const TriAttrUv& tu = M.tri_attr_uv[best.tri];
        // uvs[t.t[k]] with the third coordinate dropped: only x and y are used (src/mesh.cpp:203-207)
        const d3 texA = mk3(tu.uv[0][0], tu.uv[0][1], 0), texB = mk3(tu.uv[1][0], tu.uv[1][1], 0), texC = mk3(tu.uv[2][0], tu.uv[2][1], 0);
        const d3 tex = texA + (texB - texA) * best.l2 + (texC - texA) * best.l3;
        info.u = tex.x;
        info.v = tex.y;
        info.dNdx = ld3(tu.dNdx);
        info.dNdy = ld3(tu.dNdy);
    }
    if (M.faceted) {
        info.norm = ld3(ta.gnormal);
    } else {
        const d3 nA = ld3(ta.nrm[0]), nB = ld3(ta.nrm[1]), nC = ld3(ta.nrm[2]);
        info.norm = normalize_m(nA + (nB - nA) * best.l2 + (nC - nA) * best.l3);
    }
    info.geom = gi;
}

// Meshes with at most this many triangles (quads) are tested inline by brute force; everything bigger goes through the
// walk kernel, even a ten-triangle box: inline loops run at the lane occupancy of the rays that hit the mesh's box (measured
// on cornell_box: 3-9 of 32 lanes), the walk's cooperative leaves run full (cornell_box 1400 -> 2268 Mrays/s with 4 instead of 24)
#define HXR_SMALL_MESH 4

// gamma_limit: object-space ray parameter beyond which hits cannot matter to the caller
// (HXR_INF for "no limit"); it only prunes, it never changes which hit wins below it.
template <bool COUNT>
HXR_HD bool mesh_intersect(const DMesh& M, int gi, const Ray& ray, Hit& info, double gamma_limit, TravCounters* cnt)
{
    MeshBest best;
    const bool hit = M.brute ? mesh_bruteforce(M, ray, gamma_limit, best) : mesh_closest<COUNT>(M, ray, gamma_limit, best, cnt);
    if (!hit) return false;
    mesh_fill_hit(M, gi, ray, best, info);
    return true;
}

// ---------------------------------------------------------------- heightfield
HXR_HD float hf_height(const DHeightfield& F, int x, int y)
{
    x = x < F.W - 1 ? x : F.W - 1;
    y = y < F.H - 1 ? y : F.H - 1;
    x = x > 0 ? x : 0;
    y = y > 0 ? y : 0;
    return F.heights[y * F.W + x];
}
HXR_HD float hf_highest(const DHeightfield& F, int x, int y, int k)
{
    x = x < F.W - 1 ? x : F.W - 1;
    y = y < F.H - 1 ? y : F.H - 1;
    x = x > 0 ? x : 0;
    y = y > 0 ? y : 0;
    return F.high_map[(size_t)(y * F.W + x) * 16 + k];
}
HXR_HD d3 hf_normal(const DHeightfield& F, float x, float y)  // src/heightfield.cpp:83-104
{
    int x0 = (int)floorf(x);
    int y0 = (int)floorf(y);
    float p = (x - x0);
    float q = (y - y0);
    int x1 = (x0 + 1 < F.W - 1) ? x0 + 1 : F.W - 1;
    int y1 = (y0 + 1 < F.H - 1) ? y0 + 1 : F.H - 1;
    x0 = x0 < F.W - 1 ? x0 : F.W - 1;
    y0 = y0 < F.H - 1 ? y0 : F.H - 1;
    x0 = x0 > 0 ? x0 : 0;
    y0 = y0 > 0 ? y0 : 0;
    // the weights are float products promoted to double (Vector * double)
    d3 v = ld3(F.normals + 3 * (y0 * F.W + x0)) * (double)((1 - p) * (1 - q)) +
           ld3(F.normals + 3 * (y0 * F.W + x1)) * (double)((p) * (1 - q)) +
           ld3(F.normals + 3 * (y1 * F.W + x0)) * (double)((1 - p) * (q)) +
           ld3(F.normals + 3 * (y1 * F.W + x1)) * (double)((p) * (q));
    return normalize_m(v);
}

HXR_HD bool heightfield_intersect(const DHeightfield& F, int gi, const Ray& ray, Hit& info)
{
    d3 step = ray.d;
    double distHoriz = sqrt(step.x * step.x + step.z * step.z);
    step = div3(step, distHoriz);
    double dist = bbox_closest_intersection(F.bbmin, F.bbmax, ray);
    d3 p = ray.o + ray.d * (dist + 1e-6);
    double mx = 1.0 / ray.d.x;
    double mz = 1.0 / ray.d.z;
    // the loop is bounded: every iteration advances p by at least ~1e-6 along a ray that must
    // leave a W x H box; the cap only protects against NaN-poisoned vertical rays.
    for (int guard = 0; guard < 1 << 22 && bbox_inside(F.bbmin, F.bbmax, p); guard++) {
        int x0 = (int)floor(p.x);
        int z0 = (int)floor(p.z);
        if (x0 < 0 || x0 >= F.W || z0 < 0 || z0 >= F.H) break;
        if (F.use_opt) {
            int k = 1;
            while (k < F.max_k && p.y + step.y * (1 << k) > hf_highest(F, x0, z0, k)) k++;
            k--;
            if (k > 0) {
                p = p + step * (double)(1 << k);
                continue;
            }
        }
        double lx = ray.d.x > 0 ? (ceil(p.x) - p.x) * mx : (floor(p.x) - p.x) * mx;
        double lz = ray.d.z > 0 ? (ceil(p.z) - p.z) * mz : (floor(p.z) - p.z) * mz;
        double lmin = (lz < lx) ? lz : lx;  // std::min(lx, lz)
        d3 p_next = p + step * (lmin + 1e-6);
        double ymin = (p_next.y < p.y) ? p_next.y : p.y;
        if (ymin < F.max_h[z0 * F.W + x0]) {
            double closestDist = HXR_INF;
            d3 A = mk3(x0, hf_height(F, x0, z0), z0);
            d3 B = mk3(x0 + 1, hf_height(F, x0 + 1, z0), z0);
            d3 C = mk3(x0 + 1, hf_height(F, x0 + 1, z0 + 1), z0 + 1);
            d3 D = mk3(x0, hf_height(F, x0, z0 + 1), z0 + 1);
            bool b1 = intersect_triangle_fast(ray, A, B, D, closestDist);
            bool b2 = intersect_triangle_fast(ray, B, C, D, closestDist);
            if (b1 || b2) {
                info.dist = closestDist;
                info.ip = ray.o + ray.d * closestDist;
                info.norm = hf_normal(F, (float)info.ip.x, (float)info.ip.z);
                info.u = info.ip.x / F.W;
                info.v = info.ip.z / F.H;
                info.dNdx = mk3(1, 0, 0);
                info.dNdy = mk3(0, 0, 1);
                info.geom = gi;
                return true;
            }
        }
        p = p_next;
    }
    return false;
}

// ---------------------------------------------------------------- dispatch + CSG
#define HXR_CSG_MAX_CROSSINGS 30
#define HXR_CSG_MAX_NESTING 3

template <int DEPTH, bool COUNT>
HXR_HD_NOINLINE bool geom_intersect(const DScene& sc, int gi, const Ray& ray, Hit& info, double gamma_limit, TravCounters* cnt);

HXR_HD bool csg_inside(int op, bool inA, bool inB)  // src/geometry.h:123-136
{
    return op == HXR_CSG_UNION ? (inA || inB) : (op == HXR_CSG_INTER ? (inA && inB) : (inA && !inB));
}

// k-th boundary crossing of `child` along the ray (findAllIntersections re-shoots from
// ip + dir*1e-6, geometry.cpp:149-165); returns false if there are fewer than k+1 crossings.
template <int DEPTH, bool COUNT>
HXR_HD bool csg_nth_crossing(const DScene& sc, int child, const Ray& ray, int k, Hit& info, TravCounters* cnt)
{
    Ray cur = ray;
    for (int c = 0; c <= k; c++) {
        info.dist = HXR_INF;
        if (!geom_intersect<DEPTH + 1, COUNT>(sc, child, cur, info, HXR_INF, cnt)) return false;
        cur.o = info.ip + ray.d * 1e-6;
    }
    return true;
}

template <int DEPTH, bool COUNT>
HXR_HD bool csg_intersect(const DScene& sc, const hxr_geometry& g, int gi, const Ray& ray, Hit& info, TravCounters* cnt)
{
    const int op = g.a, left = g.b, right = g.c;
    // pass 1: distances of all crossings of both children, kept as (dist, child, ordinal, geom)
    double cd[2 * HXR_CSG_MAX_CROSSINGS];
    uint8_t cchild[2 * HXR_CSG_MAX_CROSSINGS], cord[2 * HXR_CSG_MAX_CROSSINGS], cleft[2 * HXR_CSG_MAX_CROSSINGS];
    int n = 0, cnt2[2] = {0, 0};
    for (int side = 0; side < 2; side++) {
        const int child = side ? right : left;
        Ray cur = ray;
        for (int c = 0; c < HXR_CSG_MAX_CROSSINGS; c++) {
            Hit h;
            h.dist = HXR_INF;
            if (!geom_intersect<DEPTH + 1, COUNT>(sc, child, cur, h, HXR_INF, cnt)) break;
            cd[n] = distance3(ray.o, h.ip);
            cchild[n] = (uint8_t)side;
            cord[n] = (uint8_t)c;
            cleft[n] = (uint8_t)(h.geom == left);
            n++;
            cnt2[side]++;
            cur.o = h.ip + ray.d * 1e-6;
        }
    }
    // order by distance (insertion sort keeps left-before-right on ties, like std::sort on <=16 items)
    for (int i = 1; i < n; i++) {
        double kd = cd[i];
        uint8_t kc = cchild[i], ko = cord[i], kl = cleft[i];
        int j = i - 1;
        while (j >= 0 && kd < cd[j]) {
            cd[j + 1] = cd[j]; cchild[j + 1] = cchild[j]; cord[j + 1] = cord[j]; cleft[j + 1] = cleft[j];
            j--;
        }
        cd[j + 1] = kd; cchild[j + 1] = kc; cord[j + 1] = ko; cleft[j + 1] = kl;
    }
    bool inA = cnt2[0] % 2, inB = cnt2[1] % 2;
    const bool initial = csg_inside(op, inA, inB);
    for (int i = 0; i < n; i++) {
        if (cleft[i]) inA = !inA; else inB = !inB;
        if (csg_inside(op, inA, inB) != initial) {
            // pass 2: regenerate the winning crossing's full record
            if (!csg_nth_crossing<DEPTH, COUNT>(sc, cchild[i] ? right : left, ray, cord[i], info, cnt)) return false;
            info.dist = distance3(ray.o, info.ip);
            info.norm = faceforward(ray.d, info.norm);
            info.geom = gi;
            return true;
        }
    }
    return false;
}

template <int DEPTH, bool COUNT>
HXR_HD_NOINLINE bool geom_intersect(const DScene& sc, int gi, const Ray& ray, Hit& info, double gamma_limit, TravCounters* cnt)
{
    const hxr_geometry& g = sc.geoms[gi];
    switch (g.type) {
        case HXR_GEOM_PLANE: return plane_intersect(g, gi, ray, info);
        case HXR_GEOM_SPHERE: return sphere_intersect(g, gi, ray, info);
        case HXR_GEOM_CUBE: return cube_intersect(g, gi, ray, info);
        case HXR_GEOM_MESH: return mesh_intersect<COUNT>(sc.meshes[g.a], gi, ray, info, gamma_limit, cnt);
        case HXR_GEOM_HEIGHTFIELD: return heightfield_intersect(sc.hfs[g.a], gi, ray, info);
        case HXR_GEOM_CSG:
            if constexpr (DEPTH < HXR_CSG_MAX_NESTING) return csg_intersect<DEPTH, COUNT>(sc, g, gi, ray, info, cnt);
            else return false;
    }
    return false;
}

// ---------------------------------------------------------------- instancing + lights
// world_limit: world-space distance beyond which a hit cannot matter (prunes mesh traversal only).
// SIMPLE: the scene's inline nodes are only planes, spheres, cubes and brute-force meshes (DScene::simple_inline); the
// kernels are compiled twice so that such scenes do not carry the CSG / heightfield / tree-walk code (its stack frame
// and registers) through every ray.
template <bool COUNT, bool SIMPLE>
HXR_HD bool node_intersect(const DScene& sc, const hxr_node& nd, const Ray& ray, Hit& info, double world_limit, TravCounters* cnt)
{
    const bool ident = (nd.pad & HXR_NODE_IDENT) != 0;  // v * I == v: skip the mat-vecs of an untransformed node
    Ray t;
    t.o = ident ? ray.o - ld3(nd.T.offset) : mul_vm(ray.o - ld3(nd.T.offset), nd.T.inv);
    const d3 dl = ident ? ray.d : mul_vm(ray.d, nd.T.inv);
    t.d = normalize_f(dl);
    t.depth = ray.depth;
    t.flags = ray.flags;
    bool hit;
    if (SIMPLE) {
        const hxr_geometry& g = sc.geoms[nd.geom];
        if (g.type == HXR_GEOM_PLANE) hit = plane_intersect(g, nd.geom, t, info);
        else if (g.type == HXR_GEOM_SPHERE) hit = sphere_intersect(g, nd.geom, t, info);
        else if (g.type == HXR_GEOM_CUBE) hit = cube_intersect(g, nd.geom, t, info);
        else {
            const DMesh& M = sc.meshes[g.a];
            MeshBest best;
            double gamma_limit = HXR_INF;
            if (world_limit < HXR_INF) gamma_limit = world_limit / length(ident ? t.d : mul_vm(t.d, nd.T.m)) * (1.0 + 1e-9) + 1e-9;  // prunes only
            hit = mesh_bruteforce(M, t, gamma_limit, best);
            if (hit) mesh_fill_hit(M, nd.geom, t, best, info);
        }
    } else {
        double gamma_limit = HXR_INF;
        if (world_limit < HXR_INF) {
            // world distance of the object-space point o + g*d is g * |d * m|
            const double k = length(ident ? t.d : mul_vm(t.d, nd.T.m));
            gamma_limit = world_limit / k * (1.0 + 1e-9) + 1e-9;
        }
        hit = geom_intersect<0, COUNT>(sc, nd.geom, t, info, gamma_limit, cnt);
    }
    if (!hit) return false;
    info.ip = (ident ? info.ip : mul_vm(info.ip, nd.T.m)) + ld3(nd.T.offset);
    info.norm = normalize_m(ident ? info.norm : mul_vm(info.norm, nd.T.inv_t));
    info.dist = distance3(ray.o, info.ip);
    return true;
}

// returns +1 / -1 / 0 and shortens `dist` like RectLight::intersect (src/lights.cpp:53-73)
HXR_HD int light_intersect(const hxr_light& L, const Ray& ray, double& dist)
{
    if (L.type != HXR_LIGHT_RECT) return 0;
    d3 o = mul_vm(ray.o - ld3(L.T.offset), L.T.inv);
    d3 d = normalize_f(mul_vm(ray.d, L.T.inv));
    if (fabs(d.y) < 1e-12) return 0;
    double len = -(o.y / d.y);
    if (len < 0) return 0;
    d3 p = o + d * len;
    if (fabs(p.x) < 0.5 && fabs(p.z) < 0.5) {
        double distance = length((mul_vm(p, L.T.m) + ld3(L.T.offset)) - ray.o);
        if (distance < dist) {
            dist = distance;
            return (o.y < 0) ? +1 : -1;
        }
    }
    return 0;
}

}  // namespace hxr
