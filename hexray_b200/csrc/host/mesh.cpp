// mesh.cpp — Mesh element: OBJ ingest, per-triangle precomputation, procedural meshes.
//   loadFromOBJ        semantics of src/mesh.cpp:265-358 (1-based indices with a (0,0,0) sentinel in
//                      slot 0, missing vt/vn index = 0, fan triangulation, CRLF tolerated)
//   prepareTriangles   src/mesh.cpp:360-396 (AB, AC, AB^AC, gnormal, dNdx/dNdy from a 2x2 uv solve)
//   beginRender        src/mesh.cpp:49-87 (recenter, bounds, autoSmooth, faceted fallback); the
//                      reference's median-split KD build is replaced by the library's SAH build
//                      at upload time (csrc/host/kdtree.cpp)
#include <algorithm>
#include "cache.h"
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "scene.h"

namespace hxr {
namespace host {

static double tokDouble(const char* b, const char* e)
{
    if (b == e) return 0;
    char buf[64];
    size_t n = (size_t)(e - b) < sizeof(buf) - 1 ? (size_t)(e - b) : sizeof(buf) - 1;
    memcpy(buf, b, n);
    buf[n] = 0;
    char* end;
    double v = strtod(buf, &end);
    return end == buf ? 0 : v;
}

// "v", "v/t", "v//n", "v/t/n"
static void parseCorner(const char* b, const char* e, int& v, int& t, int& n)
{
    int out[3] = {0, 0, 0};
    int k = 0;
    const char* p = b;
    while (p <= e && k < 3) {
        const char* q = p;
        while (q < e && *q != '/') q++;
        if (q > p) {
            char buf[32];
            size_t len = (size_t)(q - p) < sizeof(buf) - 1 ? (size_t)(q - p) : sizeof(buf) - 1;
            memcpy(buf, p, len);
            buf[len] = 0;
            int x;
            out[k] = sscanf(buf, "%d", &x) == 1 ? x : 0;
        }
        k++;
        if (q >= e) break;
        p = q + 1;
    }
    v = out[0]; t = out[1]; n = out[2];
}

bool Mesh::loadFromOBJ(const char* filename)
{
    // a big file parsed before comes back from its binary cache (host/cache.h): O(read) instead of O(tokens)
    {
        ObjArrays a;
        if (loadObjCache(filename, a)) {
            auto fill = [](std::vector<Vec3>& dst, const std::vector<double>& src) {
                dst.resize(src.size() / 3);
                for (size_t i = 0; i < dst.size(); i++) dst[i] = Vec3(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
            };
            fill(vertices, a.vertices);
            fill(normals, a.normals);
            fill(uvs, a.uvs);
            triangles.resize(a.tris.size() / 9);
            for (size_t t = 0; t < triangles.size(); t++) {
                memset(&triangles[t], 0, sizeof(hxr_triangle));
                for (int k = 0; k < 3; k++) {
                    triangles[t].v[k] = a.tris[9 * t + k];
                    triangles[t].n[k] = a.tris[9 * t + 3 + k];
                    triangles[t].t[k] = a.tris[9 * t + 6 + k];
                }
            }
            prepareTriangles();
            return true;
        }
    }
    FILE* f = fopen(filename, "rt");
    if (!f) return false;
    vertices.assign(1, Vec3(0, 0, 0));
    normals.assign(1, Vec3(0, 0, 0));
    uvs.assign(1, Vec3(0, 0, 0));
    triangles.clear();
    char line[10000];
    std::vector<std::pair<const char*, const char*>> tok;
    while (fgets(line, sizeof line, f)) {
        if (line[0] == '#') continue;
        tok.clear();
        for (char* p = line; *p;) {
            while (*p && isspace((unsigned char)*p)) p++;
            if (!*p) break;
            char* q = p;
            while (*q && !isspace((unsigned char)*q)) q++;
            tok.emplace_back(p, q);
            p = q;
        }
        if (tok.empty()) continue;
        const size_t tl = (size_t)(tok[0].second - tok[0].first);
        auto is = [&](const char* s) { return tl == strlen(s) && !strncmp(tok[0].first, s, tl); };
        auto num = [&](size_t i) { return i < tok.size() ? tokDouble(tok[i].first, tok[i].second) : 0.0; };
        if (is("v")) vertices.push_back(Vec3(num(1), num(2), num(3)));
        else if (is("vn")) normals.push_back(Vec3(num(1), num(2), num(3)));
        else if (is("vt")) uvs.push_back(Vec3(num(1), num(2), 0));
        else if (is("f")) {
            for (size_t i = 0; i + 3 < tok.size(); i++) {
                hxr_triangle T;
                memset(&T, 0, sizeof T);
                parseCorner(tok[1].first, tok[1].second, T.v[0], T.t[0], T.n[0]);
                parseCorner(tok[2 + i].first, tok[2 + i].second, T.v[1], T.t[1], T.n[1]);
                parseCorner(tok[3 + i].first, tok[3 + i].second, T.v[2], T.t[2], T.n[2]);
                triangles.push_back(T);
            }
        }
    }
    fclose(f);
    // the reference indexes its arrays unchecked; a file with out-of-range (or negative,
    // unsupported) indices is rejected here instead of reading out of bounds
    for (const auto& T : triangles)
        for (int k = 0; k < 3; k++)
            if (T.v[k] < 0 || T.v[k] >= (int)vertices.size() || T.n[k] < 0 || T.n[k] >= (int)normals.size() ||
                T.t[k] < 0 || T.t[k] >= (int)uvs.size())
                return false;
    if (triangles.size() >= 50000 && cacheEnabled()) {  // (small files parse in milliseconds)
        ObjArrays a;
        auto flat = [](std::vector<double>& dst, const std::vector<Vec3>& src) {
            dst.resize(src.size() * 3);
            for (size_t i = 0; i < src.size(); i++) { dst[3 * i] = src[i].x; dst[3 * i + 1] = src[i].y; dst[3 * i + 2] = src[i].z; }
        };
        flat(a.vertices, vertices);
        flat(a.normals, normals);
        flat(a.uvs, uvs);
        a.tris.resize(triangles.size() * 9);
        for (size_t t = 0; t < triangles.size(); t++)
            for (int k = 0; k < 3; k++) {
                a.tris[9 * t + k] = triangles[t].v[k];
                a.tris[9 * t + 3 + k] = triangles[t].n[k];
                a.tris[9 * t + 6 + k] = triangles[t].t[k];
            }
        storeObjCache(filename, a);
    }
    prepareTriangles();
    return true;
}

static void store3(double* d, const Vec3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

// x*A + y*B = C in 2-D (Cramer); a singular uv mapping yields non-finite tangents, as in the reference
static void solve2D(const Vec3& A, const Vec3& B, const Vec3& C, double& x, double& y)
{
    const double det = A.x * B.y - A.y * B.x;
    x = (C.x * B.y - C.y * B.x) / det;
    y = (A.x * C.y - A.y * C.x) / det;
}

void Mesh::prepareTriangles()
{
    if (normals.size() <= 1) faceted = true;
    for (auto& t : triangles) {
        const Vec3 &A = vertices[t.v[0]], &B = vertices[t.v[1]], &C = vertices[t.v[2]];
        const Vec3 AB = B - A, AC = C - A;
        const Vec3 N = cross(AB, AC);
        Vec3 g = N;
        g.normalize();
        store3(t.ab, AB);
        store3(t.ac, AC);
        store3(t.ab_cross_ac, N);
        store3(t.gnormal, g);
        const Vec3 &tA = uvs[t.t[0]], &tB = uvs[t.t[1]], &tC = uvs[t.t[2]];
        const Vec3 texAB = tB - tA, texAC = tC - tA;
        double px, py, qx, qy;
        solve2D(texAB, texAC, Vec3(1, 0, 0), px, qx);
        solve2D(texAB, texAC, Vec3(0, 1, 0), py, qy);
        Vec3 dx = px * AB + qx * AC, dy = py * AB + qy * AC;
        dx.normalize();
        dy.normalize();
        store3(t.dndx, dx);
        store3(t.dndy, dy);
    }
}

void Mesh::computeBoundingGeometry()
{
    bbmin = Vec3(+1e99, +1e99, +1e99);
    bbmax = Vec3(-1e99, -1e99, -1e99);
    for (size_t i = 1; i < vertices.size(); i++)
        for (int a = 0; a < 3; a++) {
            bbmin[a] = std::min(bbmin[a], vertices[i][a]);
            bbmax[a] = std::max(bbmax[a], vertices[i][a]);
        }
}

void Mesh::beginRender()
{
    if (recenter && vertices.size() > 1) {
        Vec3 c(0, 0, 0);
        for (size_t i = 1; i < vertices.size(); i++) c = c + vertices[i];
        c = c * (1.0 / double(vertices.size() - 1));
        for (size_t i = 1; i < vertices.size(); i++) vertices[i] = vertices[i] + (c * -1.0);
    }
    computeBoundingGeometry();
    if (normals.size() <= 1 && autoSmooth) {
        // unweighted sum of unit face normals per vertex
        normals.assign(vertices.size(), Vec3(0, 0, 0));
        for (auto& t : triangles)
            for (int j = 0; j < 3; j++) {
                t.n[j] = t.v[j];
                normals[t.n[j]] = normals[t.n[j]] + Vec3(t.gnormal[0], t.gnormal[1], t.gnormal[2]);
            }
        for (size_t i = 1; i < normals.size(); i++)
            if (normals[i].lengthSqr() > 1e-9) normals[i].normalize();
        faceted = false;
    }
    if (normals.size() <= 1) faceted = true;
}

// ------------------------------------------------------------------ procedural meshes (C5)
static inline uint32_t lattice(uint32_t x, uint32_t y, uint32_t seed)
{
    uint32_t h = x * 0x8DA6B343u ^ y * 0xD8163841u ^ seed * 0xCB1AB31Fu;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
static double valueNoise(double x, double y, uint32_t seed)
{
    const double fx = std::floor(x), fy = std::floor(y);
    const int ix = (int)fx, iy = (int)fy;
    double tx = x - fx, ty = y - fy;
    tx = tx * tx * (3 - 2 * tx);
    ty = ty * ty * (3 - 2 * ty);
    auto v = [&](int a, int b) { return lattice((uint32_t)a, (uint32_t)b, seed) * (1.0 / 4294967296.0); };
    const double a = v(ix, iy), b = v(ix + 1, iy), c = v(ix, iy + 1), d = v(ix + 1, iy + 1);
    return (a + (b - a) * tx) * (1 - ty) + (c + (d - c) * tx) * ty;
}
static double fbm(double x, double y, uint32_t seed)
{
    double amp = 0.5, f = 1, s = 0;
    for (int o = 0; o < 5; o++) {
        s += amp * valueNoise(x * f, y * f, seed + (uint32_t)o * 101u);
        amp *= 0.5;
        f *= 2.03;
    }
    return s;
}

// terrain: gridSide^2 vertices over [-500,500]^2 in xz, y = 60*fbm(x/250, z/250);
// 2*(gridSide-1)^2 triangles, per-vertex normals, uv = grid coordinates
void Mesh::generateTerrain(int n, uint64_t seed)
{
    if (n < 2) n = 2;
    vertices.assign(1, Vec3(0, 0, 0));
    normals.assign(1, Vec3(0, 0, 0));
    uvs.assign(1, Vec3(0, 0, 0));
    vertices.reserve((size_t)n * n + 1);
    uvs.reserve((size_t)n * n + 1);
    const double step = 1000.0 / (n - 1);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
            const double x = -500.0 + i * step, z = -500.0 + j * step;
            vertices.push_back(Vec3(x, 60.0 * fbm(x / 250.0 + 7.0, z / 250.0 + 3.0, (uint32_t)seed), z));
            uvs.push_back(Vec3((double)i, (double)j, 0));
        }
    normals.assign(vertices.size(), Vec3(0, 0, 0));
    triangles.clear();
    triangles.reserve((size_t)2 * (n - 1) * (n - 1));
    auto vid = [&](int i, int j) { return 1 + j * n + i; };
    for (int j = 0; j + 1 < n; j++)
        for (int i = 0; i + 1 < n; i++) {
            const int q[4] = {vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)};
            const int idx[2][3] = {{q[0], q[3], q[1]}, {q[1], q[3], q[2]}};  // +y facing
            for (int k = 0; k < 2; k++) {
                hxr_triangle T;
                memset(&T, 0, sizeof T);
                for (int c = 0; c < 3; c++) T.v[c] = T.n[c] = T.t[c] = idx[k][c];
                triangles.push_back(T);
            }
        }
    prepareTriangles();
    for (const auto& t : triangles)
        for (int c = 0; c < 3; c++) normals[t.n[c]] = normals[t.n[c]] + Vec3(t.gnormal[0], t.gnormal[1], t.gnormal[2]);
    for (size_t i = 1; i < normals.size(); i++) normals[i].normalize();
    faceted = false;
}

bool Mesh::saveOBJ(const char* filename) const
{
    FILE* f = fopen(filename, "wt");
    if (!f) return false;
    std::vector<char> buf(1 << 20);
    setvbuf(f, buf.data(), _IOFBF, buf.size());
    for (size_t i = 1; i < vertices.size(); i++) fprintf(f, "v %.17g %.17g %.17g\n", vertices[i].x, vertices[i].y, vertices[i].z);
    for (size_t i = 1; i < uvs.size(); i++) fprintf(f, "vt %.17g %.17g\n", uvs[i].x, uvs[i].y);
    for (size_t i = 1; i < normals.size(); i++) fprintf(f, "vn %.17g %.17g %.17g\n", normals[i].x, normals[i].y, normals[i].z);
    const bool hasN = normals.size() > 1, hasT = uvs.size() > 1;
    for (const auto& t : triangles) {
        fputs("f", f);
        for (int k = 0; k < 3; k++) {
            if (hasN && hasT) fprintf(f, " %d/%d/%d", t.v[k], t.t[k], t.n[k]);
            else if (hasT) fprintf(f, " %d/%d", t.v[k], t.t[k]);
            else if (hasN) fprintf(f, " %d//%d", t.v[k], t.n[k]);
            else fprintf(f, " %d", t.v[k]);
        }
        fputs("\n", f);
    }
    bool ok = !ferror(f);
    fclose(f);
    return ok;
}

// soup: nTriangles random triangles, centroids uniform in [-500,500]^3, edge ~ N(2, 0.5): worst-case incoherence
void Mesh::generateSoup(int64_t nTris, uint64_t seed)
{
    std::mt19937_64 gen(seed);
    std::uniform_real_distribution<double> pos(-500.0, 500.0), unit(-1.0, 1.0);
    std::normal_distribution<double> edge(2.0, 0.5);
    vertices.assign(1, Vec3(0, 0, 0));
    normals.assign(1, Vec3(0, 0, 0));
    uvs.assign(1, Vec3(0, 0, 0));
    uvs.push_back(Vec3(0, 0, 0));
    uvs.push_back(Vec3(1, 0, 0));
    uvs.push_back(Vec3(0, 1, 0));
    triangles.clear();
    triangles.reserve((size_t)nTris);
    for (int64_t i = 0; i < nTris; i++) {
        const Vec3 c(pos(gen), pos(gen), pos(gen));
        const double e = std::max(0.2, edge(gen));
        hxr_triangle T;
        memset(&T, 0, sizeof T);
        for (int k = 0; k < 3; k++) {
            Vec3 d(unit(gen), unit(gen), unit(gen));
            vertices.push_back(c + d * (e * 0.5));
            T.v[k] = (int)vertices.size() - 1;
            T.t[k] = 1 + k;
        }
        triangles.push_back(T);
    }
    prepareTriangles();
    faceted = true;
}

}  // namespace host
}  // namespace hxr
