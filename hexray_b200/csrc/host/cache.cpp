// cache.cpp — see cache.h.
#include "cache.h"
#include <cerrno>
#include <chrono>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

namespace hxr {
namespace host {

bool cacheEnabled()
{
    const char* e = getenv("HXR_CACHE");
    return !(e && atoi(e) == 0);
}

std::string cacheDir()
{
    if (!cacheEnabled()) return "";
    const char* e = getenv("HXR_CACHE_DIR");
    std::string d = e && *e ? e : "/tmp/hexray_b200_cache";
    struct stat st;
    if (stat(d.c_str(), &st) != 0 && mkdir(d.c_str(), 0777) != 0 && stat(d.c_str(), &st) != 0) return "";
    return d;
}

// 8 bytes per step, multiply-xorshift mixing (not cryptographic: it only has to tell meshes apart)
uint64_t hashBytes(const void* data, size_t n, uint64_t seed)
{
    const unsigned char* p = (const unsigned char*)data;
    uint64_t h = seed ^ (n * 0x9E3779B97F4A7C15ull);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 0xD6E8FEB86659FD93ull;
        h ^= h >> 32;
    }
    uint64_t w = 0;
    if (i < n) memcpy(&w, p + i, n - i);
    h = (h ^ w) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 29;
    h *= 0x94D049BB133111EBull;
    h ^= h >> 32;
    return h;
}

namespace {

struct KdHeader {
    char magic[8];  // "HXRKD\0\0\0"
    uint32_t version, n_triangles;
    uint64_t key;
    uint64_t n_blocks, n_leaf_tris, leaves;
    uint32_t max_depth, pad;
    double build_ms;
    uint64_t content;  // hashBytes over the blocks, then the leaf lists: a damaged file is rebuilt, not walked
};
const uint32_t KD_VERSION = 4;

uint64_t kdContentHash(const KdTree& kd)
{
    uint64_t h = hashBytes(kd.blocks.data(), kd.blocks.size() * sizeof(KdBlock), 0x4B444331ull);
    return hashBytes(kd.leafTris.data(), kd.leafTris.size() * sizeof(uint32_t), h);
}

std::string kdPath(uint64_t key)
{
    const std::string d = cacheDir();
    if (d.empty()) return "";
    char b[64];
    snprintf(b, sizeof b, "/kd_%016llx.bin", (unsigned long long)key);
    return d + b;
}

bool writeAtomically(const std::string& path, const std::vector<std::pair<const void*, size_t>>& parts)
{
    char tmp[64];
    snprintf(tmp, sizeof tmp, ".tmp%d", (int)getpid());
    const std::string t = path + tmp;
    FILE* f = fopen(t.c_str(), "wb");
    if (!f) return false;
    bool ok = true;
    for (const auto& p : parts) ok = ok && (p.second == 0 || fwrite(p.first, 1, p.second, f) == p.second);
    ok = (fclose(f) == 0) && ok;
    if (ok) ok = rename(t.c_str(), path.c_str()) == 0;
    if (!ok) unlink(t.c_str());
    return ok;
}

}  // namespace

uint64_t meshContentKey(const hxr_mesh& m, const KdBuildParams& P)
{
    uint64_t h = 0x48585232ull;
    h = hashBytes(m.vertices, (size_t)m.n_vertices * 3 * sizeof(double), h);
    // the triangles' vertex indices (their derived vectors follow from the vertices)
    std::vector<int32_t> idx((size_t)m.n_triangles * 3);
    for (int t = 0; t < m.n_triangles; t++)
        for (int k = 0; k < 3; k++) idx[(size_t)t * 3 + k] = m.triangles[t].v[k];
    h = hashBytes(idx.data(), idx.size() * sizeof(int32_t), h);
    h = hashBytes(m.bbox_min, sizeof m.bbox_min, h);
    h = hashBytes(m.bbox_max, sizeof m.bbox_max, h);
    const float fp[3] = {P.traversalCost, P.intersectCost, P.emptyBonus};
    const int ip[3] = {P.maxLeafSize, P.maxDepth, P.binnedAbove};
    h = hashBytes(fp, sizeof fp, h);
    h = hashBytes(ip, sizeof ip, h);
    return h;
}

bool loadKdTree(uint64_t key, const hxr_mesh& mesh, KdTree& out)
{
    const std::string path = kdPath(key);
    if (path.empty()) return false;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    KdHeader h;
    bool ok = fread(&h, sizeof h, 1, f) == 1 && !memcmp(h.magic, "HXRKD\0\0", 8) && h.version == KD_VERSION && h.key == key &&
              h.n_triangles == (uint32_t)mesh.n_triangles && h.n_blocks < (1ull << 31) && h.n_leaf_tris < (1ull << 33);
    if (ok) {
        out.nodes.clear();
        out.blocks.resize(h.n_blocks);
        out.leafTris.resize(h.n_leaf_tris);
        ok = (h.n_blocks == 0 || fread(out.blocks.data(), sizeof(KdBlock), h.n_blocks, f) == h.n_blocks) &&
             (h.n_leaf_tris == 0 || fread(out.leafTris.data(), sizeof(uint32_t), h.n_leaf_tris, f) == h.n_leaf_tris);
        out.maxDepth = h.max_depth;
        out.leaves = h.leaves;
        out.buildMs = h.build_ms;
        ok = ok && kdContentHash(out) == h.content;
    }
    fclose(f);
    if (!ok) {
        out.blocks.clear();
        out.leafTris.clear();
    }
    return ok;
}

void storeKdTree(uint64_t key, const hxr_mesh& mesh, const KdTree& kd)
{
    const std::string path = kdPath(key);
    if (path.empty()) return;
    KdHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "HXRKD\0\0", 8);
    h.version = KD_VERSION;
    h.n_triangles = (uint32_t)mesh.n_triangles;
    h.key = key;
    h.n_blocks = kd.blocks.size();
    h.n_leaf_tris = kd.leafTris.size();
    h.leaves = kd.leaves;
    h.max_depth = kd.maxDepth;
    h.build_ms = kd.buildMs;
    h.content = kdContentHash(kd);
    writeAtomically(path, {{&h, sizeof h}, {kd.blocks.data(), kd.blocks.size() * sizeof(KdBlock)}, {kd.leafTris.data(), kd.leafTris.size() * sizeof(uint32_t)}});
}

void cachedKdTree(const hxr_mesh& mesh, const KdBuildParams& params, KdTree& out, const char** how)
{
    const char* dummy;
    if (!how) how = &dummy;
    // small meshes build in milliseconds: not worth a file
    if (!cacheEnabled() || mesh.n_triangles < 100000 || cacheDir().empty()) {
        buildKdTree(mesh, params, out);
        *how = "built";
        return;
    }
    const uint64_t key = meshContentKey(mesh, params);
    if (loadKdTree(key, mesh, out)) { *how = "cache"; return; }
    // build it, unless another process of this box is at it already: then wait for its file
    // The lock file names its builder (pid): a lock whose process no longer exists - a run killed in the middle of its build -
    // is stale, and is taken over instead of being waited for.
    const std::string lock = kdPath(key) + ".lock";
    auto deadLockOwner = [&]() -> long {  // the pid the lock names if that process no longer exists, else 0
        FILE* f = fopen(lock.c_str(), "r");
        if (!f) return 0;
        long pid = 0;
        const int got = fscanf(f, "%ld", &pid);
        fclose(f);
        if (got != 1 || pid <= 0) return 0;  // (being written this instant, or a lock without a pid: the time-out below covers it)
        return (kill((pid_t)pid, 0) != 0 && errno == ESRCH) ? pid : 0;
    };
    auto takeLock = [&]() -> bool {
        const int fd = open(lock.c_str(), O_CREAT | O_EXCL | O_WRONLY, 0666);
        if (fd < 0) return false;
        char b[32];
        const int n = snprintf(b, sizeof b, "%ld\n", (long)getpid());
        if (write(fd, b, (size_t)n) != n) { /* an empty lock still excludes; its owner just cannot be probed */ }
        close(fd);
        return true;
    };
    bool mine = takeLock();
    if (!mine) {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            std::this_thread::sleep_for(std::chrono::milliseconds(50));
            if (loadKdTree(key, mesh, out)) { *how = "waited"; return; }
            struct stat st;
            const bool lockGone = stat(lock.c_str(), &st) != 0;
            if (lockGone) {
                if (loadKdTree(key, mesh, out)) { *how = "waited"; return; }
                mine = takeLock();  // the builder left without publishing: build it here (unless somebody else was quicker)
                if (mine) break;
                continue;
            }
            const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            const long dead = deadLockOwner();
            if (dead || waited > 300.0) {
                // stale (or implausibly slow): whoever removes the lock and takes a new one builds. The second look keeps a
                // waiter from removing the fresh lock of another waiter that was quicker (the worst case is a tree built twice)
                if (waited > 300.0 || deadLockOwner() == dead) unlink(lock.c_str());
                mine = takeLock();
                if (mine) break;
            }
        }
    }
    buildKdTree(mesh, params, out);
    storeKdTree(key, mesh, out);
    unlink(lock.c_str());
    *how = "built";
}

// ---- parsed OBJ
namespace {
struct ObjHeader {
    char magic[8];  // "HXROBJ\0\0"
    uint32_t version, pad;
    uint64_t key, n_vertices, n_normals, n_uvs, n_tris;
    uint64_t content;  // hashBytes over the four arrays in file order
};
const uint32_t OBJ_VERSION = 2;

uint64_t objContentHash(const ObjArrays& a)
{
    uint64_t h = hashBytes(a.vertices.data(), a.vertices.size() * 8, 0x4F424A32ull);
    h = hashBytes(a.normals.data(), a.normals.size() * 8, h);
    h = hashBytes(a.uvs.data(), a.uvs.size() * 8, h);
    return hashBytes(a.tris.data(), a.tris.size() * 4, h);
}

bool objKey(const char* objPath, uint64_t& key, std::string& path)
{
    const std::string d = cacheDir();
    struct stat st;
    if (d.empty() || stat(objPath, &st) != 0) return false;
    char real[4096];
    if (!realpath(objPath, real)) return false;
    key = hashBytes(real, strlen(real), 0x4F424A31ull);
    const long long meta[2] = {(long long)st.st_size, (long long)st.st_mtime};
    key = hashBytes(meta, sizeof meta, key);
    char b[64];
    snprintf(b, sizeof b, "/obj_%016llx.bin", (unsigned long long)key);
    path = d + b;
    return true;
}
}  // namespace

bool loadObjCache(const char* objPath, ObjArrays& out)
{
    uint64_t key;
    std::string path;
    if (!objKey(objPath, key, path)) return false;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    ObjHeader h;
    bool ok = fread(&h, sizeof h, 1, f) == 1 && !memcmp(h.magic, "HXROBJ\0", 8) && h.version == OBJ_VERSION && h.key == key &&
              h.n_vertices < (1ull << 31) && h.n_normals < (1ull << 31) && h.n_uvs < (1ull << 31) && h.n_tris < (1ull << 31);
    if (ok) {
        out.vertices.resize(h.n_vertices * 3);
        out.normals.resize(h.n_normals * 3);
        out.uvs.resize(h.n_uvs * 3);
        out.tris.resize(h.n_tris * 9);
        auto rd = [&](void* p, size_t bytes) { return bytes == 0 || fread(p, 1, bytes, f) == bytes; };
        ok = rd(out.vertices.data(), out.vertices.size() * 8) && rd(out.normals.data(), out.normals.size() * 8) && rd(out.uvs.data(), out.uvs.size() * 8) &&
             rd(out.tris.data(), out.tris.size() * 4);
        ok = ok && objContentHash(out) == h.content;
    }
    fclose(f);
    if (!ok) out = ObjArrays();
    return ok;
}

void storeObjCache(const char* objPath, const ObjArrays& a)
{
    uint64_t key;
    std::string path;
    if (!objKey(objPath, key, path)) return;
    ObjHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "HXROBJ\0", 8);
    h.version = OBJ_VERSION;
    h.key = key;
    h.n_vertices = a.vertices.size() / 3;
    h.n_normals = a.normals.size() / 3;
    h.n_uvs = a.uvs.size() / 3;
    h.n_tris = a.tris.size() / 9;
    h.content = objContentHash(a);
    writeAtomically(path, {{&h, sizeof h}, {a.vertices.data(), a.vertices.size() * 8}, {a.normals.data(), a.normals.size() * 8},
                           {a.uvs.data(), a.uvs.size() * 8}, {a.tris.data(), a.tris.size() * 4}});
}

}  // namespace host
}  // namespace hxr
