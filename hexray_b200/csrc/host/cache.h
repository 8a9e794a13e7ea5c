// cache.h — binary caches of the two expensive host steps before the first ray:
//   * a parsed OBJ (the reference tokenises text with sscanf per token on every start, src/mesh.cpp:301-358: 20 s for the
//     10 M-triangle file): vertices / normals / uvs / triangle indices as one binary blob, keyed by the file's path, size
//     and modification time - the second load is O(read);
//   * a built KD-tree (the reference rebuilds its tree on every start, src/mesh.cpp:95-122), keyed by a hash of the mesh
//     content and the build parameters. The first process to ask builds it and publishes the file; processes that ask
//     while it is being built (the other ranks of a torchrun job: one process per GPU, the same scene) WAIT for that file
//     instead of building the same tree again - the tree is built once per box, not once per GPU.
// Files live in $HXR_CACHE_DIR (default /tmp/hexray_b200_cache); HXR_CACHE=0 turns both caches off. Every file is written to a
// temporary name and renamed into place, and carries a magic, a version, its key, its element counts and a hash of its
// payload: a stale, foreign, truncated or damaged file is ignored and rebuilt. The lock file of a tree being built names its
// builder's pid; a lock whose process is gone (a run killed mid-build) is taken over, not waited for.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "kdtree.h"

namespace hxr {
namespace host {

bool cacheEnabled();
std::string cacheDir();  // created on first use; "" when the cache is off or the directory cannot be made
uint64_t hashBytes(const void* data, size_t n, uint64_t seed);

// ---- KD-tree of a mesh
uint64_t meshContentKey(const hxr_mesh& mesh, const KdBuildParams& params);
bool loadKdTree(uint64_t key, const hxr_mesh& mesh, KdTree& out);   // false: not cached (or unusable)
void storeKdTree(uint64_t key, const hxr_mesh& mesh, const KdTree& kd);
// buildKdTree through the cache: load it, or wait for another process that is building it, or build and publish it.
// how (optional): "cache", "waited", "built"
void cachedKdTree(const hxr_mesh& mesh, const KdBuildParams& params, KdTree& out, const char** how = nullptr);

// ---- parsed OBJ
struct ObjArrays {
    std::vector<double> vertices, normals, uvs;  // 3 per entry, slot 0 = the sentinel
    std::vector<int32_t> tris;                   // 9 per triangle: v[3], n[3], t[3]
};
bool loadObjCache(const char* objPath, ObjArrays& out);
void storeObjCache(const char* objPath, const ObjArrays& a);

}  // namespace host
}  // namespace hxr
