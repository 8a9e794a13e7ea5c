// kdtree.h — host-side SAH KD-tree build for triangle meshes, flattened for the GPU.
// Replaces the reference's build (Mesh::buildKD, src/mesh.cpp:95-122: median split, axis =
// depth % 3, leaf < 20 triangles or depth > 64, exact triangle-box overlap) with a surface-area-
// heuristic build; the tree only has to preserve the CLOSEST triangle hit (see csrc/device/isect.h).
#pragma once
#include <cstdint>
#include <vector>
#include "../device/scene_dev.h"

namespace hxr {
namespace host {

struct KdTree {
    std::vector<KdNode> nodes;        // node 0 = root; children stored explicitly (DFS pre-order)
    std::vector<uint32_t> leafTris;   // triangle indices, leaf after leaf, ascending inside a leaf
    uint32_t maxDepth = 0;
    uint64_t leaves = 0;
    double buildMs = 0;
};

struct KdBuildParams {
    float traversalCost = 1.0f;
    float intersectCost = 2.0f;
    float emptyBonus = 0.2f;
    int maxLeafSize = 4;
    int maxDepth = -1;      // -1: 8 + 1.3 log2(N), capped so the device stack (HXR_KD_STACK) cannot overflow
    int binnedAbove = 192;  // nodes with more triangles than this use 32-bin SAH, smaller ones an exact sweep
    int threads = 0;        // 0: hardware concurrency
};

void buildKdTree(const hxr_mesh& mesh, const KdBuildParams& params, KdTree& out);

}  // namespace host
}  // namespace hxr
