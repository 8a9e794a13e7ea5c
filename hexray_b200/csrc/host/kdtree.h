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
    std::vector<KdNode> nodes;        // binary tree (build intermediate): node 0 = root, DFS pre-order
    std::vector<KdBlock> blocks;      // what the device walks: two tree levels per 32-byte block, block 0 = root
    std::vector<uint32_t> leafTris;   // per leaf: [count, triangle indices ascending]; a leaf reference points at its count
    uint32_t maxDepth = 0;
    uint64_t leaves = 0;
    double buildMs = 0;
    double deviceMs = 0;  // device-built trees: the part of buildMs spent on the GPU passes (the rest packs the blocks on the host)
};

struct KdBuildParams {
    float traversalCost = 1.0f;
    float intersectCost = 1.0f;  // a triangle costs the walk about one tree level: leaves are filtered 32 pairs at a time (k_walk)
    float emptyBonus = 0.2f;
    int maxLeafSize = 16;
    int maxDepth = -1;      // -1: 8 + 1.3 log2(N), capped at HXR_KD_MAX_DEPTH so the device stack cannot overflow
    int binnedAbove = 192;  // nodes with more triangles than this use 32-bin SAH, smaller ones an exact sweep
    int threads = 0;        // 0: hardware concurrency
};

void buildKdTree(const hxr_mesh& mesh, const KdBuildParams& params, KdTree& out);

// The two ends of a build somebody else does (the device build, csrc/kdbuild.cpp): the parameters with the environment knobs
// and the depth limit resolved, the root box (float-rounded outward), and the packing of a finished binary tree
// (out.nodes / out.leafTris raw lists / maxDepth / leaves) into the blocks and leaf lists the device walks.
void resolveKdParams(const hxr_mesh& mesh, KdBuildParams& params);
void kdRootBox(const hxr_mesh& mesh, double mn[3], double mx[3]);
void packKdTree(KdTree& tree, const hxr_mesh& mesh);

}  // namespace host
}  // namespace hxr
