// scene_dev.h — the scene as the kernels see it: flat tables in HBM.
//
// Layout notes (B200): the per-scene tables (nodes, geometries, shaders, lights) are a few
// hundred bytes and stay L1/L2 resident; the per-mesh arrays are the HBM traffic:
//   KdNode   16 B  one 128-bit load per traversal step
//   TriTest  96 B  the four vectors the reference's triangle test reads (A, AB, AC, AB^AC),
//                  in double so the test is the reference's arithmetic (src/mesh.cpp:178-196)
//   TriAttr  96 B  read once per ray for the winning triangle (normals/uv indices, dNdx/dNdy)
#pragma once
#include "hd.h"
#include "../../../include/hxr.h"

#define HXR_KD_STACK 64 /* device traversal stack entries; the host build caps the tree depth below it */

namespace hxr {

// KD-tree node. inner: kind = axis (0..2), a = left child, b = right child.
//               leaf : kind = 3,           a = first entry in leaf_tris, b = triangle count.
struct alignas(16) KdNode {
    float split;
    uint32_t kind;
    uint32_t a, b;
};

struct alignas(16) TriTest {
    double A[3], AB[3], AC[3], N[3];
};

struct TriAttr {
    int32_t n[3], t[3];
    double gnormal[3], dNdx[3], dNdy[3];
};

struct DMesh {
    const KdNode* nodes;
    const uint32_t* leaf_tris;
    const TriTest* tri_test;
    const TriAttr* tri_attr;
    const double* normals;
    const double* uvs;
    double bbmin[3], bbmax[3];
    int32_t faceted, backface;
    int32_t n_tris, pad;
};

struct DHeightfield {
    const float* heights;
    const float* max_h;
    const double* normals;
    const float* high_map;
    double bbmin[3], bbmax[3];
    int32_t W, H, use_opt, max_k;
};

struct DImage {
    const float* rgb;
    int32_t w, h;
};

struct DScene {
    const hxr_node* nodes;
    const hxr_geometry* geoms;
    const DMesh* meshes;
    const DHeightfield* hfs;
    const hxr_shader* shaders;
    const hxr_layer* layers;
    const hxr_texture* textures;
    const DImage* images;
    const hxr_light* lights;
    // per node: slot of its traversal results if its geometry is a "big" mesh (walked by the persistent
    // traversal kernel), -1 otherwise (analytic primitives, CSG, heightfields, meshes <= HXR_SMALL_MESH)
    const int32_t* node_slot;
    int32_t n_big;
    int32_t n_nodes, n_lights;
    int32_t has_env;
    int32_t env_images[6];
    hxr_settings settings;
    hxr_camera cam;
};

// optional traversal counters (HXR_RENDER_COUNT_TRAVERSAL)
struct TravCounters {
    unsigned long long kd_inner, kd_leaves, tri_tests, mesh_queries;
};

}  // namespace hxr
