// scene_dev.h — the scene as the kernels see it: flat tables in HBM.
//
// Layout notes (B200): the per-scene tables (nodes, geometries, shaders, lights) are a few
// hundred bytes and stay L1/L2 resident; the per-mesh arrays are the HBM traffic:
//   KdBlock  32 B  one L2 sector = TWO levels of the KD-tree (a node and both its children): one fetch
//                  (two 128-bit loads) decides up to four grandchildren; leaves are not nodes at all but
//                  tagged references into leaf_tris
//   TriF32   48 B  float copy of TriTest: what the walk's conservative triangle filter reads (3 x 128-bit loads)
//   TriPacked 32 B ONE sector: A in float, AB and AC as block floating point (one exponent byte + three signed 24-bit
//                  mantissas per vector); the filter rebuilds AB x AC itself. The walk saturates the GPU's RANDOM-access
//                  DRAM rate (tools/microbench.cu: ~31 G sectors/s whatever the record size), so sectors per triangle
//                  are what the walk pays for
//   TriTest  96 B  the four vectors the reference's triangle test reads (A, AB, AC, AB^AC),
//                  in double so the test is the reference's arithmetic (src/mesh.cpp:178-196)
//   TriAttr 96 B + TriAttrUv 96 B  read once per ray for the winning triangle: normals / uvs and dNdx, dNdy inline
#pragma once
#include "hd.h"
#include "../../../include/hxr.h"

/* hxr_node::pad in the DEVICE copy of the node table (set by hxr_upload_scene): bit 0 = m, inv and inv_t are exactly the identity
 * (a node without rotate / scale): v * I == v bit for bit, so the three mat-vecs per transform are skipped */
#define HXR_NODE_IDENT 1

#define HXR_KD_MAX_DEPTH 60 /* the host build caps the tree depth here */
#define HXR_KD_STACK 96     /* traversal stack entries: a block step (two levels) pushes at most three */

namespace hxr {

// Host-side binary KD node (build intermediate, never uploaded).
//   inner: kind = axis (0..2), a = left child, b = right child.
//   leaf : kind = 3,           a = first entry in leaf_tris, b = triangle count.
struct alignas(16) KdNode {
    float split;
    uint32_t kind;
    uint32_t a, b;
};

// Device KD-tree: 32-byte blocks holding a node (sub 0) and its two children (sub 1 = left, sub 2 = right).
//   meta bits 0-1 / 2-3 / 4-5 = split axis of sub 0 / 1 / 2; axis 3 marks a child that is a LEAF.
//   ref[0], ref[1] = children of sub 1 (or ref[0] = its leaf reference), ref[2], ref[3] likewise for sub 2.
//   A reference is either a block index (bit 31 clear) or HXR_KD_LEAF | position of the leaf's list in leaf_tris
//   ([count, triangle indices...]); HXR_KD_EMPTY is a leaf without triangles.
//   "left" holds coordinates <= split, "right" >= split.
struct alignas(32) KdBlock {
    float split[3];
    uint32_t meta;
    uint32_t ref[4];
};
#define HXR_KD_LEAF 0x80000000u
#define HXR_KD_EMPTY 0xFFFFFFFFu

struct alignas(16) TriTest {
    double A[3], AB[3], AC[3], N[3];
};

struct alignas(16) TriF32 {  // 48 B: the same four vectors rounded to float, for the walk's triangle filter (isect.h: tri_filter)
    float A[3], AB[3], AC[3], N[3];
};

// 32 B: A rounded to float; AB, AC as v_i = m_i * 2^e with one exponent per vector (|v_i - m_i 2^e| <= 2^(e-1)).
//   w[0] = (eAB + 126) | (eAC + 126) << 8 | m0 low 16 << 16;  w[1] = m0 high 8 | m1 << 8;  w[2] = m2 | m3 low 8 << 24;
//   w[3] = m3 high 16 | m4 low 16 << 16;  w[4] = m4 high 8 | m5 << 8      (m0..2 = AB, m3..5 = AC, two's complement)
struct alignas(32) TriPacked {
    float A[3];
    uint32_t w[5];
};

// The winning triangle's shading data, read once per ray by finalize. Two 96-byte records (3 sectors each): every hit needs
// the normals; the texture coordinates and dNdx/dNdy only matter to textured or bump-mapped nodes (DScene::node_lean), so an
// untextured mesh costs 3 random sectors per hit instead of 6 (random sectors are what a gather pays for, see TriPacked)
struct alignas(32) TriAttr {
    double gnormal[3];
    double nrm[3][3];  // the three vertex normals
};
struct alignas(32) TriAttrUv {
    double uv[3][2];  // the three texture coordinates
    double dNdx[3], dNdy[3];
};

struct DMesh {
    const KdBlock* blocks;
    const uint32_t* leaf_tris;
    const TriTest* tri_test;
    const TriF32* tri_f32;
    const TriPacked* tri_pk;  // null when an edge of the mesh does not fit the packed form (the walk then reads tri_f32)
    const TriAttr* tri_attr;
    const TriAttrUv* tri_attr_uv;
    double bbmin[3], bbmax[3];
    int32_t faceted, backface;
    int32_t n_tris;
    float abs_max;  // largest |coordinate| of the mesh box (error bound of the float filter)
    int32_t pad;
    float fbmin[3], fbmax[3];  // the box inflated by 1e-6 (the slab test's tolerance, src/bbox.h) as floats rounded outward: the walk's entry test
    int32_t brute;  // test all triangles in index order (tiny meshes, or HXR_CFG_BRUTE_FORCE_MESHES) instead of walking the tree; 3: and skip the box gate
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
    // this struct's copy in device memory (static tables only: made by hxr_upload_scene, it does not follow later camera or
    // settings changes). For the rare non-inlined device functions: they take the scene by pointer, because a reference to
    // the kernel parameter would make the compiler keep a copy of all of it in every thread's local memory.
    const DScene* self;
    const hxr_node* nodes;
    const hxr_geometry* geoms;
    const DMesh* meshes;
    const DHeightfield* hfs;
    const hxr_shader* shaders;
    const hxr_layer* layers;
    const hxr_texture* textures;
    const DImage* images;
    const hxr_light* lights;
    // per light: conservative float world box of what a ray can hit of it (rect lights; an empty box for point lights):
    // raycast's light loop skips the double transform of a light whose box the ray certainly misses
    const float* light_box;
    // per node: slot of its traversal results if its geometry is a "big" mesh (walked by the persistent
    // traversal kernel), -1 otherwise (analytic primitives, CSG, heightfields, meshes <= HXR_SMALL_MESH)
    const int32_t* node_slot;
    // walked nodes in scene order: big_nodes[slot] = node index; big_box[6 * slot] = float copy of that node's world box,
    // rounded outward (the walk kernel skips a mesh whose box the ray certainly misses before paying for the double transform)
    const int32_t* big_nodes;
    const float* big_box;
    // inline nodes in scene order: inline_nodes[k] = node index; inline_box[6 * k] = conservative world-space box of its geometry
    // as floats rounded outward (min xyz, max xyz; +-inf when unbounded or unknown): the node loops skip a node whose box
    // the ray misses before paying for the object-space transform and intersector
    const int32_t* inline_nodes;
    const float* inline_box;
    // per node: 1 if nothing downstream reads u, v, dNdx, dNdy of a hit on it (no texture anywhere in its shader, no bump map):
    // finalize then skips the TriAttrUv gather and leaves those fields zero. full_attr = 1 overrides (the trace_closest hook).
    const int32_t* node_lean;
    int32_t full_attr;
    int32_t use_node_box;
    int32_t walk_packed;  // every walked mesh has tri_pk: the walk filters 32-byte packed triangles
    int32_t n_big;
    int32_t n_inline;  // nodes handled inline (n_nodes - n_big)
    int32_t simple_inline;  // every inline node is a plane, sphere, cube or brute-force mesh (selects the lean kernel variants)
    int32_t n_nodes, n_lights;
    int32_t has_env;
    int32_t env_images[6];
    hxr_settings settings;
    hxr_camera cam;
};

// optional traversal counters (HXR_RENDER_COUNT_TRAVERSAL)
struct TravCounters {
    unsigned long long kd_inner, kd_leaves, tri_tests, mesh_queries;
    unsigned long long reserved;
};

}  // namespace hxr
