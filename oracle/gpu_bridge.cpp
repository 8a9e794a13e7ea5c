// oracle/gpu_bridge.cpp — TEST INFRASTRUCTURE, and at the same time the reference-side binding INTEGRATION.md describes:
// the one translation unit a maintainer of anrieff/hexray would add (as src/gpu_bridge.cpp) to render on the GPU.
//
// It is compiled HERE against the UNMODIFIED reference headers and linked with the reference's own objects (parser, scene
// graph, beginRender/beginFrame code: oracle/Makefile) and with libhexray_b200.so. After the reference's own
//     scene.parseScene(); scene.beginRender(); scene.beginFrame();            (src/main.cpp:506-511, 536)
// it walks the LIVE scene graph (`Scene scene`, src/scene.h:271-294), fills the POD tables of include/hxr.h, calls
// hxr_upload_scene / hxr_set_camera / hxr_render and writes the frame into the reference's `vfb` (src/main.cpp:50), from
// where the reference's own display / screenshot code carries on. It proves that the POD tables carry everything the
// reference's classes hold: tests/test_bridge.py renders every bundled scene through it and compares with the frame of
// this repo's own front-end (must be identical) and with the reference's CPU render (within the parity tolerance).
//
// The reference keeps the members this needs private (Camera::m_topLeft, Mesh::triangles, Layered::m_layers, ...). A
// maintainer would add `friend bool renderOnGPU(hxr_stats*, unsigned long long);` or accessors; this file is instead compiled
// with GCC's -fno-access-control (oracle/Makefile), so that not one reference source line has to change.
#include <array>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "scene.h"
#include "camera.h"
#include "geometry.h"
#include "mesh.h"
#include "heightfield.h"
#include "shading.h"
#include "lights.h"
#include "environment.h"
#include "node.h"
#include "bitmap.h"
#include "sdl.h"

#include "../include/hxr.h"

extern Color vfb[VFB_MAX_SIZE][VFB_MAX_SIZE];  // src/main.cpp:50

namespace {

void put(double* d, const Vector& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
void put(float* d, const Color& c) { d[0] = c.r; d[1] = c.g; d[2] = c.b; }
void put(hxr_transform& o, const Transform& T)  // src/matrix.h:72-99
{
    put(o.offset, T.offset);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            o.m[i * 3 + j] = T.m.m[i][j];
            o.inv[i * 3 + j] = T.invM.m[i][j];
            o.inv_t[i * 3 + j] = T.transposedInverse.m[i][j];
        }
}

struct Tables {
    std::vector<hxr_geometry> geoms;
    std::vector<hxr_mesh> meshes;
    std::vector<hxr_heightfield> hfs;
    std::vector<hxr_node> nodes;
    std::vector<hxr_shader> shaders;
    std::vector<hxr_layer> layers;
    std::vector<hxr_texture> textures;
    std::vector<hxr_image> images;
    std::vector<hxr_light> lights;
    std::vector<std::vector<hxr_triangle>> triStore;
    std::vector<std::vector<float>> floatStore;
    int addImage(const Bitmap& b)  // src/bitmap.h: float RGB, row-major, top-down; Color = 3 floats
    {
        hxr_image im;
        im.width = b.isOK() ? b.m_width : 0;
        im.height = b.isOK() ? b.m_height : 0;
        im.rgb = im.width && im.height ? reinterpret_cast<const float*>(b.m_data.data()) : nullptr;
        images.push_back(im);
        return (int)images.size() - 1;
    }
};

template <class P> int indexOf(const std::vector<P*>& list, const void* p)
{
    for (size_t i = 0; i < list.size(); i++)
        if ((const void*)list[i] == p) return (int)i;
    return -1;
}

bool fillScene(Tables& t, hxr_scene& s, std::string& err)
{
    static_assert(sizeof(Vector) == 3 * sizeof(double) && sizeof(Color) == 3 * sizeof(float), "Vector / Color are plain triples");
    memset(&s, 0, sizeof s);
    // ---- geometries (src/geometry.h:57-136, src/mesh.h:72-124, src/heightfield.h:30-52), pointers -> indices in list order
    for (Geometry* g : scene.geometries) {
        hxr_geometry o;
        memset(&o, 0, sizeof o);
        if (auto* p = dynamic_cast<Plane*>(g)) {
            o.type = HXR_GEOM_PLANE;
            o.p[0] = p->y;
            o.p[1] = p->limit;
        } else if (auto* sp = dynamic_cast<Sphere*>(g)) {
            o.type = HXR_GEOM_SPHERE;
            put(o.p, sp->O);
            o.p[3] = sp->R;
            o.p[4] = sp->uvscaling;
        } else if (auto* c = dynamic_cast<Cube*>(g)) {
            o.type = HXR_GEOM_CUBE;
            put(o.p, c->O);
            o.p[3] = c->m_halfSide;  // Cube::beginFrame
        } else if (auto* csg = dynamic_cast<CSGBase*>(g)) {
            o.type = HXR_GEOM_CSG;
            o.a = dynamic_cast<CSGUnion*>(g) ? HXR_CSG_UNION : (dynamic_cast<CSGInter*>(g) ? HXR_CSG_INTER : HXR_CSG_DIFF);
            o.b = indexOf(scene.geometries, csg->left);
            o.c = indexOf(scene.geometries, csg->right);
            if (o.b < 0 || o.c < 0) { err = "CSG child is not a scene geometry"; return false; }
        } else if (auto* m = dynamic_cast<Mesh*>(g)) {
            // the arrays as Mesh::beginRender left them (src/mesh.cpp:49-87): slot 0 of vertices / normals / uvs is the OBJ
            // sentinel; Triangle (src/bbox.h:36-43) after prepareTriangles, field by field
            hxr_mesh hm;
            memset(&hm, 0, sizeof hm);
            hm.n_vertices = (int)m->vertices.size();
            hm.n_normals = (int)m->normals.size();
            hm.n_uvs = (int)m->uvs.size();
            hm.n_triangles = (int)m->triangles.size();
            hm.vertices = reinterpret_cast<const double*>(m->vertices.data());
            hm.normals = reinterpret_cast<const double*>(m->normals.data());
            hm.uvs = reinterpret_cast<const double*>(m->uvs.data());
            t.triStore.emplace_back(m->triangles.size());
            std::vector<hxr_triangle>& tr = t.triStore.back();
            for (size_t i = 0; i < tr.size(); i++) {
                const Triangle& T = m->triangles[i];
                memset(&tr[i], 0, sizeof tr[i]);
                for (int k = 0; k < 3; k++) { tr[i].v[k] = T.v[k]; tr[i].n[k] = T.n[k]; tr[i].t[k] = T.t[k]; }
                put(tr[i].gnormal, T.gnormal);
                put(tr[i].ab, T.AB);
                put(tr[i].ac, T.AC);
                put(tr[i].ab_cross_ac, T.ABcrossAC);
                put(tr[i].dndx, T.dNdx);
                put(tr[i].dndy, T.dNdy);
            }
            hm.triangles = tr.data();
            hm.faceted = m->faceted;
            hm.backface_culling = m->backfaceCulling;
            put(hm.bbox_min, m->bbox.vmin);
            put(hm.bbox_max, m->bbox.vmax);
            o.type = HXR_GEOM_MESH;
            o.a = (int)t.meshes.size();
            t.meshes.push_back(hm);
        } else if (auto* h = dynamic_cast<Heightfield*>(g)) {
            hxr_heightfield hh;
            memset(&hh, 0, sizeof hh);
            hh.width = h->W;
            hh.height = h->H;
            hh.use_optimization = h->useOptimization;
            hh.max_k = h->maxK;
            hh.heights = h->heights.data();
            hh.max_h = h->maxH.data();
            hh.normals = reinterpret_cast<const double*>(h->normals.data());
            hh.high_map = h->highMap.empty() ? nullptr : reinterpret_cast<const float*>(h->highMap.data());
            put(hh.bbox_min, h->bbox.vmin);
            put(hh.bbox_max, h->bbox.vmax);
            o.type = HXR_GEOM_HEIGHTFIELD;
            o.a = (int)t.hfs.size();
            t.hfs.push_back(hh);
        } else {
            err = "unknown Geometry subclass";
            return false;
        }
        t.geoms.push_back(o);
    }
    // ---- textures (src/shading.h:58-227)
    for (Texture* x : scene.textures) {
        hxr_texture o;
        memset(&o, 0, sizeof o);
        o.image = -1;
        if (auto* c = dynamic_cast<CheckerTexture*>(x)) {
            o.type = HXR_TEX_CHECKER;
            put(o.color1, c->color1);
            put(o.color2, c->color2);
            o.scaling = c->scaling;
        } else if (auto* b = dynamic_cast<BitmapTexture*>(x)) {
            o.type = HXR_TEX_BITMAP;
            o.image = t.addImage(b->m_bitmap);
            o.scaling = b->scaling;
        } else if (auto* f = dynamic_cast<Fresnel*>(x)) {
            o.type = HXR_TEX_FRESNEL;
            o.ior = f->ior;
        } else if (auto* bt = dynamic_cast<BumpTexture*>(x)) {
            o.type = HXR_TEX_BUMP;  // already differentiated by BumpTexture::beginRender
            o.image = t.addImage(bt->bitmap);
            o.strength = bt->strength;
            o.scaling = bt->scaling;
        } else if (auto* bs = dynamic_cast<Bumps*>(x)) {
            o.type = HXR_TEX_BUMPS;
            o.strength = bs->strength;
        } else {
            err = "unknown Texture subclass";
            return false;
        }
        t.textures.push_back(o);
    }
    // ---- shaders (src/shading.h:97-214); Phong derives from Lambert
    for (Shader* x : scene.shaders) {
        hxr_shader o;
        memset(&o, 0, sizeof o);
        o.tex = -1;
        if (auto* ph = dynamic_cast<Phong*>(x)) {
            o.type = HXR_SHADER_PHONG;
            put(o.color, ph->diffuse);
            o.tex = indexOf(scene.textures, ph->diffuseTex);
            put(o.color2, ph->specular);
            o.f0 = ph->exponent;
        } else if (auto* la = dynamic_cast<Lambert*>(x)) {
            o.type = HXR_SHADER_LAMBERT;
            put(o.color, la->diffuse);
            o.tex = indexOf(scene.textures, la->diffuseTex);
        } else if (auto* rl = dynamic_cast<Reflection*>(x)) {
            o.type = HXR_SHADER_REFLECTION;
            put(o.color, rl->reflColor);
            o.f0 = rl->glossiness;
            o.i0 = rl->numSamples;
        } else if (auto* rr = dynamic_cast<Refraction*>(x)) {
            o.type = HXR_SHADER_REFRACTION;
            put(o.color, rr->refrColor);
            o.ior = rr->ior;
        } else if (auto* ly = dynamic_cast<Layered*>(x)) {
            o.type = HXR_SHADER_LAYERED;
            o.first_layer = (int)t.layers.size();
            o.n_layers = (int)ly->m_layers.size();
            for (const auto& l : ly->m_layers) {
                hxr_layer hl;
                hl.shader = indexOf(scene.shaders, l.shader);
                hl.tex = indexOf(scene.textures, l.blendTex);
                put(hl.blend, l.blend);
                if (hl.shader < 0) { err = "Layered layer is not a scene shader"; return false; }
                t.layers.push_back(hl);
            }
        } else if (auto* co = dynamic_cast<Const*>(x)) {
            o.type = HXR_SHADER_CONST;
            put(o.color, co->color);
        } else {
            err = "unknown Shader subclass";
            return false;
        }
        t.shaders.push_back(o);
    }
    // ---- lights (src/lights.h:79-118) after beginFrame
    for (Light* x : scene.lights) {
        hxr_light o;
        memset(&o, 0, sizeof o);
        o.power = x->power;
        put(o.color, x->color);
        if (auto* pl = dynamic_cast<PointLight*>(x)) {
            o.type = HXR_LIGHT_POINT;
            o.xsubd = o.ysubd = 1;
            o.scale_factor = 1.0f;
            put(o.pos, pl->pos);
            put(o.T, Transform());
        } else if (auto* rl = dynamic_cast<RectLight*>(x)) {
            o.type = HXR_LIGHT_RECT;
            o.xsubd = rl->xSubd;
            o.ysubd = rl->ySubd;
            o.scale_factor = rl->scaleFactor;  // RectLight::beginFrame: 1 / area (src/lights.cpp:75-88)
            o.area = rl->m_area;
            put(o.pos, rl->T.offset);
            put(o.T, rl->T);
        } else {
            err = "unknown Light subclass";
            return false;
        }
        t.lights.push_back(o);
    }
    // ---- nodes: scene.nodes only (the shader-less ones have been moved to superNodes, src/scene.cpp:560-565)
    for (Node* n : scene.nodes) {
        hxr_node o;
        memset(&o, 0, sizeof o);
        o.geom = indexOf(scene.geometries, n->geom);
        o.shader = indexOf(scene.shaders, n->shader);
        o.bump_tex = indexOf(scene.textures, n->bump);
        put(o.T, n->T);
        if (o.geom < 0 || o.shader < 0) { err = "Node without geometry or shader"; return false; }
        t.nodes.push_back(o);
    }
    // ---- environment (src/environment.h:47-70): NEGX NEGY NEGZ POSX POSY POSZ
    if (auto* env = dynamic_cast<CubemapEnvironment*>(scene.environment)) {
        if (env->loaded) {
            s.has_environment = 1;
            for (int i = 0; i < 6; i++) s.env_images[i] = t.addImage(env->m_sides[i]);
        }
    }
    s.abi_version = HXR_ABI_VERSION;
    s.n_nodes = (int)t.nodes.size(); s.nodes = t.nodes.data();
    s.n_geometries = (int)t.geoms.size(); s.geometries = t.geoms.data();
    s.n_meshes = (int)t.meshes.size(); s.meshes = t.meshes.data();
    s.n_heightfields = (int)t.hfs.size(); s.heightfields = t.hfs.data();
    s.n_shaders = (int)t.shaders.size(); s.shaders = t.shaders.data();
    s.n_layers = (int)t.layers.size(); s.layers = t.layers.data();
    s.n_textures = (int)t.textures.size(); s.textures = t.textures.data();
    s.n_images = (int)t.images.size(); s.images = t.images.data();
    s.n_lights = (int)t.lights.size(); s.lights = t.lights.data();
    const GlobalSettings& gs = scene.settings;  // src/scene.h:244-269
    s.settings.frame_width = gs.frameWidth;
    s.settings.frame_height = gs.frameHeight;
    s.settings.max_trace_depth = gs.maxTraceDepth;
    s.settings.want_aa = gs.wantAA;
    s.settings.gi = gs.gi;
    s.settings.num_paths = gs.numPaths;
    put(s.settings.ambient, gs.ambientLight);
    put(s.settings.background, gs.backgroundColor);
    return true;
}

void fillCamera(hxr_camera& c)  // what Camera::beginFrame computed (src/camera.cpp:30-63)
{
    const Camera& cam = *scene.camera;
    memset(&c, 0, sizeof c);
    put(c.pos, cam.pos);
    put(c.top_left, cam.m_topLeft);
    put(c.top_right, cam.m_topRight);
    put(c.bottom_left, cam.m_bottomLeft);
    put(c.up, cam.m_upDir);
    put(c.right, cam.m_rightDir);
    put(c.front, cam.m_frontDir);
    c.aperture_size = cam.m_apertureSize;
    c.focal_plane_dist = cam.focalPlaneDist;
    c.stereo_separation = cam.stereoSeparation;
    c.dof = cam.dof;
    c.auto_focus = cam.autoFocus;
    c.num_samples = cam.numSamples;
}

hxr_ctx* g_ctx = nullptr;

}  // namespace

// Replaces render() in renderStatic (src/main.cpp:506-511): scene.beginRender() / beginFrame() have run.
// seed: Philox seed of stochastic frames. Returns false (and prints the library's message) on failure.
bool renderOnGPU(hxr_stats* stats, unsigned long long seed)
{
    if (!g_ctx) {
        hxr_config cfg;
        memset(&cfg, 0, sizeof cfg);
        if (hxr_create(&cfg, &g_ctx) != HXR_OK) {
            fprintf(stderr, "renderOnGPU: %s\n", hxr_last_error(nullptr));
            return false;
        }
    }
    Tables t;
    hxr_scene s;
    std::string err;
    if (!fillScene(t, s, err)) {
        fprintf(stderr, "renderOnGPU: %s\n", err.c_str());
        return false;
    }
    if (hxr_upload_scene(g_ctx, &s) != HXR_OK) {
        fprintf(stderr, "renderOnGPU: %s\n", hxr_last_error(g_ctx));
        return false;
    }
    hxr_camera cam;
    fillCamera(cam);
    hxr_set_camera(g_ctx, &cam);
    hxr_render_params p;
    memset(&p, 0, sizeof p);  // everything = the scene's own settings (frame size, AA, depth, spp)
    p.want_aa = -1;
    p.max_depth = -1;
    p.seed = seed;
    const int W = frameWidth(), H = frameHeight();
    std::vector<float> rgb((size_t)W * H * 3);
    hxr_stats st;
    if (hxr_render(g_ctx, &p, rgb.data(), &st) != HXR_OK) {
        fprintf(stderr, "renderOnGPU: %s\n", hxr_last_error(g_ctx));
        return false;
    }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const float* c = &rgb[((size_t)y * W + x) * 3];
            vfb[y][x] = Color(c[0], c[1], c[2]);
        }
    if (stats) *stats = st;
    return true;
}

void shutdownGPU()
{
    if (g_ctx) hxr_destroy(g_ctx);
    g_ctx = nullptr;
}
