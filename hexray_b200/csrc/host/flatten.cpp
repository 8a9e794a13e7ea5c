// flatten.cpp — scene graph -> POD tables of include/hxr.h (the layer SURVEY.md §1 adds between
// the parser and the C ABI). Runs the reference's pre-render callbacks first, in the reference's
// visiting order (Scene::beginRender / beginFrame, src/scene.cpp:741-762).
#include <cstring>
#include "scene.h"

namespace hxr {
namespace host {

void toPodTransform(const Transform& T, hxr_transform& o)
{
    o.offset[0] = T.offset.x; o.offset[1] = T.offset.y; o.offset[2] = T.offset.z;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            o.m[i * 3 + j] = T.m.m[i][j];
            o.inv[i * 3 + j] = T.invM.m[i][j];
            o.inv_t[i * 3 + j] = T.transposedInverse.m[i][j];
        }
}

static void putc3(float* d, const Color3& c) { d[0] = c.r; d[1] = c.g; d[2] = c.b; }
static void putv3(double* d, const Vec3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

int FlatScene::addImage(const Bitmap& bmp)
{
    hxr_image im;
    im.width = bmp.isOK() ? bmp.getWidth() : 0;
    im.height = bmp.isOK() ? bmp.getHeight() : 0;
    imageStore.emplace_back();
    std::vector<float>& st = imageStore.back();
    st.resize((size_t)im.width * im.height * 3);
    const auto& px = bmp.data();
    for (size_t i = 0; i < px.size() && i * 3 + 2 < st.size() + 0; i++) {
        st[i * 3] = px[i].r; st[i * 3 + 1] = px[i].g; st[i * 3 + 2] = px[i].b;
    }
    im.rgb = st.empty() ? nullptr : st.data();
    images.push_back(im);
    return (int)images.size() - 1;
}

const double* FlatScene::keepDoubles(const std::vector<Vec3>& v)
{
    doubleStore.emplace_back(v.size() * 3);
    std::vector<double>& d = doubleStore.back();
    for (size_t i = 0; i < v.size(); i++) { d[i * 3] = v[i].x; d[i * 3 + 1] = v[i].y; d[i * 3 + 2] = v[i].z; }
    return d.data();
}

void FlatScene::finalize()
{
    pod.abi_version = HXR_ABI_VERSION;
    pod.n_nodes = (int)nodes.size(); pod.nodes = nodes.data();
    pod.n_geometries = (int)geometries.size(); pod.geometries = geometries.data();
    pod.n_meshes = (int)meshes.size(); pod.meshes = meshes.data();
    pod.n_heightfields = (int)heightfields.size(); pod.heightfields = heightfields.data();
    pod.n_shaders = (int)shaders.size(); pod.shaders = shaders.data();
    pod.n_layers = (int)layers.size(); pod.layers = layers.data();
    pod.n_textures = (int)textures.size(); pod.textures = textures.data();
    // imageStore vectors may have been moved while growing: refresh the pointers
    for (size_t i = 0; i < images.size(); i++) images[i].rgb = imageStore[i].empty() ? nullptr : imageStore[i].data();
    pod.n_images = (int)images.size(); pod.images = images.data();
    pod.n_lights = (int)lights.size(); pod.lights = lights.data();
}

// ---- geometry
void Plane::flatten(FlatScene&, hxr_geometry& g) const { g.type = HXR_GEOM_PLANE; g.p[0] = y; g.p[1] = limit; }
void Sphere::flatten(FlatScene&, hxr_geometry& g) const
{
    g.type = HXR_GEOM_SPHERE;
    putv3(g.p, O);
    g.p[3] = R;
    g.p[4] = uvscaling;
}
void Cube::flatten(FlatScene&, hxr_geometry& g) const
{
    g.type = HXR_GEOM_CUBE;
    putv3(g.p, O);
    g.p[3] = side * 0.5;  // Cube::beginFrame (src/geometry.h:102-105)
}
void CSGBase::flatten(FlatScene&, hxr_geometry& g) const
{
    g.type = HXR_GEOM_CSG;
    g.a = op();
    g.b = left ? left->index : -1;
    g.c = right ? right->index : -1;
}
void Mesh::flatten(FlatScene& fs, hxr_geometry& g) const
{
    hxr_mesh m;
    memset(&m, 0, sizeof m);
    m.n_vertices = (int)vertices.size();
    m.n_normals = (int)normals.size();
    m.n_uvs = (int)uvs.size();
    m.n_triangles = (int)triangles.size();
    m.vertices = fs.keepDoubles(vertices);
    m.normals = fs.keepDoubles(normals);
    m.uvs = fs.keepDoubles(uvs);
    m.triangles = triangles.data();
    m.faceted = faceted;
    m.backface_culling = backfaceCulling;
    putv3(m.bbox_min, bbmin);
    putv3(m.bbox_max, bbmax);
    g.type = HXR_GEOM_MESH;
    g.a = (int)fs.meshes.size();
    fs.meshes.push_back(m);
}
void Heightfield::flatten(FlatScene& fs, hxr_geometry& g) const
{
    hxr_heightfield h;
    memset(&h, 0, sizeof h);
    h.width = W;
    h.height = H;
    h.use_optimization = useOptimization;
    h.max_k = maxK;
    h.heights = heights.data();
    h.max_h = maxH.data();
    h.normals = normals.data();
    h.high_map = highMap.empty() ? nullptr : highMap.data();
    putv3(h.bbox_min, bbmin);
    putv3(h.bbox_max, bbmax);
    g.type = HXR_GEOM_HEIGHTFIELD;
    g.a = (int)fs.heightfields.size();
    fs.heightfields.push_back(h);
}

// ---- textures
void CheckerTexture::flatten(FlatScene&, hxr_texture& t) const
{
    t.type = HXR_TEX_CHECKER;
    putc3(t.color1, color1);
    putc3(t.color2, color2);
    t.scaling = scaling;
}
void BitmapTexture::flatten(FlatScene& fs, hxr_texture& t) const
{
    t.type = HXR_TEX_BITMAP;
    t.image = fs.addImage(bitmap);
    t.scaling = scaling;
}
void Fresnel::flatten(FlatScene&, hxr_texture& t) const { t.type = HXR_TEX_FRESNEL; t.ior = ior; }
void BumpTexture::flatten(FlatScene& fs, hxr_texture& t) const
{
    t.type = HXR_TEX_BUMP;
    t.image = fs.addImage(bitmap);
    t.strength = strength;
    t.scaling = scaling;
}
void Bumps::flatten(FlatScene&, hxr_texture& t) const { t.type = HXR_TEX_BUMPS; t.strength = strength; }

// ---- shaders
void Lambert::flatten(FlatScene&, hxr_shader& s) const
{
    s.type = HXR_SHADER_LAMBERT;
    putc3(s.color, diffuse);
    s.tex = diffuseTex ? diffuseTex->index : -1;
}
void Phong::flatten(FlatScene& fs, hxr_shader& s) const
{
    Lambert::flatten(fs, s);
    s.type = HXR_SHADER_PHONG;
    putc3(s.color2, specular);
    s.f0 = exponent;
}
void Reflection::flatten(FlatScene&, hxr_shader& s) const
{
    s.type = HXR_SHADER_REFLECTION;
    putc3(s.color, reflColor);
    s.f0 = glossiness;
    s.i0 = numSamples;
}
void Refraction::flatten(FlatScene&, hxr_shader& s) const
{
    s.type = HXR_SHADER_REFRACTION;
    putc3(s.color, refrColor);
    s.ior = ior;
}
void Layered::flatten(FlatScene& fs, hxr_shader& s) const
{
    s.type = HXR_SHADER_LAYERED;
    s.first_layer = (int)fs.layers.size();
    s.n_layers = (int)layers.size();
    for (const Layer& l : layers) {
        hxr_layer o;
        o.shader = l.shader->index;
        o.tex = l.blendTex ? l.blendTex->index : -1;
        putc3(o.blend, l.blend);
        fs.layers.push_back(o);
    }
}
void Const::flatten(FlatScene&, hxr_shader& s) const { s.type = HXR_SHADER_CONST; putc3(s.color, color); }

// ---- lights
void PointLight::flatten(hxr_light& l) const
{
    l.type = HXR_LIGHT_POINT;
    l.xsubd = l.ysubd = 1;
    l.power = power;
    putc3(l.color, color);
    l.scale_factor = 1.0f;
    l.area = 0;
    putv3(l.pos, pos);
    Transform I;
    toPodTransform(I, l.T);
}
void RectLight::flatten(hxr_light& l) const
{
    l.type = HXR_LIGHT_RECT;
    l.xsubd = xSubd;
    l.ysubd = ySubd;
    l.power = power;
    putc3(l.color, color);
    // RectLight::beginFrame (src/lights.cpp:75-88): area from three transformed corners
    const Vec3 p0 = T.transformPoint(Vec3(0.5, 0, -0.5)), p1 = T.transformPoint(Vec3(-0.5, 0, -0.5)),
               p2 = T.transformPoint(Vec3(0.5, 0, 0.5));
    l.area = distance(p0, p1) * distance(p0, p2);
    l.scale_factor = (float)(1 / l.area);
    putv3(l.pos, T.offset);
    toPodTransform(T, l.T);
}

bool flattenScene(Scene& sc, FlatScene& fs, std::string& err)
{
    sc.beginRender();
    sc.beginFrame();
    for (size_t i = 0; i < sc.geometries.size(); i++) sc.geometries[i]->index = (int)i;
    for (size_t i = 0; i < sc.textures.size(); i++) sc.textures[i]->index = (int)i;
    for (size_t i = 0; i < sc.shaders.size(); i++) sc.shaders[i]->index = (int)i;
    for (size_t i = 0; i < sc.nodes.size(); i++) sc.nodes[i]->index = (int)i;
    fs = FlatScene();
    memset(&fs.pod, 0, sizeof fs.pod);

    for (Geometry* g : sc.geometries) {
        hxr_geometry o;
        memset(&o, 0, sizeof o);
        g->flatten(fs, o);
        if (o.type == HXR_GEOM_CSG && (o.b < 0 || o.c < 0)) { err = "CSG geometry `" + g->name + "' lacks a child"; return false; }
        fs.geometries.push_back(o);
    }
    for (Texture* t : sc.textures) {
        hxr_texture o;
        memset(&o, 0, sizeof o);
        o.image = -1;
        t->flatten(fs, o);
        fs.textures.push_back(o);
    }
    for (Shader* s : sc.shaders) {
        hxr_shader o;
        memset(&o, 0, sizeof o);
        o.tex = -1;
        s->flatten(fs, o);
        fs.shaders.push_back(o);
    }
    for (Light* l : sc.lights) {
        hxr_light o;
        memset(&o, 0, sizeof o);
        l->flatten(o);
        fs.lights.push_back(o);
    }
    for (Node* n : sc.nodes) {
        if (!n->geom) { err = "Node `" + n->name + "' has no geometry"; return false; }
        hxr_node o;
        memset(&o, 0, sizeof o);
        o.geom = n->geom->index;
        o.shader = n->shader->index;
        o.bump_tex = n->bump ? n->bump->index : -1;
        toPodTransform(n->T, o.T);
        fs.nodes.push_back(o);
    }
    fs.pod.has_environment = 0;
    if (auto* env = dynamic_cast<CubemapEnvironment*>(sc.environment)) {
        if (env->loaded) {
            fs.pod.has_environment = 1;
            for (int i = 0; i < 6; i++) fs.pod.env_images[i] = fs.addImage(env->sides[i]);
        }
    }
    hxr_settings& st = fs.pod.settings;
    st.frame_width = sc.settings.frameWidth;
    st.frame_height = sc.settings.frameHeight;
    st.max_trace_depth = sc.settings.maxTraceDepth;
    st.want_aa = sc.settings.wantAA;
    st.gi = sc.settings.gi;
    st.num_paths = sc.settings.numPaths;
    putc3(st.ambient, sc.settings.ambientLight);
    putc3(st.background, sc.settings.backgroundColor);
    fs.finalize();
    return true;
}

}  // namespace host
}  // namespace hxr
