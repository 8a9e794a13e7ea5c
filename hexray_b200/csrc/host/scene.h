// scene.h — the host-side scene front-end: the `.hexray` language and the element model.
//
// This is the UPSTREAM seam of the hot path: it accepts the reference's scene
// language unchanged and keeps the reference's element interface —
//   SceneElement::fillProperties(ParsedBlock&) / beginRender() / beginFrame()   src/scene.h:60-134
//   ParsedBlock typed getters, same names / ranges / error behaviour            src/scene.h:136-186
//   SceneParser::find*ByName, resolveFullPath                                   src/scene.h:188-210
//   class names and property names of every element                            src/scene.cpp:780-809
//   errors: SyntaxError / FileNotFoundError thrown inside fillProperties, caught by the parser,
//           parseScene() returns false                                          src/scene.cpp:541-551
// — but the elements carry no intersect()/computeColor() code: per-ray work happens on the GPU.
// Instead every element can flatten itself into the POD tables of include/hxr.h.
#pragma once
#include <climits>
#include <memory>
#include <string>
#include <vector>
#include "../../../include/hxr.h"
#include "bitmap.h"
#include "math_types.h"

namespace hxr {
namespace host {

enum ElementType {
    ELEM_GEOMETRY, ELEM_SHADER, ELEM_NODE, ELEM_TEXTURE, ELEM_ENVIRONMENT, ELEM_CAMERA, ELEM_LIGHT, ELEM_SETTINGS,
};

const float LARGE_FLOAT = 1e17f;
const double LARGE_DOUBLE = 1e120;

class ParsedBlock;
class SceneParser;
class Geometry;
class Shader;
class Texture;
struct Node;
struct Scene;
struct FlatScene;

struct SyntaxError {
    std::string msg;
    int line;
    SyntaxError(int line, const std::string& msg) : msg(msg), line(line) {}
};
struct FileNotFoundError {
    std::string filename;
    int line;
    FileNotFoundError(int line, const std::string& fn) : filename(fn), line(line) {}
};

class SceneElement {
public:
    std::string name;
    int index = -1;  // position in its Scene list (assigned after parsing)
    virtual ~SceneElement() {}
    virtual ElementType getElementType() const = 0;
    virtual void fillProperties(ParsedBlock&) {}
    virtual void beginRender() {}
    virtual void beginFrame() {}
};

class ParsedBlock {
public:
    virtual ~ParsedBlock() {}
    virtual bool getIntProp(const char* name, int* value, int minValue = INT_MIN, int maxValue = INT_MAX) = 0;
    virtual bool getBoolProp(const char* name, bool* value) = 0;
    virtual bool getFloatProp(const char* name, float* value, float minValue = -LARGE_FLOAT, float maxValue = LARGE_FLOAT) = 0;
    virtual bool getDoubleProp(const char* name, double* value, double minValue = -LARGE_DOUBLE, double maxValue = LARGE_DOUBLE) = 0;
    virtual bool getColorProp(const char* name, Color3* value, float minCompValue = -LARGE_FLOAT, float maxCompValue = LARGE_FLOAT) = 0;
    virtual bool getVectorProp(const char* name, Vec3* value) = 0;
    virtual bool getGeometryProp(const char* name, Geometry** value) = 0;
    virtual bool getShaderProp(const char* name, Shader** value) = 0;
    virtual bool getTextureProp(const char* name, Texture** value) = 0;
    virtual bool getNodeProp(const char* name, Node** value) = 0;
    virtual bool getStringProp(const char* name, std::string* value) = 0;
    virtual bool getFilenameProp(const char* name, std::string* value) = 0;
    virtual bool getBitmapFileProp(const char* name, Bitmap& value) = 0;
    virtual void getTransformProp(Transform& T) = 0;
    virtual void requiredProp(const char* name) = 0;
    virtual void signalError(const char* msg) = 0;
    virtual void signalWarning(const char* msg) = 0;
    virtual int getBlockLines() = 0;
    virtual void getBlockLine(int idx, int& srcLine, std::string& head, std::string& tail) = 0;
    virtual SceneParser& getParser() = 0;
};

class SceneParser {
public:
    virtual ~SceneParser() {}
    virtual Shader* findShaderByName(const char* name) = 0;
    virtual Texture* findTextureByName(const char* name) = 0;
    virtual Geometry* findGeometryByName(const char* name) = 0;
    virtual Node* findNodeByName(const char* name) = 0;
    virtual bool resolveFullPath(std::string& path) = 0;
};

// ---- elements (property tables: SURVEY.md Appendix B) -------------------------------------

struct GlobalSettings : public SceneElement {
    int frameWidth = 800, frameHeight = 600;
    Color3 ambientLight = Color3(0.15f, 0.15f, 0.15f);
    Color3 backgroundColor = Color3(0, 0, 0);
    bool wantAA = true;
    int maxTraceDepth = 8;
    bool dbg = false;
    int prepassSamples = 5;
    bool gi = false;
    int numPaths = 32;
    int numThreads = 0;
    bool interactive = false;
    int foveatedRadius = 0;
    ElementType getElementType() const override { return ELEM_SETTINGS; }
    void fillProperties(ParsedBlock& pb) override;
};

class Camera : public SceneElement {
public:
    Vec3 pos;
    double yaw = 0, pitch = 0, roll = 0;
    double aspectRatio = 4.0 / 3.0;
    double fov = 90;
    double fNumber = 2.0;
    int numSamples = 32;
    double focalPlaneDist = 100;
    bool dof = false;
    bool autoFocus = false;
    double stereoSeparation = 0;
    ElementType getElementType() const override { return ELEM_CAMERA; }
    void fillProperties(ParsedBlock& pb) override;
    // Camera::beginFrame (src/camera.cpp:30-63) as a pure function of the properties
    void computeFrame(hxr_camera& out) const;
};

class Geometry : public SceneElement {
public:
    ElementType getElementType() const override { return ELEM_GEOMETRY; }
    virtual void flatten(FlatScene& fs, hxr_geometry& g) const = 0;
};

class Plane : public Geometry {
public:
    double y = 0, limit = 1e99;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_geometry& g) const override;
};
class Sphere : public Geometry {
public:
    Vec3 O;
    double R = 1, uvscaling = 1;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_geometry& g) const override;
};
class Cube : public Geometry {
public:
    Vec3 O;
    double side = 1;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_geometry& g) const override;
};
class CSGBase : public Geometry {
public:
    Geometry* left = nullptr;
    Geometry* right = nullptr;
    virtual hxr_csg_op op() const = 0;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_geometry& g) const override;
};
class CSGUnion : public CSGBase { public: hxr_csg_op op() const override { return HXR_CSG_UNION; } };
class CSGInter : public CSGBase { public: hxr_csg_op op() const override { return HXR_CSG_INTER; } };
class CSGDiff : public CSGBase { public: hxr_csg_op op() const override { return HXR_CSG_DIFF; } };

class Mesh : public Geometry {
public:
    std::vector<Vec3> vertices, normals, uvs;  // slot 0 = OBJ sentinel
    std::vector<hxr_triangle> triangles;
    Vec3 bbmin, bbmax;
    bool faceted = false, backfaceCulling = false, useKDTree = true, autoSmooth = false, recenter = false;
    bool loadFromOBJ(const char* filename);
    void prepareTriangles();
    void computeBoundingGeometry();
    // procedural meshes for the large-scene configuration
    void generateTerrain(int gridSide, uint64_t seed);
    void generateSoup(int64_t nTriangles, uint64_t seed);
    bool saveOBJ(const char* filename) const;  // v/vt/vn/f text, so the reference can load a procedural mesh
    void fillProperties(ParsedBlock& pb) override;
    void beginRender() override;
    void flatten(FlatScene& fs, hxr_geometry& g) const override;
};

class Heightfield : public Geometry {
public:
    std::vector<float> heights, maxH, highMap;
    std::vector<double> normals;
    Vec3 bbmin, bbmax;
    int W = 0, H = 0, maxK = 0;
    bool useOptimization = true;
    void buildHighMap();
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_geometry& g) const override;
};

class Texture : public SceneElement {
public:
    ElementType getElementType() const override { return ELEM_TEXTURE; }
    virtual void flatten(FlatScene& fs, hxr_texture& t) const = 0;
};
class CheckerTexture : public Texture {
public:
    Color3 color1 = Color3(1, 1, 1), color2 = Color3(0, 0, 0);
    double scaling = 20.0;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_texture& t) const override;
};
class BitmapTexture : public Texture {
public:
    Bitmap bitmap;
    double scaling = 100.0;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_texture& t) const override;
};
class Fresnel : public Texture {
public:
    double ior = 1.33;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_texture& t) const override;
};
class BumpTexture : public Texture {
public:
    Bitmap bitmap;
    double strength = 1, scaling = 1;
    void fillProperties(ParsedBlock& pb) override;
    void beginRender() override { bitmap.differentiate(); }
    void flatten(FlatScene& fs, hxr_texture& t) const override;
};
class Bumps : public Texture {
public:
    float strength = 1;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_texture& t) const override;
};

class Shader : public SceneElement {
public:
    ElementType getElementType() const override { return ELEM_SHADER; }
    virtual void flatten(FlatScene& fs, hxr_shader& s) const = 0;
};
class Lambert : public Shader {
public:
    Color3 diffuse = Color3(0.5f, 0.5f, 0.5f);
    Texture* diffuseTex = nullptr;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_shader& s) const override;
};
class Phong : public Lambert {
public:
    Color3 specular = Color3(1, 1, 1);
    float exponent = 10.0f;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_shader& s) const override;
};
class Reflection : public Shader {
public:
    float glossiness = 1.0f;
    Color3 reflColor = Color3(0.95f, 0.95f, 0.95f);
    int numSamples = 50;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_shader& s) const override;
};
class Refraction : public Shader {
public:
    Color3 refrColor = Color3(0.95f, 0.95f, 0.95f);
    double ior = 1.33;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_shader& s) const override;
};
class Layered : public Shader {
public:
    struct Layer {
        Shader* shader;
        Color3 blend;
        Texture* blendTex;
    };
    std::vector<Layer> layers;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_shader& s) const override;
};
class Const : public Shader {
public:
    Color3 color = Color3(0.5f, 0.5f, 0.5f);
    void fillProperties(ParsedBlock& pb) override;
    void flatten(FlatScene& fs, hxr_shader& s) const override;
};

class Light : public SceneElement {
public:
    Color3 color = Color3(1, 1, 1);
    float power = 1.0f;
    ElementType getElementType() const override { return ELEM_LIGHT; }
    void fillProperties(ParsedBlock& pb) override;
    virtual void flatten(hxr_light& l) const = 0;
};
class PointLight : public Light {
public:
    Vec3 pos;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(hxr_light& l) const override;
};
class RectLight : public Light {
public:
    Transform T;
    int xSubd = 3, ySubd = 3;
    void fillProperties(ParsedBlock& pb) override;
    void flatten(hxr_light& l) const override;
};

class Environment : public SceneElement {
public:
    bool loaded = false;
    ElementType getElementType() const override { return ELEM_ENVIRONMENT; }
};
class CubemapEnvironment : public Environment {
public:
    Bitmap sides[6];  // NEGX NEGY NEGZ POSX POSY POSZ
    bool loadMaps(const std::string& folder, float gamma);
    void fillProperties(ParsedBlock& pb) override;
};

struct Node : public SceneElement {
    Geometry* geom = nullptr;
    Shader* shader = nullptr;
    Transform T;
    Texture* bump = nullptr;
    ElementType getElementType() const override { return ELEM_NODE; }
    void fillProperties(ParsedBlock& pb) override;
};

// ---- the scene ---------------------------------------------------------------------------

struct Scene {
    std::vector<std::unique_ptr<SceneElement>> owned;
    std::vector<Geometry*> geometries;
    std::vector<Shader*> shaders;
    std::vector<Node*> nodes, superNodes;
    std::vector<Texture*> textures;
    std::vector<Light*> lights;
    Environment* environment = nullptr;
    Camera* camera = nullptr;
    GlobalSettings settings;
    std::string lastError;

    bool parseScene(const char* sceneFile);  // false on any syntax / missing-file error (message in lastError)
    void beginRender();
    void beginFrame();
};

// POD tables + the storage they point into
struct FlatScene {
    hxr_scene pod;
    std::vector<hxr_node> nodes;
    std::vector<hxr_geometry> geometries;
    std::vector<hxr_mesh> meshes;
    std::vector<hxr_heightfield> heightfields;
    std::vector<hxr_shader> shaders;
    std::vector<hxr_layer> layers;
    std::vector<hxr_texture> textures;
    std::vector<hxr_image> images;
    std::vector<hxr_light> lights;
    std::vector<std::vector<float>> imageStore;
    std::vector<std::vector<double>> doubleStore;
    int addImage(const Bitmap& bmp);
    const double* keepDoubles(const std::vector<Vec3>& v);
    void finalize();
};

// scene.beginRender() + beginFrame() + flatten, in the reference's visiting order
bool flattenScene(Scene& scene, FlatScene& out, std::string& err);
void toPodTransform(const Transform& T, hxr_transform& out);

}  // namespace host
}  // namespace hxr
