// scene.cpp — `.hexray` parser and element property tables (see scene.h).
// Language reference: SURVEY.md Appendix B, derived from reference src/scene.cpp:401-568
// (lexing, two-pass processing order), :135-356 (typed getters) and the per-class
// fillProperties() in src/*.h. Written from that description; no reference code is reused.
#include "scene.h"
#include <algorithm>
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <random>

namespace hxr {
namespace host {

// ======================================================================= lexical helpers
static std::string trimmed(const std::string& s)
{
    size_t b = 0, e = s.size();
    while (e > b && isspace((unsigned char)s[e - 1])) e--;
    while (b < e && isspace((unsigned char)s[b])) b++;
    return s.substr(b, e - b);
}

static std::vector<std::string> splitWhitespace(const std::string& s)
{
    std::vector<std::string> out;
    size_t i = 0, n = s.size();
    while (i < n) {
        while (i < n && isspace((unsigned char)s[i])) i++;
        if (i >= n) break;
        size_t j = i;
        while (j < n && !isspace((unsigned char)s[j])) j++;
        out.push_back(s.substr(i, j - i));
        i = j;
    }
    return out;
}

static std::string formatted(const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return buf;
}

// "(1, 2, 3)" / "1 2 3" / "(278. 273. -800)" -> three numbers
static bool parseTriple(std::string s, double& a, double& b, double& c)
{
    for (auto& ch : s)
        if (ch == ',' || ch == '(' || ch == ')') ch = ' ';
    return sscanf(s.c_str(), "%lf%lf%lf", &a, &b, &c) == 3;
}

// randfloat(a,b) / randint(a,b) are replaced textually, drawing from ONE default-seeded
// mt19937 shared by the whole parse (src/scene.cpp:607-654)
static void substituteRandoms(int srcLine, std::string& s, std::mt19937& gen)
{
    for (int pass = 0; pass < 2; pass++) {
        const char* key = pass == 0 ? "randfloat" : "randint";
        size_t p;
        while ((p = s.find(key)) != std::string::npos) {
            size_t open = s.find('(', p), close = open == std::string::npos ? open : s.find(')', open);
            if (open == std::string::npos || close == std::string::npos)
                throw SyntaxError(srcLine, std::string(key) + " in inexpected format");
            std::string args = s.substr(open + 1, close - open - 1);
            std::string repl;
            if (pass == 0) {
                float f1, f2;
                if (sscanf(args.c_str(), "%f,%f", &f1, &f2) != 2) throw SyntaxError(srcLine, "bad randfloat format (expected: randfloat(<min>, <max>))");
                if (f1 > f2) throw SyntaxError(srcLine, "bad randfloat format (min > max)");
                std::uniform_real_distribution<float> d(f1, f2);
                repl = formatted("%.5f", d(gen));
            } else {
                int i1, i2;
                if (sscanf(args.c_str(), "%d,%d", &i1, &i2) != 2) throw SyntaxError(srcLine, "bad randint format (expected: randint(<min>, <max>))");
                if (i1 > i2) throw SyntaxError(srcLine, "bad randint format (min > max)");
                std::uniform_int_distribution<int> d(i1, i2);
                repl = formatted("%d", d(gen));
            }
            // the call is overwritten in place and padded with blanks
            std::string pad(close - p + 1, ' ');
            pad.replace(0, std::min(repl.size(), pad.size()), repl.substr(0, pad.size()));
            s.replace(p, close - p + 1, pad);
        }
    }
}

// ======================================================================= ParsedBlock
namespace {

class Parser;

class Block : public ParsedBlock {
public:
    struct Line {
        int line;
        std::string name, value;
        bool recognized;
    };
    std::vector<Line> lines;
    int blockBegin = 0, blockEnd = 0;
    Parser* parser = nullptr;
    SceneElement* element = nullptr;

    Line* find(const char* name)
    {
        for (auto& l : lines)
            if (l.name == name) {
                l.recognized = true;
                return &l;
            }
        return nullptr;
    }
    bool getIntProp(const char* name, int* value, int lo, int hi) override
    {
        Line* l = find(name);
        if (!l) return false;
        int x;
        if (sscanf(l->value.c_str(), "%d", &x) != 1) throw SyntaxError(l->line, "Invalid integer");
        if (x < lo || x > hi) throw SyntaxError(l->line, formatted("Value outside the allowed bounds (%d .. %d)\n", lo, hi));
        *value = x;
        return true;
    }
    bool getBoolProp(const char* name, bool* value) override
    {
        Line* l = find(name);
        if (!l) return false;
        *value = !(l->value == "off" || l->value == "false" || l->value == "0");
        return true;
    }
    bool getFloatProp(const char* name, float* value, float lo, float hi) override
    {
        Line* l = find(name);
        if (!l) return false;
        float x;
        if (sscanf(l->value.c_str(), "%f", &x) != 1) throw SyntaxError(l->line, "Invalid float");
        if (x < lo || x > hi) throw SyntaxError(l->line, formatted("Value outside the allowed bounds (%f .. %f)\n", lo, hi));
        *value = x;
        return true;
    }
    bool getDoubleProp(const char* name, double* value, double lo, double hi) override
    {
        Line* l = find(name);
        if (!l) return false;
        double x;
        if (sscanf(l->value.c_str(), "%lf", &x) != 1) throw SyntaxError(l->line, "Invalid double");
        if (x < lo || x > hi) throw SyntaxError(l->line, formatted("Value outside the allowed bounds (%f .. %f)\n", lo, hi));
        *value = x;
        return true;
    }
    bool getColorProp(const char* name, Color3* value, float lo, float hi) override
    {
        Line* l = find(name);
        if (!l) return false;
        double r, g, b;
        if (!parseTriple(l->value, r, g, b)) throw SyntaxError(l->line, "Invalid color");
        Color3 c((float)r, (float)g, (float)b);
        const char* ch[3] = {"R", "G", "B"};
        const float v[3] = {c.r, c.g, c.b};
        for (int i = 0; i < 3; i++)
            if (v[i] < lo || v[i] > hi)
                throw SyntaxError(l->line, formatted("Color %s value outside the allowed bounds (%f .. %f)\n", ch[i], lo, hi));
        *value = c;
        return true;
    }
    bool getVectorProp(const char* name, Vec3* value) override
    {
        Line* l = find(name);
        if (!l) return false;
        Vec3 v;
        if (!parseTriple(l->value, v.x, v.y, v.z)) throw SyntaxError(l->line, "Invalid vector");
        *value = v;
        return true;
    }
    bool getGeometryProp(const char* name, Geometry** value) override;
    bool getShaderProp(const char* name, Shader** value) override;
    bool getTextureProp(const char* name, Texture** value) override;
    bool getNodeProp(const char* name, Node** value) override;
    bool getStringProp(const char* name, std::string* value) override
    {
        Line* l = find(name);
        if (!l) return false;
        *value = l->value;
        return true;
    }
    bool getFilenameProp(const char* name, std::string* value) override;
    bool getBitmapFileProp(const char* name, Bitmap& bmp) override;
    void getTransformProp(Transform& T) override
    {
        for (auto& l : lines) {
            const bool sc = l.name == "scale", ro = l.name == "rotate", tr = l.name == "translate";
            if (!sc && !ro && !tr) continue;
            l.recognized = true;
            double x, y, z;
            if (!parseTriple(l.value, x, y, z)) throw SyntaxError(l.line, "Expected three double values");
            if (sc) T.scale(x, y, z);
            else if (ro) T.rotate(x, y, z);
            else T.translate(Vec3(x, y, z));
        }
    }
    void requiredProp(const char* name) override
    {
        if (!find(name)) throw SyntaxError(blockEnd, formatted("Required property `%s' not defined", name));
    }
    void signalError(const char* msg) override { throw SyntaxError(blockEnd, msg); }
    void signalWarning(const char* msg) override { fprintf(stderr, "Warning (at line %d): %s\n", blockEnd, msg); }
    int getBlockLines() override { return (int)lines.size(); }
    void getBlockLine(int idx, int& srcLine, std::string& head, std::string& tail) override
    {
        lines[idx].recognized = true;
        srcLine = lines[idx].line;
        head = lines[idx].name;
        tail = lines[idx].value;
    }
    SceneParser& getParser() override;
};

class Parser : public SceneParser {
public:
    Scene* s = nullptr;
    std::string rootDir;
    template <class T> static T* byName(const std::vector<T*>& v, const char* name)
    {
        for (T* e : v)
            if (e->name == name) return e;
        return nullptr;
    }
    Shader* findShaderByName(const char* name) override { return byName(s->shaders, name); }
    Texture* findTextureByName(const char* name) override { return byName(s->textures, name); }
    Geometry* findGeometryByName(const char* name) override { return byName(s->geometries, name); }
    Node* findNodeByName(const char* name) override { return byName(s->nodes, name); }
    bool resolveFullPath(std::string& path) override
    {
        std::string full = rootDir + path;
        if (!std::filesystem::exists(full)) return false;
        path = full;
        return true;
    }
    SceneElement* create(const std::string& cls);
    bool parse(const char* filename, Scene* scene);
};

SceneParser& Block::getParser() { return *parser; }

bool Block::getGeometryProp(const char* name, Geometry** value)
{
    Line* l = find(name);
    if (!l) return false;
    Geometry* g = parser->findGeometryByName(l->value.c_str());
    if (!g) throw SyntaxError(l->line, "Geometry not defined");
    *value = g;
    return true;
}
bool Block::getShaderProp(const char* name, Shader** value)
{
    Line* l = find(name);
    if (!l) return false;
    Shader* x = parser->findShaderByName(l->value.c_str());
    if (!x) throw SyntaxError(l->line, "Shader not defined");
    *value = x;
    return true;
}
bool Block::getTextureProp(const char* name, Texture** value)
{
    Line* l = find(name);
    if (!l) return false;
    Texture* x = parser->findTextureByName(l->value.c_str());
    if (!x) throw SyntaxError(l->line, "Texture not defined");
    *value = x;
    return true;
}
bool Block::getNodeProp(const char* name, Node** value)
{
    Line* l = find(name);
    if (!l) return false;
    Node* x = parser->findNodeByName(l->value.c_str());
    if (!x) throw SyntaxError(l->line, "Node not defined");
    *value = x;
    return true;
}
bool Block::getFilenameProp(const char* name, std::string* value)
{
    Line* l = find(name);
    if (!l) return false;
    std::string p = l->value;
    if (!parser->resolveFullPath(p)) throw FileNotFoundError(l->line, l->value);
    *value = p;
    return true;
}
bool Block::getBitmapFileProp(const char* name, Bitmap& bmp)
{
    Line* l = find(name);
    if (!l) return false;
    std::string p = l->value;
    if (!parser->resolveFullPath(p)) throw FileNotFoundError(l->line, p);
    return bmp.loadImage(p.c_str());
}

SceneElement* Parser::create(const std::string& c)
{
    if (c == "GlobalSettings") return &s->settings;
    SceneElement* e = nullptr;
    if (c == "Plane") e = new Plane;
    else if (c == "Sphere") e = new Sphere;
    else if (c == "Cube") e = new Cube;
    else if (c == "CSGUnion") e = new CSGUnion;
    else if (c == "CSGInter") e = new CSGInter;
    else if (c == "CSGDiff") e = new CSGDiff;
    else if (c == "Lambert") e = new Lambert;
    else if (c == "Phong") e = new Phong;
    else if (c == "CheckerTexture") e = new CheckerTexture;
    else if (c == "BitmapTexture") e = new BitmapTexture;
    else if (c == "Reflection") e = new Reflection;
    else if (c == "Refraction") e = new Refraction;
    else if (c == "Layered") e = new Layered;
    else if (c == "Fresnel") e = new Fresnel;
    else if (c == "Node") e = new Node;
    else if (c == "CubemapEnvironment") e = new CubemapEnvironment;
    else if (c == "Camera") e = new Camera;
    else if (c == "Mesh") e = new Mesh;
    else if (c == "Heightfield") e = new Heightfield;
    else if (c == "BumpTexture") e = new BumpTexture;
    else if (c == "Bumps") e = new Bumps;
    else if (c == "Const") e = new Const;
    else if (c == "PointLight") e = new PointLight;
    else if (c == "RectLight") e = new RectLight;
    if (e) s->owned.emplace_back(e);
    return e;
}

bool Parser::parse(const char* filename, Scene* scene)
{
    s = scene;
    auto fail = [&](const std::string& m) {
        s->lastError = m;
        fprintf(stderr, "%s\n", m.c_str());
        return false;
    };
    FILE* f = fopen(filename, "rt");
    if (!f) return fail(formatted("Cannot open scene file `%s'!", filename));
    {
        std::string fn(filename);
        size_t slash = fn.find_last_of("/\\");
        rootDir = slash == std::string::npos ? "" : fn.substr(0, slash + 1);
    }
    std::vector<std::unique_ptr<Block>> blocks;
    Block* cur = nullptr;
    SceneElement* curObj = nullptr;
    bool inComment = false;
    int lineNo = 0;
    char raw[1024];
    std::mt19937 randGen;
    std::string failMsg;
    while (fgets(raw, sizeof raw, f)) {
        lineNo++;
        if (inComment) {
            if (raw[0] == '*' && raw[1] == '/') inComment = false;
            continue;
        }
        std::string line(raw);
        size_t c1 = line.find("//"), c2 = line.find('#');
        size_t cut = std::min(c1, c2);
        if (cut != std::string::npos) line.erase(cut);
        line = trimmed(line);
        if (line.empty()) continue;
        if (line.size() >= 2 && line[0] == '/' && line[1] == '*') {
            inComment = true;
            continue;
        }
        try {
            substituteRandoms(lineNo, line, randGen);
        } catch (SyntaxError& e) {
            failMsg = formatted("%s:%d: Syntax error on line %d: %s", filename, e.line, e.line, e.msg.c_str());
            break;
        }
        std::vector<std::string> tok = splitWhitespace(line);
        if (tok.empty()) continue;
        if (!curObj) {
            if (tok.size() == 1) {
                failMsg = tok[0] == "{" ? formatted("Excess `}' on line %d", lineNo) : formatted("Unexpected token `%s' on line %d", tok[0].c_str(), lineNo);
                break;
            }
            if (tok.size() > 3) { failMsg = formatted("Unexpected content on line %d!", lineNo); break; }
            if (tok.back() != "{") {
                failMsg = formatted(tok.size() == 2 ? "A singleton object definition should end with a `{' (on line %d)" : "A object definition should end with a `{' (on line %d)", lineNo);
                break;
            }
            curObj = create(tok[0]);
            if (!curObj) { failMsg = formatted("Unknown object class `%s' on line %d", tok[0].c_str(), lineNo); break; }
            curObj->name = tok.size() == 3 ? tok[1] : std::string();
            blocks.emplace_back(new Block);
            cur = blocks.back().get();
            cur->parser = this;
            cur->element = curObj;
            cur->blockBegin = lineNo;
            switch (curObj->getElementType()) {
                case ELEM_GEOMETRY: s->geometries.push_back(static_cast<Geometry*>(curObj)); break;
                case ELEM_SHADER: s->shaders.push_back(static_cast<Shader*>(curObj)); break;
                case ELEM_TEXTURE: s->textures.push_back(static_cast<Texture*>(curObj)); break;
                case ELEM_LIGHT: s->lights.push_back(static_cast<Light*>(curObj)); break;
                case ELEM_NODE: s->nodes.push_back(static_cast<Node*>(curObj)); break;
                case ELEM_ENVIRONMENT: s->environment = static_cast<Environment*>(curObj); break;
                case ELEM_CAMERA: s->camera = static_cast<Camera*>(curObj); break;
                case ELEM_SETTINGS: break;
            }
        } else if (tok.size() == 1) {
            if (tok[0] != "}") {
                failMsg = formatted("Unexpected token in object definition on line %d: `%s'", lineNo, tok[0].c_str());
                break;
            }
            cur->blockEnd = lineNo;
            cur = nullptr;
            curObj = nullptr;
        } else {
            // property: first token is the name, the rest of the line the value (quotes stripped)
            size_t i = tok[0].size();
            while (i < line.size() && isspace((unsigned char)line[i])) i++;
            std::string value = line.substr(i);
            if (value.size() >= 2 && value.front() == '"' && value.back() == '"') value = value.substr(1, value.size() - 2);
            if (value.size() > 255) value.resize(255);
            cur->lines.push_back(Block::Line{lineNo, tok[0], value, false});
        }
    }
    fclose(f);
    if (!failMsg.empty()) return fail(failMsg);
    if (curObj) return fail("Unfinished object definition at EOF!");

    static const ElementType order[] = {ELEM_SETTINGS, ELEM_CAMERA, ELEM_ENVIRONMENT, ELEM_GEOMETRY,
                                        ELEM_TEXTURE, ELEM_SHADER, ELEM_LIGHT, ELEM_NODE};
    for (ElementType et : order)
        for (auto& b : blocks) {
            if (b->element->getElementType() != et) continue;
            try {
                b->element->fillProperties(*b);
            } catch (SyntaxError& e) {
                return fail(formatted("%s:%d: Syntax error on line %d: %s", filename, e.line, e.line, e.msg.c_str()));
            } catch (FileNotFoundError& e) {
                return fail(formatted("%s:%d: Required file not found (%s) (required at line %d)", filename, e.line, e.filename.c_str(), e.line));
            }
            for (auto& l : b->lines)
                if (!l.recognized)
                    fprintf(stderr, "%s:%d: Warning: the property `%s' isn't recognized!\n", filename, l.line, l.name.c_str());
        }
    // nodes without a shader are not scene objects (src/scene.cpp:560-565)
    for (int i = (int)s->nodes.size() - 1; i >= 0; i--)
        if (!s->nodes[i]->shader) {
            s->superNodes.push_back(s->nodes[i]);
            s->nodes.erase(s->nodes.begin() + i);
        }
    if (!s->camera) return fail(formatted("%s: the scene defines no Camera", filename));
    return true;
}

}  // namespace

bool Scene::parseScene(const char* sceneFile)
{
    Parser p;
    return p.parse(sceneFile, this);
}

static void visitAll(Scene& sc, void (SceneElement::*fn)())
{
    for (auto* e : sc.geometries) (e->*fn)();
    for (auto* e : sc.textures) (e->*fn)();
    for (auto* e : sc.shaders) (e->*fn)();
    for (auto* e : sc.superNodes) (e->*fn)();
    for (auto* e : sc.nodes) (e->*fn)();
    for (auto* e : sc.lights) (e->*fn)();
    if (sc.camera) (sc.camera->*fn)();
    (sc.settings.*fn)();
    if (sc.environment) (sc.environment->*fn)();
}
void Scene::beginRender() { visitAll(*this, &SceneElement::beginRender); }
void Scene::beginFrame() { visitAll(*this, &SceneElement::beginFrame); }

// ======================================================================= property tables
void GlobalSettings::fillProperties(ParsedBlock& pb)
{
    pb.getIntProp("frameWidth", &frameWidth);
    pb.getIntProp("frameHeight", &frameHeight);
    pb.getColorProp("ambientLight", &ambientLight);
    pb.getIntProp("maxTraceDepth", &maxTraceDepth);
    pb.getBoolProp("dbg", &dbg);
    pb.getBoolProp("wantAA", &wantAA);
    pb.getIntProp("prepassSamples", &prepassSamples, 0);
    pb.getBoolProp("gi", &gi);
    pb.getIntProp("numPaths", &numPaths, 1);
    pb.getIntProp("numThreads", &numThreads, 0, 1024);
    pb.getBoolProp("interactive", &interactive);
    pb.getIntProp("foveatedRadius", &foveatedRadius, 0, 1000);
}

void Camera::fillProperties(ParsedBlock& pb)
{
    if (!pb.getVectorProp("pos", &pos)) pb.requiredProp("pos");
    pb.getDoubleProp("aspectRatio", &aspectRatio, 1e-6);
    pb.getDoubleProp("fov", &fov, 0.0001, 179);
    pb.getDoubleProp("yaw", &yaw);
    pb.getDoubleProp("pitch", &pitch, -90, 90);
    pb.getDoubleProp("roll", &roll);
    pb.getDoubleProp("fNumber", &fNumber, 0.5, 128.0);
    pb.getIntProp("numSamples", &numSamples, 1);
    pb.getDoubleProp("focalPlaneDist", &focalPlaneDist, 1e-3, 1e+6);
    pb.getBoolProp("dof", &dof);
    pb.getBoolProp("autoFocus", &autoFocus);
    pb.getDoubleProp("stereoSeparation", &stereoSeparation, 0.0);
}

static void put3(double* dst, const Vec3& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

void Camera::computeFrame(hxr_camera& out) const
{
    // the fov is the corner-to-centre angle: scale the unit screen so that its corner sits at tan(fov/2)
    const double corner = distance(Vec3(0, 0, 1), Vec3(-aspectRatio, 1, 1));
    const double k = std::tan(toRadians(fov / 2)) / corner;
    Vec3 tl(-aspectRatio * k, +k, 1), tr(+aspectRatio * k, +k, 1), bl(-aspectRatio * k, -k, 1);
    const Mat3 R = rotationAroundZ(toRadians(roll)) * rotationAroundX(toRadians(pitch)) * rotationAroundY(toRadians(yaw));
    tl = tl * R + pos;
    tr = tr * R + pos;
    bl = bl * R + pos;
    const Vec3 up = normalized(tl - bl), right = normalized(tr - tl);
    memset(&out, 0, sizeof out);
    put3(out.pos, pos);
    put3(out.top_left, tl);
    put3(out.top_right, tr);
    put3(out.bottom_left, bl);
    put3(out.up, up);
    put3(out.right, right);
    put3(out.front, cross(right, up));
    out.aperture_size = 2.5 / fNumber;
    out.focal_plane_dist = focalPlaneDist;
    out.stereo_separation = stereoSeparation;
    out.dof = dof;
    out.auto_focus = autoFocus;
    out.num_samples = numSamples;
}

void Plane::fillProperties(ParsedBlock& pb)
{
    pb.getDoubleProp("y", &y);
    pb.getDoubleProp("limit", &limit);
}
void Sphere::fillProperties(ParsedBlock& pb)
{
    pb.getVectorProp("O", &O);
    pb.getDoubleProp("R", &R, 0.0);
    pb.getDoubleProp("uvscaling", &uvscaling, 1e-6);
}
void Cube::fillProperties(ParsedBlock& pb)
{
    pb.getVectorProp("O", &O);
    pb.getDoubleProp("side", &side, 0.0);
}
void CSGBase::fillProperties(ParsedBlock& pb)
{
    pb.requiredProp("left");
    pb.requiredProp("right");
    pb.getGeometryProp("left", &left);
    pb.getGeometryProp("right", &right);
}

void Mesh::fillProperties(ParsedBlock& pb)
{
    pb.getBoolProp("faceted", &faceted);
    pb.getBoolProp("backfaceCulling", &backfaceCulling);
    pb.getBoolProp("useKDTree", &useKDTree);
    pb.getBoolProp("autoSmooth", &autoSmooth);
    pb.getBoolProp("recenter", &recenter);
    std::string fn;
    if (pb.getStringProp("file", &fn) && fn.rfind("synthetic:", 0) == 0) {
        // extension for the large-scene configuration: "synthetic:terrain:<gridSide>:<seed>" / "synthetic:soup:<nTris>:<seed>"
        char kind[32];
        long long n = 0;
        unsigned long long seed = 0;
        if (sscanf(fn.c_str(), "synthetic:%31[^:]:%lld:%lli", kind, &n, (long long*)&seed) < 2) pb.signalError("bad synthetic mesh spec");
        if (!strcmp(kind, "terrain")) generateTerrain((int)n, seed);
        else if (!strcmp(kind, "soup")) generateSoup(n, seed);
        else pb.signalError("unknown synthetic mesh kind");
        return;
    }
    if (pb.getFilenameProp("file", &fn)) {
        if (!loadFromOBJ(fn.c_str())) pb.signalError("Could not parse OBJ file!");
    } else {
        pb.requiredProp("file");
    }
}

void Heightfield::fillProperties(ParsedBlock& pb)
{
    pb.getBoolProp("useOptimization", &useOptimization);
    Bitmap bmp;
    if (!pb.getBitmapFileProp("file", bmp)) pb.requiredProp("file");
    W = bmp.getWidth();
    H = bmp.getHeight();
    if (W <= 0 || H <= 0) pb.signalError("Heightfield: could not load the height bitmap");
    double blur = 0;
    pb.getDoubleProp("blur", &blur, 0, 1000);
    heights.assign((size_t)W * H, 0.0f);
    float minY = LARGE_FLOAT, maxY = -LARGE_FLOAT;
    if (blur <= 0) {
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                float h = bmp.getPixel(x, y).intensity();
                heights[(size_t)y * W + x] = h;
                minY = std::min(minY, h);
                maxY = std::max(maxY, h);
            }
    } else {
        // grey-scale, then a truncated, un-normalised Gaussian of radius R = min(128, round(3*blur))
        std::vector<float> grey((size_t)W * H);
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) grey[(size_t)y * W + x] = bmp.getPixel(x, y).intensity();
        auto greyAt = [&](int x, int y) -> float { return (x < 0 || x >= W || y < 0 || y >= H) ? 0.0f : grey[(size_t)y * W + x]; };
        const int R = std::min(128, (int)std::floor(float(3 * blur) + 0.5f));
        std::vector<float> gauss((size_t)R * R);
        for (int y = 0; y < R; y++)
            for (int x = 0; x < R; x++)
                gauss[(size_t)y * R + x] = float(std::exp(-(double(x) * x + double(y) * y) / (2 * blur * blur)) / (2 * kPi * blur * blur));
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                float sum = 0;
                for (int dy = -R + 1; dy < R; dy++)
                    for (int dx = -R + 1; dx < R; dx++) sum += gauss[(size_t)std::abs(dy) * R + std::abs(dx)] * greyAt(x + dx, y + dy);
                heights[(size_t)y * W + x] = sum;
                minY = std::min(minY, sum);
                maxY = std::max(maxY, sum);
            }
    }
    bbmin = Vec3(0, minY, 0);
    bbmax = Vec3(W, maxY, H);
    maxH.assign((size_t)W * H, 0.0f);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float m = heights[(size_t)y * W + x];
            if (x < W - 1) m = std::max(m, heights[(size_t)y * W + x + 1]);
            if (y < H - 1) {
                m = std::max(m, heights[(size_t)(y + 1) * W + x]);
                if (x < W - 1) m = std::max(m, heights[(size_t)(y + 1) * W + x + 1]);
            }
            maxH[(size_t)y * W + x] = m;
        }
    normals.assign((size_t)W * H * 3, 0.0);
    auto setN = [&](int x, int y, const Vec3& n) { double* p = &normals[((size_t)y * W + x) * 3]; p[0] = n.x; p[1] = n.y; p[2] = n.z; };
    auto getN = [&](int x, int y) { const double* p = &normals[((size_t)y * W + x) * 3]; return Vec3(p[0], p[1], p[2]); };
    for (int y = 0; y < H - 1; y++)
        for (int x = 0; x < W - 1; x++) {
            float h0 = heights[(size_t)y * W + x], hdx = heights[(size_t)y * W + x + 1], hdy = heights[(size_t)(y + 1) * W + x];
            Vec3 n = cross(Vec3(0, hdy - h0, 1), Vec3(1, hdx - h0, 0));
            n.normalize();
            setN(x, y, n);
        }
    if (W >= 2) for (int y = 0; y < H; y++) setN(W - 1, y, getN(W - 2, y));
    if (H >= 2) for (int x = 0; x < W; x++) setN(x, H - 1, getN(x, H - 2));
    if (useOptimization) buildHighMap();
}

// max-height pyramid: level 0 = 3x3 neighbourhood, level k = 4 diagonal taps of level k-1 at
// offset 2^(k-1), all with clamped addressing (src/heightfield.cpp:49-81)
void Heightfield::buildHighMap()
{
    highMap.assign((size_t)W * H * 16, 0.0f);
    maxK = (int)std::ceil(std::log((double)W) / std::log(2.0));
    if (maxK > 16) maxK = 16;
    auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
    auto hAt = [&](int x, int y) { return heights[(size_t)clampi(y, 0, H - 1) * W + clampi(x, 0, W - 1)]; };
    auto mAt = [&](int x, int y, int k) { return highMap[((size_t)clampi(y, 0, H - 1) * W + clampi(x, 0, W - 1)) * 16 + k]; };
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float r = hAt(x, y);
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) r = std::max(r, hAt(x + dx, y + dy));
            highMap[((size_t)y * W + x) * 16] = r;
        }
    for (int k = 1; k < maxK; k++) {
        const int o = 1 << (k - 1);
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                float r = mAt(x - o, y - o, k - 1);
                r = std::max(r, mAt(x + o, y - o, k - 1));
                r = std::max(r, mAt(x - o, y + o, k - 1));
                r = std::max(r, mAt(x + o, y + o, k - 1));
                highMap[((size_t)y * W + x) * 16 + k] = r;
            }
    }
}

void CheckerTexture::fillProperties(ParsedBlock& pb)
{
    pb.getColorProp("color1", &color1);
    pb.getColorProp("color2", &color2);
    pb.getDoubleProp("scaling", &scaling);
}
void BitmapTexture::fillProperties(ParsedBlock& pb)
{
    pb.getDoubleProp("scaling", &scaling);
    if (!pb.getBitmapFileProp("file", bitmap)) pb.requiredProp("file");
    float assumedGamma = 1.0f;
    pb.getFloatProp("assumedGamma", &assumedGamma, 0.1f, 100.0f);
    if (assumedGamma != 1.0f) bitmap.decompressGamma(assumedGamma);
}
void Fresnel::fillProperties(ParsedBlock& pb) { pb.getDoubleProp("ior", &ior, 1e-6, 10); }
void BumpTexture::fillProperties(ParsedBlock& pb)
{
    pb.getDoubleProp("strength", &strength);
    pb.getDoubleProp("scaling", &scaling);
    if (!pb.getBitmapFileProp("file", bitmap)) pb.requiredProp("file");
}
void Bumps::fillProperties(ParsedBlock& pb) { pb.getFloatProp("strength", &strength); }

void Lambert::fillProperties(ParsedBlock& pb)
{
    pb.getColorProp("color", &diffuse);
    pb.getTextureProp("texture", &diffuseTex);
}
void Phong::fillProperties(ParsedBlock& pb)
{
    pb.getColorProp("color", &diffuse);
    pb.getColorProp("specular", &specular);
    pb.getTextureProp("texture", &diffuseTex);
    pb.getFloatProp("exponent", &exponent);
}
void Reflection::fillProperties(ParsedBlock& pb)
{
    double multiplier;
    if (pb.getDoubleProp("multiplier", &multiplier)) reflColor = Color3((float)multiplier, (float)multiplier, (float)multiplier);
    else pb.getColorProp("reflColor", &reflColor);
    pb.getFloatProp("glossiness", &glossiness, 0, 1);
    pb.getIntProp("numSamples", &numSamples, 1);
}
void Refraction::fillProperties(ParsedBlock& pb)
{
    double multiplier;
    if (pb.getDoubleProp("multiplier", &multiplier)) refrColor = Color3((float)multiplier, (float)multiplier, (float)multiplier);
    else pb.getColorProp("refrColor", &refrColor);
    pb.getDoubleProp("ior", &ior, 1e-6, 10);
}

// "layer <shader>, (r, g, b)[, <texture>]" — repeated; texture "NULL" means none (src/shading.cpp:270-312)
void Layered::fillProperties(ParsedBlock& pb)
{
    auto stripPunct = [](std::string s) {
        std::string o;
        for (char c : s)
            if (!isspace((unsigned char)c) && c != ',') o += c;
        return o;
    };
    for (int i = 0; i < pb.getBlockLines(); i++) {
        int srcLine;
        std::string head, tail;
        pb.getBlockLine(i, srcLine, head, tail);
        if (head != "layer") continue;
        const char* expect = "Expected a line like `layer <shader>, <color>[, <texture>]'";
        // front token = shader name
        size_t b = 0;
        while (b < tail.size() && isspace((unsigned char)tail[b])) b++;
        size_t e = b;
        while (e < tail.size() && !isspace((unsigned char)tail[e])) e++;
        if (b == tail.size() || e == tail.size()) throw SyntaxError(srcLine, expect);
        std::string shaderName = stripPunct(tail.substr(b, e - b));
        std::string rest = tail.substr(e);
        std::string textureName;
        if (rest.empty()) throw SyntaxError(srcLine, expect);
        if (rest.back() != ')') {
            // last token = texture name
            size_t te = rest.size();
            while (te > 0 && isspace((unsigned char)rest[te - 1])) te--;
            size_t tb = te;
            while (tb > 0 && !isspace((unsigned char)rest[tb - 1])) tb--;
            if (te == 0 || tb == 0) throw SyntaxError(srcLine, expect);
            textureName = stripPunct(rest.substr(tb, te - tb));
            rest = rest.substr(0, tb);
        }
        if (textureName == "NULL") textureName.clear();
        Shader* sh = pb.getParser().findShaderByName(shaderName.c_str());
        if (!sh) throw SyntaxError(srcLine, expect);
        Texture* tx = nullptr;
        if (!textureName.empty()) {
            tx = pb.getParser().findTextureByName(textureName.c_str());
            if (!tx) throw SyntaxError(srcLine, expect);
        }
        double x, y, z;
        if (!parseTriple(rest, x, y, z)) throw SyntaxError(srcLine, "Expected three double values");
        layers.push_back(Layer{sh, Color3((float)x, (float)y, (float)z), tx});
    }
}
void Const::fillProperties(ParsedBlock& pb) { pb.getColorProp("color", &color); }

void Light::fillProperties(ParsedBlock& pb)
{
    pb.getColorProp("color", &color);
    pb.getFloatProp("power", &power);
}
void PointLight::fillProperties(ParsedBlock& pb)
{
    Light::fillProperties(pb);
    pb.getVectorProp("pos", &pos);
}
void RectLight::fillProperties(ParsedBlock& pb)
{
    Light::fillProperties(pb);
    pb.getTransformProp(T);
    pb.getIntProp("xSubd", &xSubd, 1, 1000);
    pb.getIntProp("ySubd", &ySubd, 1, 1000);
}

bool CubemapEnvironment::loadMaps(const std::string& folder, float gamma)
{
    static const char* prefixes[2] = {"neg", "pos"};
    static const char* axes[3] = {"x", "y", "z"};
    static const char* suffixes[2] = {".bmp", ".exr"};
    int n = 0;
    for (int pi = 0; pi < 2; pi++)
        for (int a = 0; a < 3; a++) {
            for (int si = 0; si < 2; si++) {
                std::string fn = folder + "/" + prefixes[pi] + axes[a] + suffixes[si];
                if (std::filesystem::exists(fn) && sides[n].loadImage(fn.c_str())) break;
            }
            if (!sides[n].isOK()) return false;
            if (gamma != 1.0f) sides[n].decompressGamma(gamma);
            n++;
        }
    loaded = true;
    return true;
}
void CubemapEnvironment::fillProperties(ParsedBlock& pb)
{
    float gamma = 1.0f;
    pb.getFloatProp("assumedGamma", &gamma, 0.1f, 10.0f);
    std::string folder;
    if (!pb.getFilenameProp("folder", &folder)) pb.requiredProp("folder");
    // a cubemap that fails to load is only a warning: the environment renders black
    if (!loadMaps(folder, gamma)) fprintf(stderr, "CubemapEnvironment: Could not load maps from `%s'\n", folder.c_str());
}

void Node::fillProperties(ParsedBlock& pb)
{
    pb.getGeometryProp("geometry", &geom);
    pb.getShaderProp("shader", &shader);
    pb.getTransformProp(T);
    pb.getTextureProp("bump", &bump);
}

}  // namespace host
}  // namespace hxr
