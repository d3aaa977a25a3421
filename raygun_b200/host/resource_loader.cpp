// resource_loader.cpp -- asset ingestion for the ray-tracing path without Assimp / nlohmann-json:
//   * a Collada subset reader that reproduces what Assimp's importer hands raygun::Entity (entity.cpp:31-122) for the
//     reference's Blender-exported files: <triangles> groups with interleaved index tuples, one output vertex per tuple
//     (3 per triangle, no welding), one sub-mesh per group in file order merged per node, node <matrix> -> TRS,
//     material index = position in the file's material library sorted by id (Assimp keeps it in a std::map);
//   * the .rgmat.json reader of raygun/material.cpp:60-107 (keys = gpu_material.def names, "basedOn"),
//     with the underscore fallback of resource_manager.cpp:41-50.
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

#include "raygun_host.hpp"

namespace raygun {

namespace {

string readFile(const string& path) {
    std::ifstream f(path, std::ios::binary);
    if(!f) throw std::runtime_error("cannot open " + path);
    std::ostringstream ss; ss << f.rdbuf();
    return ss.str();
}
bool fileExists(const string& path) { std::ifstream f(path); return (bool)f; }

// ---------------------------------------------------------------------------------------------- minimal XML DOM
struct XmlNode {
    string tag, text;
    std::vector<std::pair<string, string>> attrs;
    std::vector<std::unique_ptr<XmlNode>> children;
    const string& attr(const string& k) const {
        static const string empty;
        for(auto& a: attrs) if(a.first == k) return a.second;
        return empty;
    }
    const XmlNode* child(const string& t) const { for(auto& c: children) if(c->tag == t) return c.get(); return nullptr; }
    void all(const string& t, std::vector<const XmlNode*>& out) const { for(auto& c: children) if(c->tag == t) out.push_back(c.get()); }
    void descendants(const string& t, std::vector<const XmlNode*>& out) const { for(auto& c: children) { if(c->tag == t) out.push_back(c.get()); c->descendants(t, out); } }
};

std::unique_ptr<XmlNode> parseXml(const string& s) {
    auto root = std::make_unique<XmlNode>();
    std::vector<XmlNode*> stack{root.get()};
    size_t i = 0;
    while(i < s.size()) {
        if(s[i] != '<') {
            const size_t j = s.find('<', i);
            stack.back()->text.append(s, i, (j == string::npos ? s.size() : j) - i);
            i = j == string::npos ? s.size() : j;
            continue;
        }
        if(s.compare(i, 4, "<!--") == 0) { i = s.find("-->", i) + 3; continue; }
        if(s[i + 1] == '?' || s[i + 1] == '!') { i = s.find('>', i) + 1; continue; }
        const size_t close = s.find('>', i);
        if(s[i + 1] == '/') { if(stack.size() > 1) stack.pop_back(); i = close + 1; continue; }
        const bool selfClosing = s[close - 1] == '/';
        const string inner = s.substr(i + 1, close - i - 1 - (selfClosing ? 1 : 0));
        auto node = std::make_unique<XmlNode>();
        size_t p = 0;
        while(p < inner.size() && !std::isspace((unsigned char)inner[p])) ++p;
        node->tag = inner.substr(0, p);
        while(p < inner.size()) {
            while(p < inner.size() && std::isspace((unsigned char)inner[p])) ++p;
            const size_t eq = inner.find('=', p);
            if(eq == string::npos) break;
            const string key = inner.substr(p, eq - p);
            const char q = inner[eq + 1];
            const size_t e = inner.find(q, eq + 2);
            node->attrs.emplace_back(key, inner.substr(eq + 2, e - eq - 2));
            p = e + 1;
        }
        XmlNode* raw = node.get();
        stack.back()->children.push_back(std::move(node));
        if(!selfClosing) stack.push_back(raw);
        i = close + 1;
    }
    return root;
}

template <class T, class Conv>
std::vector<T> parseNumbers(const string& text, Conv conv) {
    std::vector<T> out;
    const char* p = text.c_str();
    char* end = nullptr;
    while(true) {
        while(*p && std::isspace((unsigned char)*p)) ++p;
        if(!*p) break;
        out.push_back(conv(p, &end));
        if(end == p) break;
        p = end;
    }
    return out;
}
std::vector<float> parseFloats(const string& t) { return parseNumbers<float>(t, [](const char* p, char** e) { return std::strtof(p, e); }); }
std::vector<uint32_t> parseUints(const string& t) { return parseNumbers<uint32_t>(t, [](const char* p, char** e) { return (uint32_t)std::strtoul(p, e, 10); }); }

// ---------------------------------------------------------------------------------------------- minimal JSON
struct Json {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    double num = 0; bool b = false; string str;
    std::vector<Json> arr;
    std::vector<std::pair<string, Json>> obj;
};
struct JsonParser {
    const string& s; size_t i = 0;
    void ws() { while(i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    Json parse() {
        ws();
        Json j;
        if(i >= s.size()) throw std::runtime_error("json: unexpected end");
        const char c = s[i];
        if(c == '{') {
            j.kind = Json::Object; ++i; ws();
            if(s[i] == '}') { ++i; return j; }
            while(true) {
                ws(); Json k = parse(); ws();
                if(s[i] != ':') throw std::runtime_error("json: expected ':'");
                ++i;
                j.obj.emplace_back(k.str, parse()); ws();
                if(s[i] == ',') { ++i; continue; }
                if(s[i] == '}') { ++i; break; }
                throw std::runtime_error("json: expected ',' or '}'");
            }
        } else if(c == '[') {
            j.kind = Json::Array; ++i; ws();
            if(s[i] == ']') { ++i; return j; }
            while(true) {
                j.arr.push_back(parse()); ws();
                if(s[i] == ',') { ++i; continue; }
                if(s[i] == ']') { ++i; break; }
                throw std::runtime_error("json: expected ',' or ']'");
            }
        } else if(c == '"') {
            j.kind = Json::String; ++i;
            while(i < s.size() && s[i] != '"') { if(s[i] == '\\' && i + 1 < s.size()) ++i; j.str.push_back(s[i++]); }
            ++i;
        } else if(s.compare(i, 4, "true") == 0) { j.kind = Json::Bool; j.b = true; i += 4; }
        else if(s.compare(i, 5, "false") == 0) { j.kind = Json::Bool; j.b = false; i += 5; }
        else if(s.compare(i, 4, "null") == 0) { i += 4; }
        else { j.kind = Json::Number; char* e = nullptr; j.num = std::strtod(s.c_str() + i, &e); if(e == s.c_str() + i) throw std::runtime_error("json: bad token"); i = (size_t)(e - s.c_str()); }
        return j;
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------------ Material (material.cpp:60-107)
Material::Material(string_view name_, const string& path, const std::function<std::shared_ptr<Material>(const string&)>& loadBase)
    : name(name_), gpuMaterial(gpu::defaultMaterial()) {
    const string text = readFile(path);
    JsonParser p{text};
    const Json data = p.parse();
    if(data.kind != Json::Object) throw std::runtime_error("material " + path + ": not an object");
    for(const auto& kv: data.obj)
        if(kv.first == "basedOn" && loadBase) gpuMaterial = loadBase(kv.second.str)->gpuMaterial;
    auto vec = [&](const Json& v, float* dst) { for(size_t k = 0; k < 3 && k < v.arr.size(); ++k) dst[k] = (float)v.arr[k].num; };
    for(const auto& [key, value]: data.obj) {
        if(key == "type" || key == "basedOn" || key == "staticFriction" || key == "dynamicFriction") continue;
        if(key == "diffuse") vec(value, gpuMaterial.diffuse);
        else if(key == "specular") vec(value, gpuMaterial.specular);
        else if(key == "transparency") gpuMaterial.transparency = (float)value.num;
        else if(key == "reflectivity") gpuMaterial.reflectivity = (float)value.num;
        else if(key == "roughness") gpuMaterial.roughness = (float)value.num;
        else if(key == "ior") gpuMaterial.ior = (float)value.num;
        else if(key == "effectId") gpuMaterial.effect_id = (uint32_t)value.num;
        else if(key == "rayConsumption") gpuMaterial.ray_consumption = (uint32_t)value.num;
        else if(key == "emission") gpuMaterial.emission = (float)value.num;
        // unknown fields are warnings in the reference (material.cpp:104)
    }
}

std::shared_ptr<Material> ResourceManager::loadMaterial(const string& name) {
    for(auto& m: m_materials) if(m.first == name) return m.second;
    string path = m_dir + "/materials/" + name + ".rgmat.json";
    if(!fileExists(path)) {  // e.g. ui_button -> ui/button (resource_manager.cpp:41-50)
        const size_t us = name.find('_');
        if(us != string::npos) path = m_dir + "/materials/" + name.substr(0, us) + "/" + name.substr(us + 1) + ".rgmat.json";
    }
    auto mat = std::make_shared<Material>(name, path, [this](const string& base) { return loadMaterial(base); });
    m_materials.emplace_back(name, mat);
    return mat;
}

// ------------------------------------------------------------------------------------------------ Collada -> Entity tree
std::shared_ptr<Entity> ResourceManager::loadEntity(string_view name) {
    const string path = m_dir + "/models/" + string(name) + ".dae";
    const auto doc = parseXml(readFile(path));
    const XmlNode* collada = doc->child("COLLADA");
    if(!collada) throw std::runtime_error(path + ": not a COLLADA document");

    // material library, in Assimp's order (std::map keyed by material id)
    std::map<string, string> matById;
    if(const XmlNode* lib = collada->child("library_materials")) {
        std::vector<const XmlNode*> mats; lib->all("material", mats);
        for(auto m: mats) matById[m->attr("id")] = m->attr("name");
    }
    std::vector<string> matIds;
    std::vector<std::shared_ptr<Material>> materials;
    for(auto& kv: matById) { matIds.push_back(kv.first); materials.push_back(loadMaterial(kv.second)); }

    // geometries: per <triangles> group the expanded vertex stream
    struct Group { string materialSymbol; std::vector<render::Vertex> vertices; };
    std::map<string, std::vector<Group>> geometries;
    if(const XmlNode* lib = collada->child("library_geometries")) {
        std::vector<const XmlNode*> geoms; lib->all("geometry", geoms);
        for(auto g: geoms) {
            const XmlNode* mesh = g->child("mesh");
            if(!mesh) continue;
            std::map<string, std::pair<std::vector<float>, uint32_t>> sources;  // id -> (floats, stride)
            std::vector<const XmlNode*> srcs; mesh->all("source", srcs);
            for(auto s: srcs) {
                const XmlNode* fa = s->child("float_array");
                uint32_t stride = 3;
                if(const XmlNode* tc = s->child("technique_common")) if(const XmlNode* acc = tc->child("accessor")) stride = (uint32_t)std::atoi(acc->attr("stride").c_str());
                sources["#" + s->attr("id")] = {fa ? parseFloats(fa->text) : std::vector<float>{}, stride};
            }
            std::map<string, string> vertexPositions;  // <vertices id> -> POSITION source
            std::vector<const XmlNode*> verts; mesh->all("vertices", verts);
            for(auto v: verts) {
                std::vector<const XmlNode*> inputs; v->all("input", inputs);
                for(auto in: inputs) if(in->attr("semantic") == "POSITION") vertexPositions["#" + v->attr("id")] = in->attr("source");
            }
            std::vector<const XmlNode*> tris; mesh->all("triangles", tris);
            std::vector<Group> groups;
            for(auto t: tris) {
                uint32_t stride = 0, offVertex = 0, offNormal = 0;
                string srcVertex, srcNormal;
                std::vector<const XmlNode*> inputs; t->all("input", inputs);
                for(auto in: inputs) {
                    const uint32_t off = (uint32_t)std::atoi(in->attr("offset").c_str());
                    stride = std::max(stride, off + 1);
                    if(in->attr("semantic") == "VERTEX") { offVertex = off; srcVertex = vertexPositions[in->attr("source")]; }
                    if(in->attr("semantic") == "NORMAL") { offNormal = off; srcNormal = in->attr("source"); }
                }
                const XmlNode* p = t->child("p");
                const std::vector<uint32_t> idx = p ? parseUints(p->text) : std::vector<uint32_t>{};
                const auto& pos = sources[srcVertex]; const auto& nrm = sources[srcNormal];
                Group grp; grp.materialSymbol = t->attr("material");
                const size_t n = idx.size() / stride;
                grp.vertices.resize(n);
                for(size_t k = 0; k < n; ++k) {
                    render::Vertex v{};
                    const uint32_t ip = idx[k * stride + offVertex], in = idx[k * stride + offNormal];
                    for(int a = 0; a < 3; ++a) { v.position[a] = pos.first[(size_t)ip * pos.second + a]; v.normal[a] = nrm.first.empty() ? 0.0f : nrm.first[(size_t)in * nrm.second + a]; }
                    grp.vertices[k] = v;
                }
                groups.push_back(std::move(grp));
            }
            geometries["#" + g->attr("id")] = std::move(groups);
        }
    }

    auto entity = std::make_shared<Entity>(name);
    const XmlNode* lvs = collada->child("library_visual_scenes");
    const XmlNode* vs = lvs ? lvs->child("visual_scene") : nullptr;
    if(!vs) return entity;
    std::vector<const XmlNode*> nodes; vs->all("node", nodes);
    for(auto n: nodes) {  // aiscene->mRootNode->mChildren, entity.cpp:107-119
        const XmlNode* ig = n->child("instance_geometry");
        if(!ig) continue;
        std::map<string, string> bind;  // symbol -> material id
        std::vector<const XmlNode*> ims; ig->descendants("instance_material", ims);
        for(auto im: ims) bind[im->attr("symbol")] = im->attr("target").substr(1);
        auto mesh = std::make_shared<render::Mesh>();
        for(const Group& grp: geometries[ig->attr("url")]) {  // loadMesh + merge, entity.cpp:31-81
            render::Mesh sub;
            uint32_t matIndex = 0;
            const string matId = bind.count(grp.materialSymbol) ? bind[grp.materialSymbol] : grp.materialSymbol;
            for(size_t k = 0; k < matIds.size(); ++k) if(matIds[k] == matId) matIndex = (uint32_t)k;
            sub.vertices = grp.vertices;
            for(auto& v: sub.vertices) v.mat_index = matIndex;
            sub.indices.resize(sub.vertices.size());
            for(size_t k = 0; k < sub.indices.size(); ++k) sub.indices[k] = (uint32_t)k;
            mesh->merge(sub);
        }
        auto model = std::make_shared<render::Model>();
        model->mesh = mesh;
        model->materials = materials;
        registerModel(model);
        auto child = entity->emplaceChild(n->attr("name"));
        if(const XmlNode* mx = n->child("matrix")) {
            const std::vector<float> f = parseFloats(mx->text);  // row-major in the file
            mat4 m;
            if(f.size() == 16) for(int r = 0; r < 4; ++r) for(int c = 0; c < 4; ++c) m[c][r] = f[(size_t)r * 4 + c];
            child->setTransform(Transform{m});  // utils::toTransform, assimp_utils.hpp:29-33
        }
        child->model = model;
    }
    return entity;
}


// ---------------------------------------------------------------------------------------------- fonts
// ResourceManager::loadFont (resource_manager.cpp:107-135): resources/fonts/<name>.obj holds one object per glyph, named by its
// ASCII code ("o 65").  What Assimp's OBJ importer hands raygun::Entity for such a file (aiProcess_Triangulate only, entity.cpp:
// 88-90): one node per object, one output vertex per face corner in face order (no welding), position + normal, default material
// index 0.  The glyph mesh is then shifted so that its left edge is x = 0 and its width is recorded.
std::shared_ptr<ui::Font> ResourceManager::loadFont(string_view nameView) {
    const string name(nameView);
    for(auto& f: m_fonts) if(f.first == name) return f.second;
    const string text = readFile(m_dir + "/fonts/" + name + ".obj");
    auto font = std::make_shared<ui::Font>();
    font->name = name;
    std::vector<vec3> pos, nrm;
    std::shared_ptr<render::Mesh> mesh;
    long glyph = -1;
    auto finish = [&] {
        if(!mesh || glyph < 0 || (size_t)glyph >= font->charMap.size() || mesh->vertices.empty()) return;
        const auto b = mesh->bounds();
        for(auto& v: mesh->vertices) v.position[0] -= b.lower.x;
        font->charMap[(size_t)glyph] = mesh;
        font->charWidth[(size_t)glyph] = mesh->width();
    };
    size_t i = 0;
    while(i < text.size()) {
        size_t e = text.find('\n', i);
        if(e == string::npos) e = text.size();
        const char* l = text.c_str() + i;
        if(l[0] == 'o' && l[1] == ' ') {
            finish();
            mesh = std::make_shared<render::Mesh>();
            char* end = nullptr;
            glyph = std::strtol(l + 2, &end, 10);
            if(end == l + 2) glyph = -1;   // not a number: std::stoul would throw in the reference; skipped here
        } else if(l[0] == 'v' && (l[1] == ' ' || (l[1] == 'n' && l[2] == ' '))) {
            char* q = const_cast<char*>(l + (l[1] == 'n' ? 3 : 2));
            vec3 v;
            v.x = (float)std::strtod(q, &q); v.y = (float)std::strtod(q, &q); v.z = (float)std::strtod(q, &q);
            (l[1] == 'n' ? nrm : pos).push_back(v);
        } else if(l[0] == 'f' && l[1] == ' ' && mesh) {
            // "f v//vn v//vn v//vn" (also v, v/vt, v/vt/vn); polygons are fan-triangulated like aiProcess_Triangulate does for convex faces
            std::vector<std::pair<long, long>> corners;
            char* q = const_cast<char*>(l + 2);
            const char* lineEnd = text.c_str() + e;
            while(q < lineEnd) {
                while(q < lineEnd && (*q == ' ' || *q == '\r')) ++q;
                if(q >= lineEnd) break;
                char* n0 = q;
                const long vi = std::strtol(q, &q, 10);
                if(q == n0) break;
                long ni = 0;
                if(*q == '/') { ++q; if(*q != '/') std::strtol(q, &q, 10); if(*q == '/') { ++q; ni = std::strtol(q, &q, 10); } }
                corners.push_back({vi, ni});
            }
            auto emit = [&](const std::pair<long, long>& c) {
                render::Vertex v{};
                const long vi = c.first > 0 ? c.first - 1 : (long)pos.size() + c.first, ni = c.second > 0 ? c.second - 1 : (long)nrm.size() + c.second;
                if(vi >= 0 && (size_t)vi < pos.size()) { v.position[0] = pos[(size_t)vi].x; v.position[1] = pos[(size_t)vi].y; v.position[2] = pos[(size_t)vi].z; }
                if(c.second != 0 && ni >= 0 && (size_t)ni < nrm.size()) { v.normal[0] = nrm[(size_t)ni].x; v.normal[1] = nrm[(size_t)ni].y; v.normal[2] = nrm[(size_t)ni].z; }
                v.mat_index = 0;
                mesh->indices.push_back((uint32_t)mesh->vertices.size());
                mesh->vertices.push_back(v);
            };
            for(size_t k = 2; k < corners.size(); ++k) { emit(corners[0]); emit(corners[k - 1]); emit(corners[k]); }
        }
        i = e + 1;
    }
    finish();
    m_fonts.push_back({name, font});
    return font;
}

}  // namespace raygun
