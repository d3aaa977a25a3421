// raygun_host.cpp -- see raygun_host.hpp.  Host logic only; the device work is behind include/rgb200.h.
#include "raygun_host.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <set>
#include <stdexcept>

namespace raygun {

namespace gpu {
Material defaultMaterial() {  // gpu_material.def:11-26
    Material m{};
    m.diffuse[0] = 1.0f; m.diffuse[1] = 0.0f; m.diffuse[2] = 1.0f; m.transparency = 0.f;
    m.specular[0] = m.specular[1] = m.specular[2] = 1.f; m.reflectivity = 0.f;
    m.roughness = 0.f; m.ior = 1.f; m.effect_id = 0; m.ray_consumption = 1; m.emission = 0.f;
    return m;
}
}  // namespace gpu

namespace render {

vec3 Mesh::center() const {  // mesh.cpp:27-31
    vec3 sum(0.0f);
    for(auto index: indices) sum += vec3(vertices[index].position[0], vertices[index].position[1], vertices[index].position[2]);
    return sum * (1.0f / (float)indices.size());
}
Mesh::Bounds Mesh::bounds() const {  // mesh.cpp:33-44
    vec3 lower(std::numeric_limits<float>::max()), upper(std::numeric_limits<float>::lowest());
    for(const auto& v: vertices)
        for(int a = 0; a < 3; ++a) { lower[a] = std::min(v.position[a], lower[a]); upper[a] = std::max(v.position[a], upper[a]); }
    return {lower, upper};
}
float Mesh::width() const { const auto b = bounds(); return b.upper.x - b.lower.x; }
void Mesh::merge(const Mesh& other) {  // mesh.cpp:52-62: indices shift by the vertices already present
    const auto indexOffset = (uint32_t)vertices.size();
    indices.reserve(indices.size() + other.indices.size());
    for(auto i: other.indices) indices.push_back(i + indexOffset);
    vertices.insert(vertices.end(), other.vertices.begin(), other.vertices.end());
}
void Mesh::forEachFace(std::function<void(const Vertex&, const Vertex&, const Vertex&)> action) const {
    for(size_t i = 0; i + 2 < indices.size(); i += 3) action(vertices[indices[i]], vertices[indices[i + 1]], vertices[indices[i + 2]]);
}

}  // namespace render

// ------------------------------------------------------------------------------------------------ Entity (entity.cpp:124-260)
void Entity::addChild(std::shared_ptr<Entity> child) {
    if(child->m_parent) throw std::logic_error("Entity::addChild: child already has a parent");
    child->setParent(this);
    m_children.push_back(child);
}
std::shared_ptr<Entity> Entity::emplaceChild(string_view childName) {
    auto child = std::make_shared<Entity>(childName);
    addChild(child);
    return child;
}
void Entity::removeChild(const std::shared_ptr<Entity>& child) {
    auto it = std::find(m_children.begin(), m_children.end(), child);
    if(it == m_children.end()) return;
    (*it)->setParent(nullptr);
    m_children.erase(it);
}
void Entity::clearChildren() {
    for(auto& c: m_children) c->setParent(nullptr);
    m_children.clear();
}
Transform Entity::parentTransform() const {  // entity.cpp:187-194
    if(!m_cachedParentTransform) m_cachedParentTransform = m_parent ? m_parent->globalTransform() : Transform{};
    return *m_cachedParentTransform;
}
void Entity::invalidateCachedParentTransform() {
    m_cachedParentTransform.reset();
    invalidateChildrenCachedParentTransform();
}
void Entity::invalidateChildrenCachedParentTransform() {
    for(auto& c: m_children) c->invalidateCachedParentTransform();
}

// ------------------------------------------------------------------------------------------------ Camera (camera.cpp:34-47)
void Camera::updateProjection(uint32_t width, uint32_t height) {
    auto aspectRatio = (float)width / (float)height;
    if(!std::isfinite(aspectRatio)) aspectRatio = 16.0f / 9.0f;
    m_projection = perspectiveRH_ZO(FOV * 0.01745329251994329576923690768489f, aspectRatio, NEAR, FAR);
    m_projection[1][1] *= -1;  // GLM's flipped Y
}

namespace render {

// ------------------------------------------------------------------------------------------------ Raytracer
static void check(rg_ctx* ctx, int rc, const char* what) {
    if(rc) throw std::runtime_error(string("rgb200: ") + what + ": " + (ctx ? rg_last_error(ctx) : "context creation failed (no CUDA device; there is no CPU fallback)"));
}
Raytracer::Raytracer(uint32_t width, uint32_t height, int device) { check(nullptr, rg_create(&ctx, device, width, height), "rg_create"); }
Raytracer::~Raytracer() { rg_destroy(ctx); }
void Raytracer::setupBottomLevelAS() { check(ctx, rg_build_blas(ctx), "rg_build_blas"); }

// TopLevelAS::TopLevelAS, acceleration_structure.cpp:55-85: DFS pre-order, prune invisible / zero-volume subtrees,
// instance i = {transpose(globalTransform().toMat4()) as 3x4 row-major, customIndex i} + its offset-table entry.
void Raytracer::gatherInstances(const Scene& scene, std::vector<rg_instance>& out) {
    out.clear();
    scene.root->forEachEntity([&](Entity& entity) {
        if(!entity.isVisible()) return false;
        if(entity.transform().isZeroVolume()) return false;
        if(!entity.model) return true;
        rg_instance in{};
        const mat4 t = transpose(entity.globalTransform().toMat4());
        std::memcpy(in.xform, &t.m[0][0], sizeof in.xform);
        in.mesh = entity.model->mesh->meshIndex;
        in.vtx_off = entity.model->mesh->vertexBufferRef.offsetInElements();
        in.idx_off = entity.model->mesh->indexBufferRef.offsetInElements();
        in.mat_off = entity.model->materialBufferRef.offsetInElements();
        out.push_back(in);
        return true;
    });
}
// The same walk without the transform composition: every entity (pruning is the device's job too) as its LOCAL TRS + parent index,
// in the order forEachEntity visits them.  rg_set_entities composes globalTransform and the 3x4 on the GPU (bit-identical records).
void Raytracer::gatherEntities(const Scene& scene, std::vector<rg_entity>& out) {
    out.clear();
    std::vector<std::pair<const Entity*, int32_t>> stack;   // (entity, index of its parent in `out`)
    stack.push_back({scene.root.get(), -1});
    while(!stack.empty()) {
        const auto [entity, parent] = stack.back();
        stack.pop_back();
        rg_entity e{};
        const Transform& t = entity->transform();
        e.position[0] = t.position.x; e.position[1] = t.position.y; e.position[2] = t.position.z;
        e.rotation[0] = t.rotation.w; e.rotation[1] = t.rotation.x; e.rotation[2] = t.rotation.y; e.rotation[3] = t.rotation.z;
        e.scaling[0] = t.scaling.x; e.scaling[1] = t.scaling.y; e.scaling[2] = t.scaling.z;
        e.parent = parent;
        e.flags = (entity->isVisible() ? RG_ENTITY_VISIBLE : 0u) | (entity->model ? RG_ENTITY_HAS_MODEL : 0u);
        if(entity->model) {
            e.mesh = entity->model->mesh->meshIndex;
            e.vtx_off = entity->model->mesh->vertexBufferRef.offsetInElements();
            e.idx_off = entity->model->mesh->indexBufferRef.offsetInElements();
            e.mat_off = entity->model->materialBufferRef.offsetInElements();
        }
        const int32_t self = (int32_t)out.size();
        out.push_back(e);
        const auto& ch = entity->children();
        for(auto it = ch.rbegin(); it != ch.rend(); ++it) stack.push_back({it->get(), self});   // reversed: pre-order pops the first child first
    }
}
void Raytracer::setupTopLevelAS(const Scene& scene) {
    if(deviceSceneWalk) {
        gatherEntities(scene, entities);
        uint32_t n = 0;
        check(ctx, rg_set_entities(ctx, entities.data(), (uint32_t)entities.size(), &n), "rg_set_entities");
        return;
    }
    gatherInstances(scene, instances);
    check(ctx, rg_set_instances(ctx, instances.data(), (uint32_t)instances.size()), "rg_set_instances");
}
void Raytracer::updateRenderTarget(const gpu::UniformBufferObject& ubo) { check(ctx, rg_set_ubo(ctx, &ubo), "rg_set_ubo"); }
void Raytracer::doRaytracing(bool useFXAA) { check(ctx, rg_render(ctx, useFXAA ? RG_FXAA : 0u), "rg_render"); }

// ------------------------------------------------------------------------------------------------ RenderSystem
RenderSystem::RenderSystem(uint32_t width, uint32_t height, int device) : m_width(width), m_height(height) {
    // resetUniformBuffer, render_system.cpp:235-244
    std::memset(&m_ubo, 0, sizeof m_ubo);
    const vec3 l = normalize(vec3(.4f, -.6f, -.8f));
    m_ubo.light_dir[0] = l.x; m_ubo.light_dir[1] = l.y; m_ubo.light_dir[2] = l.z;
    m_ubo.num_samples = 1;
    m_ubo.max_recursions = 5;
    m_raytracer = std::make_unique<Raytracer>(width, height, device);
    clock = [t0 = std::chrono::steady_clock::now()] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
}

// One vertex / index buffer over the DISTINCT meshes and one material buffer over all models' material lists;
// BufferRef offsets recorded on the meshes / models (render_system.cpp:270-330).  The reference iterates a
// std::set<Mesh*> (pointer order); registration order is used here -- the order is not observable in the image.
void RenderSystem::packModelBuffers(const std::vector<std::shared_ptr<Model>>& models, std::vector<Vertex>& v, std::vector<uint32_t>& i,
                                    std::vector<gpu::Material>& m, std::vector<rg_mesh_range>& ranges) {
    v.clear(); i.clear(); m.clear(); ranges.clear();
    std::set<Mesh*> seen;
    for(const auto& model: models) {
        Mesh* mesh = model->mesh.get();
        if(seen.insert(mesh).second) {
            mesh->meshIndex = (uint32_t)ranges.size();
            mesh->vertexBufferRef = {(uint32_t)(v.size() * sizeof(Vertex)), (uint32_t)(mesh->vertices.size() * sizeof(Vertex)), sizeof(Vertex)};
            mesh->indexBufferRef = {(uint32_t)(i.size() * sizeof(uint32_t)), (uint32_t)(mesh->indices.size() * sizeof(uint32_t)), sizeof(uint32_t)};
            ranges.push_back({(uint32_t)v.size(), (uint32_t)mesh->vertices.size(), (uint32_t)i.size(), (uint32_t)mesh->indices.size()});
            v.insert(v.end(), mesh->vertices.begin(), mesh->vertices.end());
            i.insert(i.end(), mesh->indices.begin(), mesh->indices.end());
        }
        model->materialBufferRef = {(uint32_t)(m.size() * sizeof(gpu::Material)), (uint32_t)(model->materials.size() * sizeof(gpu::Material)), sizeof(gpu::Material)};
        for(const auto& mat: model->materials) m.push_back(mat->gpuMaterial);
    }
}

void RenderSystem::setupModelBuffers(const std::vector<std::shared_ptr<Model>>& models) {
    packModelBuffers(models, vertexBuffer, indexBuffer, materialBuffer, meshRanges);
    rg_ctx* ctx = m_raytracer->ctx;
    check(ctx, rg_upload_geometry(ctx, vertexBuffer.data(), (uint32_t)vertexBuffer.size(), indexBuffer.data(), (uint32_t)indexBuffer.size(), meshRanges.data(),
                                  (uint32_t)meshRanges.size()), "rg_upload_geometry");
    check(ctx, rg_upload_materials(ctx, materialBuffer.data(), (uint32_t)materialBuffer.size()), "rg_upload_materials");
}

void RenderSystem::fillUniformBuffer(gpu::UniformBufferObject& ubo, const Camera& camera) {  // render_system.cpp:246-252
    const mat4 vi = camera.viewInverse(), pi = camera.projInverse();
    std::memcpy(ubo.view_inverse, &vi.m[0][0], 64);
    std::memcpy(ubo.proj_inverse, &pi.m[0][0], 64);
    ubo.clear_color[0] = ubo.clear_color[1] = ubo.clear_color[2] = 0.2f;
}

void RenderSystem::render(Scene& scene) {  // render_system.cpp:88-100 (then blit / ImGui / present: replaced by readFrame / writeFrame*)
    fillUniformBuffer(m_ubo, *scene.camera);
    m_ubo.time = (float)std::fmod(clock(), 32.0 * 3.14159265358979323846);   // render_system.cpp:253-255
    if(m_currentFade) {                                                      // render_system.cpp:257-259
        const vec4 c = m_currentFade->curColor();
        m_ubo.fade_color[0] = c.x; m_ubo.fade_color[1] = c.y; m_ubo.fade_color[2] = c.z; m_ubo.fade_color[3] = c.w;
    }
    m_raytracer->setupTopLevelAS(scene);
    m_raytracer->updateRenderTarget(m_ubo);
    m_raytracer->doRaytracing(useFXAA);
}

void RenderSystem::readFrame(std::vector<uint8_t>& rgba8) {
    rgba8.resize((size_t)m_width * m_height * 4);
    check(m_raytracer->ctx, rg_read_rgba8(m_raytracer->ctx, rgba8.data()), "rg_read_rgba8");
}
// The reference presents the frame through the swapchain (render_system.cpp:130-159); headless, the frame goes to a file.
void RenderSystem::writeFramePPM(const string& path) { std::vector<uint8_t> rgba; readFrame(rgba); writeImagePPM(path, rgba.data(), m_width, m_height); }
void RenderSystem::writeFramePNG(const string& path) { std::vector<uint8_t> rgba; readFrame(rgba); writeImagePNG(path, rgba.data(), m_width, m_height); }
void RenderSystem::writeImagePPM(const string& path, const uint8_t* rgba, uint32_t m_width, uint32_t m_height) {   // binary PPM (P6): RGB, alpha dropped
    FILE* f = std::fopen(path.c_str(), "wb");
    if(!f) throw std::runtime_error("writeFramePPM: cannot open " + path);
    std::fprintf(f, "P6\n%u %u\n255\n", m_width, m_height);
    std::vector<uint8_t> row((size_t)m_width * 3);
    for(uint32_t y = 0; y < m_height; ++y) {
        for(uint32_t x = 0; x < m_width; ++x) std::memcpy(&row[3 * (size_t)x], &rgba[4 * ((size_t)y * m_width + x)], 3);
        std::fwrite(row.data(), 1, row.size(), f);
    }
    std::fclose(f);
}
namespace {
uint32_t crc32Update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool ready = false;
    if(!ready) {
        for(uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for(int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        ready = true;
    }
    for(size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return crc;
}
void pngChunk(FILE* f, const char type[4], const std::vector<uint8_t>& data) {
    const uint32_t n = (uint32_t)data.size();
    const uint8_t len[4] = {uint8_t(n >> 24), uint8_t(n >> 16), uint8_t(n >> 8), uint8_t(n)};
    std::fwrite(len, 1, 4, f);
    std::fwrite(type, 1, 4, f);
    if(n) std::fwrite(data.data(), 1, n, f);
    uint32_t crc = crc32Update(0xffffffffu, reinterpret_cast<const uint8_t*>(type), 4);
    crc = crc32Update(crc, data.data(), n) ^ 0xffffffffu;
    const uint8_t c[4] = {uint8_t(crc >> 24), uint8_t(crc >> 16), uint8_t(crc >> 8), uint8_t(crc)};
    std::fwrite(c, 1, 4, f);
}
}  // namespace
void RenderSystem::writeImagePNG(const string& path, const uint8_t* rgba, uint32_t m_width, uint32_t m_height) {   // 8-bit RGBA, stored (uncompressed) deflate blocks: no zlib needed
    std::vector<uint8_t> raw;
    raw.reserve(((size_t)m_width * 4 + 1) * m_height);
    for(uint32_t y = 0; y < m_height; ++y) {
        raw.push_back(0);   // filter type none
        raw.insert(raw.end(), rgba + 4 * (size_t)y * m_width, rgba + 4 * (size_t)(y + 1) * m_width);
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;   // Adler-32 of the raw stream
    for(size_t off = 0; off < raw.size() || off == 0; off += 65535) {
        const size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n >= raw.size() ? 1 : 0);
        z.push_back(uint8_t(n)); z.push_back(uint8_t(n >> 8)); z.push_back(uint8_t(~n)); z.push_back(uint8_t((~n) >> 8));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        for(size_t i = 0; i < n; ++i) { a = (a + raw[off + i]) % 65521u; b = (b + a) % 65521u; }
        if(raw.empty()) break;
    }
    const uint32_t adler = (b << 16) | a;
    z.push_back(uint8_t(adler >> 24)); z.push_back(uint8_t(adler >> 16)); z.push_back(uint8_t(adler >> 8)); z.push_back(uint8_t(adler));
    FILE* f = std::fopen(path.c_str(), "wb");
    if(!f) throw std::runtime_error("writeFramePNG: cannot open " + path);
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::fwrite(sig, 1, 8, f);
    std::vector<uint8_t> hdr = {uint8_t(m_width >> 24), uint8_t(m_width >> 16), uint8_t(m_width >> 8), uint8_t(m_width),
                                uint8_t(m_height >> 24), uint8_t(m_height >> 16), uint8_t(m_height >> 8), uint8_t(m_height), 8, 6, 0, 0, 0};
    pngChunk(f, "IHDR", hdr);
    pngChunk(f, "IDAT", z);
    pngChunk(f, "IEND", {});
    std::fclose(f);
}
rg_timings RenderSystem::timings() {
    rg_timings t{};
    check(m_raytracer->ctx, rg_get_timings(m_raytracer->ctx, &t), "rg_get_timings");
    return t;
}

}  // namespace render

// ------------------------------------------------------------------------------------------------ ui::TextGenerator
// Text becomes ordinary ray-traced geometry: one instance per glyph (raygun/ui/text.hpp:39-66 is the interface kept; what a caller of
// ui/text.cpp:53-138 observes is the set of glyph instances, their pen positions and the bounds).  Here the layout is ONE flat pass
// over the string that yields glyph placements -- the form the instance upload wants (rg_entity / rg_instance records) -- and the
// entity tree is built from that list afterwards.
namespace ui {
TextGenerator::TextGenerator(const Font& font, std::shared_ptr<Material> material, const RegisterModel& registerModel, float letterPadding_, float lineSpacing_)
    : m_charWidth(font.charWidth), letterPadding(letterPadding_), lineSpacing(lineSpacing_) {
    for(size_t code = 0; code < font.charMap.size(); ++code) {
        if(!font.charMap[code]) continue;
        auto glyph = std::make_shared<render::Model>();
        glyph->mesh = font.charMap[code];
        glyph->materials = {material};
        if(registerModel) registerModel(glyph);
        m_charMap[code] = std::move(glyph);
    }
}

// Layout rules (numerically those of the reference, so that glyph instances land on the same binary32 positions):
//   pen starts at (0, 0); a glyph is placed at the pen, then the pen advances by letterPadding + its width;
//   ' ' advances the pen by 5 letterPaddings and places nothing; a line feed returns the pen to x = 0, one lineSpacing down;
//   code points without a glyph are skipped; the extent is (pen.x - letterPadding, pen.y + 0.66 lineSpacing) after the LAST glyph.
TextGenerator::Layout TextGenerator::layout(string_view input) const {
    Layout out;
    float penX = 0.0f, penY = 0.0f;
    for(const char ch: input) {
        const auto code = (unsigned char)ch;
        if(ch == '\n') { penX = 0.0f; penY -= lineSpacing; continue; }
        if(ch == ' ') penX += 5 * letterPadding;
        if(code >= m_charMap.size() || !m_charMap[code]) continue;
        out.glyphs.push_back({code, penX, penY});
        penX += letterPadding + m_charWidth[code];
        out.extent = vec2{penX - letterPadding, penY + lineSpacing * 0.66f};
    }
    return out;
}

// Alignment = (row, column) of a 3 x 3 anchor grid: the block moves left by 0, 1/2 or 1 extent and down by 0, 1/2 or 1 extent.
vec3 TextGenerator::anchorOffset(Alignment align, vec2 extent) {
    const int column = (int)align % 3, row = (int)align / 3;
    const float dx = column == 0 ? 0.0f : (column == 1 ? -extent.x / 2 : -extent.x);
    const float dy = row == 0 ? 0.0f : (row == 1 ? -extent.y / 2 : -extent.y);
    return vec3(dx, dy, 0.0f);
}

std::pair<std::shared_ptr<Entity>, render::Mesh::Bounds> TextGenerator::textWithBounds(string_view input, Alignment align) const {
    const Layout placed = layout(input);
    const vec3 anchor = anchorOffset(align, placed.extent);
    // two levels, as callers of the reference see them: the returned entity is free to be moved by the caller, its single child
    // carries the anchor offset and the glyph instances hang below it (global position = (caller + anchor) + pen, in this order)
    auto block = std::make_shared<Entity>("text:" + string(input));
    auto line = block->emplaceChild("glyphs");
    line->moveTo(anchor);
    for(const Placement& g: placed.glyphs) {
        auto e = line->emplaceChild(string(1, (char)g.code));
        e->moveTo({g.x, g.y, 0.0f});
        e->model = m_charMap[g.code];
    }
    render::Mesh::Bounds bounds{};
    bounds.lower = anchor;
    bounds.upper = vec3(placed.extent.x, placed.extent.y, 0.0f) + anchor;
    return {block, bounds};
}
}  // namespace ui

// ------------------------------------------------------------------------------------------------ render::Fade (fade.hpp:27-63)
namespace render {
// alpha envelopes over progress = elapsed / duration (fade.cpp:47-52, :71-88)
vec4 Fade::curColor() { return vec4{0.0f, 0.0f, 0.0f, 0.0f}; }
bool Fade::over() const { return true; }

FadeIn::FadeIn(const Clock& clock, double duration, vec3 fromColor) : Fade(clock), m_duration(duration), m_color(fromColor) {}
vec4 FadeIn::curColor() {
    const double progress = std::clamp(elapsed() / m_duration, 0.0, 1.0);
    return vec4{m_color.x, m_color.y, m_color.z, float(1 - progress)};
}
bool FadeIn::over() const { return elapsed() > m_duration; }

FadeTransition::FadeTransition(const Clock& clock, double halfDuration, std::function<void()> atPeak, vec3 color)
    : Fade(clock), m_half(halfDuration), m_atPeak(std::move(atPeak)), m_color(color) {}
vec4 FadeTransition::curColor() {
    const double progress = elapsed() / m_half;
    if(!m_switched) {                       // rising edge: opaque at progress 1, where the scene is switched
        m_alpha = float(std::clamp(progress, 0.0, 1.0));
        if(progress >= 1) { m_switched = true; if(m_atPeak) m_atPeak(); }
    } else {
        m_alpha = float(std::clamp(2 - progress, 0.0, 1.0));
    }
    return vec4{m_color.x, m_color.y, m_color.z, m_alpha};
}
bool FadeTransition::over() const { return elapsed() > 2 * m_half; }
}  // namespace render

}  // namespace raygun
