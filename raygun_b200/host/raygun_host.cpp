// raygun_host.cpp -- see raygun_host.hpp.  Host logic only; the device work is behind include/rgb200.h.
#include "raygun_host.hpp"

#include <algorithm>
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

void RenderSystem::render(Scene& scene) {  // render_system.cpp:88-100 (then blit / ImGui / present: dropped)
    fillUniformBuffer(m_ubo, *scene.camera);
    m_raytracer->setupTopLevelAS(scene);
    m_raytracer->updateRenderTarget(m_ubo);
    m_raytracer->doRaytracing(useFXAA);
}

void RenderSystem::readFrame(std::vector<uint8_t>& rgba8) {
    rgba8.resize((size_t)m_width * m_height * 4);
    check(m_raytracer->ctx, rg_read_rgba8(m_raytracer->ctx, rgba8.data()), "rg_read_rgba8");
}
rg_timings RenderSystem::timings() {
    rg_timings t{};
    check(m_raytracer->ctx, rg_get_timings(m_raytracer->ctx, &t), "rg_get_timings");
    return t;
}

}  // namespace render

// ------------------------------------------------------------------------------------------------ ui::TextGenerator (ui/text.cpp:30-138)
namespace ui {
TextGenerator::TextGenerator(const Font& font, std::shared_ptr<Material> material, const RegisterModel& registerModel, float letterPadding_, float lineSpacing_)
    : m_charWidth(font.charWidth), letterPadding(letterPadding_), lineSpacing(lineSpacing_) {
    for(size_t k = 0; k < font.charMap.size(); ++k) {
        if(!font.charMap[k]) continue;
        auto model = std::make_shared<render::Model>();
        model->mesh = font.charMap[k];
        model->materials.push_back(material);
        if(registerModel) registerModel(model);
        m_charMap[k] = model;
    }
}
std::pair<std::shared_ptr<Entity>, render::Mesh::Bounds> TextGenerator::textInternal(string_view input) const {
    auto result = std::make_shared<Entity>("char_group_" + string(input));
    render::Mesh::Bounds bounds{};
    vec2 offset{};
    for(const char c: input) {
        if(c == ' ') {
            offset.x += 5 * letterPadding;
        } else if(c == '\n') {
            offset.x = 0;
            offset.y -= lineSpacing;
            continue;
        }
        const auto code = (unsigned char)c;
        if(code >= m_charMap.size()) continue;
        const auto& model = m_charMap[code];
        if(!model) continue;
        auto entity = result->emplaceChild(std::to_string((int)c));
        entity->move({offset.x, offset.y, 0.0f});
        entity->model = model;
        offset.x += letterPadding + m_charWidth[code];
        bounds.upper.x = offset.x - letterPadding;
        bounds.upper.y = offset.y + lineSpacing * 0.66f;
    }
    return {result, bounds};
}
std::pair<std::shared_ptr<Entity>, render::Mesh::Bounds> TextGenerator::textWithBounds(string_view input, Alignment align) const {
    auto [textEnt, bounds] = textInternal(input);
    vec3 offset(0.0f);
    switch(align) {
    case Alignment::TopLeft: break;
    case Alignment::TopCenter: offset = vec3(-bounds.upper.x / 2, 0, 0); break;
    case Alignment::TopRight: offset = vec3(-bounds.upper.x, 0, 0); break;
    case Alignment::MiddleLeft: offset = vec3(0, -bounds.upper.y / 2, 0); break;
    case Alignment::MiddleCenter: offset = vec3(-bounds.upper.x / 2, -bounds.upper.y / 2, 0); break;
    case Alignment::MiddleRight: offset = vec3(-bounds.upper.x, -bounds.upper.y / 2, 0); break;
    case Alignment::BottomLeft: offset = vec3(0, -bounds.upper.y, 0); break;
    case Alignment::BottomCenter: offset = vec3(-bounds.upper.x / 2, -bounds.upper.y, 0); break;
    case Alignment::BottomRight: offset = vec3(-bounds.upper.x, -bounds.upper.y, 0); break;
    }
    textEnt->moveTo(offset);
    bounds.upper += offset;
    bounds.lower += offset;
    auto result = std::make_shared<Entity>("string_" + string(input));
    result->addChild(textEnt);
    return {result, bounds};
}
}  // namespace ui

}  // namespace raygun
