// raygun_host.hpp -- host-side mirror of the reference's scene API for the ray-tracing path, above the C ABI.
// Names and semantics follow the reference so that application code written against it carries over:
//   Transform            raygun/transform.hpp:27-118
//   Mesh / Model         raygun/render/mesh.hpp:31-54, mesh.cpp:52-62; raygun/render/model.hpp:33-42
//   Material             raygun/material.hpp:31-40, material.cpp:60-107 (rgmat JSON, "basedOn")
//   Entity               raygun/entity.hpp:32-131, entity.cpp:60-122, :187-199
//   Camera               raygun/camera.hpp:29-46, camera.cpp:34-47
//   Scene                raygun/scene.hpp:36-52, scene.cpp:31-32
//   Raytracer            raygun/render/raytracer.hpp:36-101   (thin wrapper over include/rgb200.h)
//   RenderSystem         raygun/render/render_system.hpp:39-95, render_system.cpp:88-162, :192-330 (headless: no swapchain / ImGui)
// Physics, audio, window, Vulkan objects are out of scope and absent.
#pragma once
#include <array>
#include <functional>
#include <memory>
#include <optional>
#include <string>
#include <string_view>
#include <type_traits>
#include <vector>

#include "../../include/rgb200.h"
#include "rg_math.hpp"

namespace raygun {

using string = std::string;
using string_view = std::string_view;

struct Transform {
    Transform() {}
    explicit Transform(const vec3& position_) : position(position_) {}
    explicit Transform(const mat4& mat) { decompose(mat, scaling, rotation, position); }

    mat4 toMat4() const { return translate(position) * raygun::toMat4(rotation) * raygun::scale(scaling); }
    vec3 up() const { return raygun::rotate(rotation, UP); }
    vec3 right() const { return raygun::rotate(rotation, RIGHT); }
    vec3 forward() const { return raygun::rotate(rotation, FORWARD); }
    Transform inverse() const { Transform r; r.position = -position; r.rotation = raygun::inverse(rotation); r.scaling = 1.0f / scaling; return r; }
    bool isIdentity() const { return position == vec3(0.0f) && rotation == quat{} && scaling == vec3(1.0f); }
    bool isZeroVolume() const { return scaling.x * scaling.y * scaling.z == 0.0f; }

    void move(const vec3& translation) { position += translation; }
    void rotate(float angle, vec3 axis) { rotation = raygun::rotate(rotation, angle, axis); }
    void rotate(vec3 angles) { rotation = quatFromEuler(angles) * rotation; }
    void rotateAround(vec3 pivot, vec3 angles) { move(-pivot); rotate(angles); move(pivot); }
    void lookAt(const vec3& target) { rotation = quatLookAt(normalize(target - position), UP); }
    void scale(float factor) { scaling *= factor; }
    void scale(vec3 factors) { scaling *= factors; }

    vec3 position = vec3(0.0f);
    quat rotation = quat{};
    vec3 scaling = vec3(1.0f);
};

inline Transform operator*(const Transform& x, const Transform& y) {
    Transform result;
    result.position = rotate(x.rotation, x.scaling * y.position) + x.position;
    result.rotation = x.rotation * y.rotation;
    result.scaling = x.scaling * y.scaling;
    return result;
}

namespace gpu {
using Material = rg_material;   // resources/shaders/gpu_material.def
using UniformBufferObject = rg_ubo;
struct BufferRef {              // raygun/gpu/gpu_buffer.hpp:67-76 without the device address
    uint32_t offsetInBytes = 0, sizeInBytes = 0, elementSize = 1;
    uint32_t offsetInElements() const { return offsetInBytes / elementSize; }
};
gpu::Material defaultMaterial();  // defaults of gpu_material.def:11-26
}  // namespace gpu

namespace render {

using Vertex = rg_vertex;

struct Mesh {
    std::vector<Vertex> vertices;
    std::vector<uint32_t> indices;
    gpu::BufferRef vertexBufferRef, indexBufferRef;
    uint32_t meshIndex = 0;  // index of this mesh's BLAS (rg_mesh_range) after setupModelBuffers

    size_t numFaces() const { return indices.size() / 3; }
    vec3 center() const;
    struct Bounds { vec3 lower, upper; };
    Bounds bounds() const;
    float width() const;
    void merge(const Mesh& other);
    void forEachFace(std::function<void(const Vertex&, const Vertex&, const Vertex&)>) const;
};

}  // namespace render

struct Material {
    Material() : gpuMaterial(gpu::defaultMaterial()) {}
    /// Loads <path> (an .rgmat.json file); "basedOn" names are resolved through `loadBase`.
    Material(string_view name, const string& path, const std::function<std::shared_ptr<Material>(const string&)>& loadBase = {});
    string name = "default";
    gpu::Material gpuMaterial;
};

namespace render {
struct Model {
    std::shared_ptr<Mesh> mesh;
    std::vector<std::shared_ptr<Material>> materials;
    gpu::BufferRef materialBufferRef;
};
}  // namespace render

class Entity {
  public:
    explicit Entity(string_view name_) : name(name_) {}
    virtual ~Entity() {}

    const Transform& transform() const { return m_transform; }
    void setTransform(Transform transform) { invalidateChildrenCachedParentTransform(); m_transform = transform; }
    Transform parentTransform() const;
    Transform globalTransform() const { return parentTransform() * m_transform; }

    bool isVisible() const { return m_visible; }
    void setVisible(bool visible) { m_visible = visible; }
    void show() { setVisible(true); }
    void hide() { setVisible(false); }

    const std::vector<std::shared_ptr<Entity>>& children() const { return m_children; }
    void addChild(std::shared_ptr<Entity> child);
    std::shared_ptr<Entity> emplaceChild(string_view childName = {});
    void removeChild(const std::shared_ptr<Entity>& child);
    void clearChildren();

    template <typename Fun>
    void forEachEntity(Fun f) {
        bool descend = true;
        if constexpr(std::is_invocable_r_v<bool, Fun, Entity&>) descend = f(*this);
        else f(*this);
        if(!descend) return;
        for(auto& child: m_children) child->forEachEntity(f);
    }

    void move(const vec3& translation) { invalidateChildrenCachedParentTransform(); m_transform.move(translation); }
    void moveTo(const vec3& position) { invalidateChildrenCachedParentTransform(); m_transform.position = position; }
    void rotate(float angle, vec3 axis) { invalidateChildrenCachedParentTransform(); m_transform.rotate(angle, axis); }
    void rotate(vec3 rotation) { invalidateChildrenCachedParentTransform(); m_transform.rotate(rotation); }
    void rotateAround(vec3 pivot, vec3 rotation) { invalidateChildrenCachedParentTransform(); m_transform.rotateAround(pivot, rotation); }
    void lookAt(const vec3& target) { invalidateChildrenCachedParentTransform(); m_transform.lookAt(target); }
    void scale(vec3 s) { invalidateChildrenCachedParentTransform(); m_transform.scale(s); }
    void scale(float s) { invalidateChildrenCachedParentTransform(); m_transform.scale(s); }

    string name;
    std::shared_ptr<render::Model> model;

  private:
    void setParent(const Entity* parent) { m_parent = parent; invalidateCachedParentTransform(); }
    void invalidateCachedParentTransform();
    void invalidateChildrenCachedParentTransform();

    Transform m_transform;
    bool m_visible = true;
    const Entity* m_parent = nullptr;
    mutable std::optional<Transform> m_cachedParentTransform;
    std::vector<std::shared_ptr<Entity>> m_children;
};

class Camera : public Entity {
  public:
    Camera(uint32_t width, uint32_t height) : Entity("Camera") { updateProjection(width, height); }
    mat4 viewInverse() const { return transform().toMat4(); }
    mat4 projInverse() const { return inverse(m_projection); }
    void updateProjection(uint32_t width, uint32_t height);

  private:
    mat4 m_projection;
    static constexpr float FOV = 45.f, NEAR = 0.1f, FAR = 100.0f;
};

struct Scene {
    Scene(uint32_t width, uint32_t height) : camera(std::make_shared<Camera>(width, height)) { root->addChild(camera); }
    virtual ~Scene() {}
    std::shared_ptr<Camera> camera;
    std::shared_ptr<Entity> root = std::make_shared<Entity>("root");
    virtual void update(double) {}
};

/// Asset ingestion the way the reference does it for this path (entity.cpp:60-122 on top of Assimp's Collada importer with
/// aiProcess_Triangulate only, resource_manager.cpp:35-50 for materials): every child node of the file becomes a child entity
/// with its own Model; one vertex per index tuple, sub-meshes merged per node in file order, every model carries all materials.
namespace ui {

/// raygun/ui/text.hpp:31-37: one mesh per ASCII code point, shifted so that its left edge is x = 0, and its width.
struct Font {
    string name;
    std::array<std::shared_ptr<render::Mesh>, 128> charMap = {};
    std::array<float, 128> charWidth = {};
};

enum class Alignment { TopLeft, TopCenter, TopRight, MiddleLeft, MiddleCenter, MiddleRight, BottomLeft, BottomCenter, BottomRight };

/// raygun/ui/text.{hpp,cpp}: text as ordinary entities, one instance per glyph (the ray-traced UI of the reference).
class TextGenerator {
  public:
    using RegisterModel = std::function<void(std::shared_ptr<render::Model>)>;
    TextGenerator(const Font& font, std::shared_ptr<Material> material, const RegisterModel& registerModel, float letterPadding = 0.1f, float lineSpacing = 1.f);
    std::shared_ptr<Entity> text(string_view input, Alignment align = Alignment::TopLeft) const { return textWithBounds(input, align).first; }
    std::pair<std::shared_ptr<Entity>, render::Mesh::Bounds> textWithBounds(string_view input, Alignment align = Alignment::TopLeft) const;

    /// The flat form of a laid-out string: one record per glyph instance (pen position before alignment) + the extent of the block.
    struct Placement { unsigned code; float x, y; };
    struct Layout { std::vector<Placement> glyphs; vec2 extent{}; };
    Layout layout(string_view input) const;
    static vec3 anchorOffset(Alignment align, vec2 extent);

  private:
    std::array<std::shared_ptr<render::Model>, 128> m_charMap = {};
    std::array<float, 128> m_charWidth = {};
    float letterPadding, lineSpacing;
};

}  // namespace ui

class ResourceManager {
  public:
    explicit ResourceManager(string resourcesDir) : m_dir(std::move(resourcesDir)) {}
    std::shared_ptr<Entity> loadEntity(string_view name);          // resources/models/<name>.dae
    std::shared_ptr<Material> loadMaterial(const string& name);   // resources/materials/<name>.rgmat.json (+ underscore fallback)
    std::shared_ptr<ui::Font> loadFont(string_view name);          // resources/fonts/<name>.obj (resource_manager.cpp:107-135)
    const std::vector<std::shared_ptr<render::Model>>& models() const { return m_models; }
    void registerModel(std::shared_ptr<render::Model> m) { m_models.push_back(std::move(m)); }

  private:
    string m_dir;
    std::vector<std::shared_ptr<render::Model>> m_models;
    std::vector<std::pair<string, std::shared_ptr<Material>>> m_materials;
    std::vector<std::pair<string, std::shared_ptr<ui::Font>>> m_fonts;
};

namespace render {

/// raygun::render::Raytracer over the C ABI.
struct Raytracer {
    Raytracer(uint32_t width, uint32_t height, int device);
    ~Raytracer();
    Raytracer(const Raytracer&) = delete;
    void setupBottomLevelAS();
    void setupTopLevelAS(const Scene& scene);
    void updateRenderTarget(const gpu::UniformBufferObject& ubo);
    void doRaytracing(bool useFXAA);
    rg_ctx* ctx = nullptr;
    std::vector<rg_instance> instances;   // the last TLAS input (also useful without a GPU)
    static void gatherInstances(const Scene& scene, std::vector<rg_instance>& out);
    /// true: upload local TRS + parent links and let the GPU walk the scene graph (rg_set_entities) -- for scenes with
    /// thousands of animated entities, where composing every globalTransform on the host dominates the frame.
    bool deviceSceneWalk = false;
    std::vector<rg_entity> entities;
    static void gatherEntities(const Scene& scene, std::vector<rg_entity>& out);
};

/// Seconds since the application started (the reference's RG().time()); injected so that a headless run can step time itself.
using Clock = std::function<double()>;

/// raygun/render/fade.hpp:27-63: a colour laid over the frame by the postprocess pass (ubo.fadeColor); alpha follows an envelope in time.
class Fade {
  public:
    explicit Fade(const Clock& clock) : m_clock(clock), m_start(clock()) {}
    virtual ~Fade() {}
    virtual vec4 curColor();
    virtual bool over() const;

  protected:
    double elapsed() const { return m_clock() - m_start; }

  private:
    Clock m_clock;
    double m_start;
};
class FadeIn : public Fade {   // from opaque `fromColor` to the frame within `duration`
  public:
    FadeIn(const Clock& clock, double duration, vec3 fromColor = vec3(0.f));
    vec4 curColor() override;
    bool over() const override;

  private:
    double m_duration;
    vec3 m_color;
};
class FadeTransition : public Fade {   // frame -> colour (callback at the peak) -> frame, `halfDuration` each way
  public:
    FadeTransition(const Clock& clock, double halfDuration, std::function<void()> atPeak, vec3 color = vec3(0.f));
    vec4 curColor() override;
    bool over() const override;

  private:
    double m_half;
    std::function<void()> m_atPeak;
    vec3 m_color;
    float m_alpha = 0.0f;
    bool m_switched = false;
};

/// Headless RenderSystem: same buffer packing and per-frame order as the reference, no swapchain / ImGui.
class RenderSystem {
  public:
    RenderSystem(uint32_t width, uint32_t height, int device = 0);
    /// render_system.hpp:65-71: a new fade starts only when none is running.  The clock is passed on to the fade.
    template <typename F, typename... Args>
    void makeFade(Args&&... args) {
        if(!m_currentFade || m_currentFade->over()) m_currentFade = std::make_unique<F>(clock, std::forward<Args>(args)...);
    }
    /// time source of fades and ubo.time; default: wall clock since construction.  Set it to step time deterministically.
    Clock clock;
    void writeFramePPM(const string& path);   // instead of the swapchain present (render_system.cpp:159)
    void writeFramePNG(const string& path);
    static void writeImagePPM(const string& path, const uint8_t* rgba, uint32_t width, uint32_t height);
    static void writeImagePNG(const string& path, const uint8_t* rgba, uint32_t width, uint32_t height);
    void setupModelBuffers(const std::vector<std::shared_ptr<Model>>& models);   // render_system.cpp:192-223, :270-330
    void render(Scene& scene);                                                   // render_system.cpp:88-162
    void readFrame(std::vector<uint8_t>& rgba8);
    rg_timings timings();
    gpu::UniformBufferObject& ubo() { return m_ubo; }
    bool useFXAA = true;
    Raytracer& raytracer() { return *m_raytracer; }

    // the packed host buffers (inspectable without a GPU)
    std::vector<Vertex> vertexBuffer;
    std::vector<uint32_t> indexBuffer;
    std::vector<gpu::Material> materialBuffer;
    std::vector<rg_mesh_range> meshRanges;
    static void packModelBuffers(const std::vector<std::shared_ptr<Model>>& models, std::vector<Vertex>& v, std::vector<uint32_t>& i,
                                 std::vector<gpu::Material>& m, std::vector<rg_mesh_range>& ranges);
    static void fillUniformBuffer(gpu::UniformBufferObject& ubo, const Camera& camera);   // render_system.cpp:235-268

  private:
    uint32_t m_width, m_height;
    gpu::UniformBufferObject m_ubo{};
    std::unique_ptr<Raytracer> m_raytracer;
    std::unique_ptr<Fade> m_currentFade;
};

}  // namespace render
}  // namespace raygun
