// host_capi.cpp -- C entry points over the C++ host shim, for the Python tests (ctypes) and as usage examples.
#include <cstring>
#include <stdexcept>
#include <string>

#include "raygun_host.hpp"

using namespace raygun;

namespace {
thread_local std::string g_err;
template <class F>
int guarded(F&& f) {
    try { f(); return 0; } catch(const std::exception& e) { g_err = e.what(); return 1; }
}
void put3x4(const Transform& t, float* out) { const mat4 m = transpose(t.toMat4()); std::memcpy(out, &m.m[0][0], 48); }
void putMat4(const mat4& m, float* out) { std::memcpy(out, &m.m[0][0], 64); }
}  // namespace

extern "C" {

const char* rgh_last_error() { return g_err.c_str(); }

// The quantities of tests/golden/glm_golden.json, computed by rg_math.hpp / Transform / Camera, in this order:
// instance 3x4 of Raygun, ph3_games, room, Ball (48 floats); viewInverse (16); cam quat wxyz (4); projInverse 640x360 (16), 100x60 (16);
// lightDir (3); trs_compose_3x4 (12); decompose pos (3) scale (3) quat wxyz (4); viewInverse_c3 (16).  141 floats.
void rgh_math_golden(float* out) {
    auto rowMajor = [](const float* f) { mat4 m; for(int r = 0; r < 4; ++r) for(int c = 0; c < 4; ++c) m[c][r] = f[r * 4 + c]; return m; };
    const float raygunM[16] = {7.5f, 0, 0, 3, 0, 7.5f, 0, 0, 0, 0, 7.5f, -21, 0, 0, 0, 1};
    const float ph3M[16] = {0.7071068f, 0, 0.7071068f, -9, 0, 1, 0, 0, -0.7071068f, 0, 0.7071068f, -21, 0, 0, 0, 1};
    const float roomM[16] = {1, 0, 0, -24, 0, 1, 0, -4, 0, 0, 1, -24, 0, 0, 0, 1};
    Entity root("root");
    auto level = root.emplaceChild("room");
    auto a = level->emplaceChild("Raygun"); a->setTransform(Transform(rowMajor(raygunM)));
    auto b = level->emplaceChild("ph3_games"); b->setTransform(Transform(rowMajor(ph3M)));
    auto c = level->emplaceChild("room"); c->setTransform(Transform(rowMajor(roomM)));
    auto ball = root.emplaceChild("Ball"); ball->moveTo({3.0f, 0.0f, -3.0f});
    float* o = out;
    put3x4(a->globalTransform(), o); o += 12; put3x4(b->globalTransform(), o); o += 12; put3x4(c->globalTransform(), o); o += 12;
    put3x4(ball->globalTransform(), o); o += 12;
    Camera cam(640, 360);
    cam.moveTo(ball->transform().position + vec3(5.0f, 10.0f, 10.0f));
    cam.lookAt(ball->transform().position);
    putMat4(cam.viewInverse(), o); o += 16;
    const quat q = cam.transform().rotation; o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z; o += 4;
    putMat4(cam.projInverse(), o); o += 16;
    cam.updateProjection(100, 60); putMat4(cam.projInverse(), o); o += 16;
    const vec3 l = normalize(vec3(.4f, -.6f, -.8f)); o[0] = l.x; o[1] = l.y; o[2] = l.z; o += 3;
    Transform parent; parent.position = {1.5f, -2.25f, 0.75f}; parent.rotation = rotate(quat{}, 0.7f, normalize(vec3(1, 2, 3))); parent.scaling = vec3(2.0f);
    Transform child; child.position = {-0.5f, 4.0f, 1.0f}; child.rotation = quatFromEuler({0.1f, -0.4f, 0.9f}); child.scaling = {0.5f, 1.5f, 1.0f};
    const Transform pc = parent * child;
    put3x4(pc, o); o += 12;
    const Transform dec(pc.toMat4());
    o[0] = dec.position.x; o[1] = dec.position.y; o[2] = dec.position.z; o += 3;
    o[0] = dec.scaling.x; o[1] = dec.scaling.y; o[2] = dec.scaling.z; o += 3;
    o[0] = dec.rotation.w; o[1] = dec.rotation.x; o[2] = dec.rotation.y; o[3] = dec.rotation.z; o += 4;
    Transform cam2; cam2.position = {35.f, 18.f, -20.f}; cam2.lookAt({33.75f, 1.f, 33.75f});
    putMat4(cam2.toMat4(), o);
}

// ---- the example scene through ResourceManager (needs the reference's resources directory: authoring container only)
struct rgh_scene {
    std::unique_ptr<ResourceManager> rm;
    std::unique_ptr<Scene> scene;
    std::shared_ptr<Entity> ball;
    std::vector<render::Vertex> v; std::vector<uint32_t> i; std::vector<gpu::Material> m; std::vector<rg_mesh_range> ranges;
    std::vector<rg_instance> inst;
    gpu::UniformBufferObject ubo{};
    std::unique_ptr<render::RenderSystem> rs;
};

// example/example_scene.cpp:9-32 + :58-62 (camera follows the ball), ball.cpp:11-22 (the Ball entity takes its child's model)
rgh_scene* rgh_example_scene_load(const char* resourcesDir, uint32_t W, uint32_t H) {
    auto* s = new rgh_scene();
    if(guarded([&] {
           s->rm = std::make_unique<ResourceManager>(resourcesDir);
           s->scene = std::make_unique<Scene>(W, H);
           auto level = s->rm->loadEntity("room");
           s->scene->root->addChild(level);
           auto loaded = s->rm->loadEntity("ball");
           s->ball = std::make_shared<Entity>("Ball");
           s->ball->model = loaded->children().at(0)->model;
           s->ball->moveTo({3.0f, 0.0f, -3.0f});
           s->scene->root->addChild(s->ball);
           s->scene->camera->moveTo(s->ball->transform().position + vec3(5.0f, 10.0f, 10.0f));
           s->scene->camera->lookAt(s->ball->transform().position);
           render::RenderSystem::packModelBuffers(s->rm->models(), s->v, s->i, s->m, s->ranges);
           render::Raytracer::gatherInstances(*s->scene, s->inst);
           std::memset(&s->ubo, 0, sizeof s->ubo);
           render::RenderSystem::fillUniformBuffer(s->ubo, *s->scene->camera);
       })) { delete s; return nullptr; }
    return s;
}

// ---- ray-traced text (ui/text.cpp): `text` laid out by TextGenerator with the font resources/fonts/<font>.obj and the material
// resources/materials/<material>.rgmat.json, above the example scene's floor material as a ground quad; only the glyphs the text
// uses are turned into models (the reference registers all 124; the fixture stays small).  widths128 receives Font::charWidth.
rgh_scene* rgh_text_scene_load(const char* resourcesDir, const char* fontName, const char* materialName, const char* text, int align, uint32_t W, uint32_t H,
                               float* widths128, uint32_t* glyphCount) {
    auto* s = new rgh_scene();
    if(guarded([&] {
           s->rm = std::make_unique<ResourceManager>(resourcesDir);
           s->scene = std::make_unique<Scene>(W, H);
           auto font = s->rm->loadFont(fontName);
           uint32_t n = 0;
           for(size_t k = 0; k < font->charMap.size(); ++k) { if(widths128) widths128[k] = font->charWidth[k]; n += font->charMap[k] ? 1u : 0u; }
           if(glyphCount) *glyphCount = n;
           ui::Font used = *font;
           for(size_t k = 0; k < used.charMap.size(); ++k)
               if(!std::strchr(text, (int)k) || k == 0) used.charMap[k].reset();
           ui::TextGenerator gen(used, s->rm->loadMaterial(materialName), [&](std::shared_ptr<render::Model> m) { s->rm->registerModel(std::move(m)); });
           auto [ent, bounds] = gen.textWithBounds(text, (ui::Alignment)align);
           ent->moveTo({0.0f, 1.0f, 0.0f});
           s->scene->root->addChild(ent);
           // ground: one quad with the floor material (grid effect, reflections of the glyphs)
           auto ground = std::make_shared<render::Model>();
           ground->mesh = std::make_shared<render::Mesh>();
           const float e = 12.0f;
           const float q[4][2] = {{-e, -e}, {e, -e}, {e, e}, {-e, e}};
           for(auto& c: q) { render::Vertex v{}; v.position[0] = c[0]; v.position[2] = c[1]; v.normal[1] = 1.0f; ground->mesh->vertices.push_back(v); }
           ground->mesh->indices = {0, 2, 1, 0, 3, 2};
           ground->materials.push_back(s->rm->loadMaterial("floor"));
           s->rm->registerModel(ground);
           auto g = std::make_shared<Entity>("ground");
           g->model = ground;
           s->scene->root->addChild(g);
           const float cx = 0.5f * (bounds.lower.x + bounds.upper.x);
           s->scene->camera->moveTo({cx + 1.5f, 2.2f, 6.0f});
           s->scene->camera->lookAt({cx, 1.2f, 0.0f});
           render::RenderSystem::packModelBuffers(s->rm->models(), s->v, s->i, s->m, s->ranges);
           render::Raytracer::gatherInstances(*s->scene, s->inst);
           std::memset(&s->ubo, 0, sizeof s->ubo);
           render::RenderSystem::fillUniformBuffer(s->ubo, *s->scene->camera);
       })) { delete s; return nullptr; }
    return s;
}
void rgh_scene_free(rgh_scene* s) { delete s; }
void rgh_scene_counts(const rgh_scene* s, uint32_t* out5) {
    out5[0] = (uint32_t)s->v.size(); out5[1] = (uint32_t)s->i.size(); out5[2] = (uint32_t)s->m.size(); out5[3] = (uint32_t)s->ranges.size(); out5[4] = (uint32_t)s->inst.size();
}
void rgh_scene_copy(const rgh_scene* s, void* vertices, uint32_t* indices, void* materials, uint32_t* ranges, void* instances, void* ubo) {
    std::memcpy(vertices, s->v.data(), s->v.size() * sizeof(render::Vertex));
    std::memcpy(indices, s->i.data(), s->i.size() * 4);
    std::memcpy(materials, s->m.data(), s->m.size() * sizeof(gpu::Material));
    std::memcpy(ranges, s->ranges.data(), s->ranges.size() * sizeof(rg_mesh_range));
    std::memcpy(instances, s->inst.data(), s->inst.size() * sizeof(rg_instance));
    std::memcpy(ubo, &s->ubo, sizeof s->ubo);
}

// ---- a scene assembled from plain arrays, rendered through Entity / Scene / Camera / RenderSystem (GPU)
// models: per model {vtx_off, vtx_cnt, idx_off, idx_cnt, mat_first, mat_cnt}; entities: per entity {parent (-1 = root), model (-1 = none), visible}
// + TRS {px,py,pz, qw,qx,qy,qz, sx,sy,sz}.  Writes the RGBA8 frame and the instance list the DFS produced.
int rgh_render_entities(const void* vertices, uint32_t nVtx, const uint32_t* indices, uint32_t nIdx, const void* materials, uint32_t nMat,
                        const uint32_t* models, uint32_t nModels, const int32_t* entities, const float* trs, uint32_t nEntities, const float* camPos,
                        const float* camTarget, uint32_t W, uint32_t H, int numSamples, int maxRecursions, int useFXAA, int device, uint8_t* rgba8,
                        void* instancesOut, uint32_t* nInstancesOut, float* timingsOut /*as_build, rt_total*/) {
    return guarded([&] {
        (void)nVtx; (void)nIdx; (void)nMat;
        const auto* V = (const render::Vertex*)vertices;
        const auto* M = (const gpu::Material*)materials;
        std::vector<std::shared_ptr<render::Model>> mods;
        for(uint32_t k = 0; k < nModels; ++k) {
            const uint32_t* d = models + 6 * k;
            auto model = std::make_shared<render::Model>();
            model->mesh = std::make_shared<render::Mesh>();
            model->mesh->vertices.assign(V + d[0], V + d[0] + d[1]);
            model->mesh->indices.assign(indices + d[2], indices + d[2] + d[3]);
            for(uint32_t j = 0; j < d[5]; ++j) { auto mat = std::make_shared<Material>(); mat->gpuMaterial = M[d[4] + j]; model->materials.push_back(mat); }
            mods.push_back(model);
        }
        Scene scene(W, H);
        std::vector<std::shared_ptr<Entity>> ents;
        for(uint32_t k = 0; k < nEntities; ++k) {
            const int32_t* e = entities + 3 * k; const float* t = trs + 10 * k;
            auto ent = std::make_shared<Entity>("e" + std::to_string(k));
            Transform tr; tr.position = {t[0], t[1], t[2]}; tr.rotation = {t[3], t[4], t[5], t[6]}; tr.scaling = {t[7], t[8], t[9]};
            ent->setTransform(tr);
            if(e[1] >= 0) ent->model = mods.at((size_t)e[1]);
            ent->setVisible(e[2] != 0);
            (e[0] < 0 ? scene.root : ents.at((size_t)e[0]))->addChild(ent);
            ents.push_back(ent);
        }
        scene.camera->moveTo({camPos[0], camPos[1], camPos[2]});
        scene.camera->lookAt({camTarget[0], camTarget[1], camTarget[2]});
        render::RenderSystem rs(W, H, device);
        rs.ubo().num_samples = numSamples; rs.ubo().max_recursions = maxRecursions;
        rs.useFXAA = (useFXAA & 1) != 0;
        rs.raytracer().deviceSceneWalk = (useFXAA & 2) != 0;   // bit 1: scene-graph walk on the GPU (rg_set_entities)
        rs.setupModelBuffers(mods);
        rs.raytracer().setupBottomLevelAS();
        rs.render(scene);
        std::vector<uint8_t> frame;
        rs.readFrame(frame);
        std::memcpy(rgba8, frame.data(), frame.size());
        if(rs.raytracer().deviceSceneWalk) {   // what the device composed
            uint32_t n = 0;
            rg_debug_read_instances(rs.raytracer().ctx, nullptr, 0, &n);
            if(instancesOut && n) rg_debug_read_instances(rs.raytracer().ctx, (rg_instance*)instancesOut, n, &n);
            if(nInstancesOut) *nInstancesOut = n;
        } else {
            const auto& inst = rs.raytracer().instances;
            if(instancesOut) std::memcpy(instancesOut, inst.data(), inst.size() * sizeof(rg_instance));
            if(nInstancesOut) *nInstancesOut = (uint32_t)inst.size();
        }
        if(timingsOut) { const rg_timings t = rs.timings(); timingsOut[0] = t.as_build_ms; timingsOut[1] = t.rt_total_ms; }
    });
}

// DFS only (no GPU): same entity description as above -> instance list.
int rgh_gather_instances(const uint32_t* models, uint32_t nModels, const int32_t* entities, const float* trs, uint32_t nEntities, void* instancesOut,
                         uint32_t* nInstancesOut) {
    return guarded([&] {
        std::vector<std::shared_ptr<render::Model>> mods;
        for(uint32_t k = 0; k < nModels; ++k) {
            const uint32_t* d = models + 6 * k;
            auto model = std::make_shared<render::Model>();
            model->mesh = std::make_shared<render::Mesh>();
            model->mesh->meshIndex = k;
            model->mesh->vertexBufferRef = {d[0] * 32u, d[1] * 32u, 32u};
            model->mesh->indexBufferRef = {d[2] * 4u, d[3] * 4u, 4u};
            model->materialBufferRef = {d[4] * 64u, d[5] * 64u, 64u};
            mods.push_back(model);
        }
        Scene scene(16, 9);
        std::vector<std::shared_ptr<Entity>> ents;
        for(uint32_t k = 0; k < nEntities; ++k) {
            const int32_t* e = entities + 3 * k; const float* t = trs + 10 * k;
            auto ent = std::make_shared<Entity>("e" + std::to_string(k));
            Transform tr; tr.position = {t[0], t[1], t[2]}; tr.rotation = {t[3], t[4], t[5], t[6]}; tr.scaling = {t[7], t[8], t[9]};
            ent->setTransform(tr);
            if(e[1] >= 0) ent->model = mods.at((size_t)e[1]);
            ent->setVisible(e[2] != 0);
            (e[0] < 0 ? scene.root : ents.at((size_t)e[0]))->addChild(ent);
            ents.push_back(ent);
        }
        std::vector<rg_instance> inst;
        render::Raytracer::gatherInstances(scene, inst);
        std::memcpy(instancesOut, inst.data(), inst.size() * sizeof(rg_instance));
        *nInstancesOut = (uint32_t)inst.size();
    });
}


// ---- ui::TextGenerator::layout without a font file: widths128[c] > 0 marks the code points that have a glyph.  out_xy receives the
// aligned glyph positions (anchor + pen), bounds4 = lower.xy, upper.xy.  Returns the number of glyph instances.
int rgh_text_layout(const char* text, int align, const float* widths128, float letterPadding, float lineSpacing, float* out_xy, uint32_t cap, float* bounds4) {
    int n = -1;
    guarded([&] {
        ui::Font font;
        auto dummy = std::make_shared<render::Mesh>();
        for(size_t k = 0; k < 128; ++k) { font.charWidth[k] = widths128[k]; if(widths128[k] > 0.0f) font.charMap[k] = dummy; }
        ui::TextGenerator gen(font, nullptr, nullptr, letterPadding, lineSpacing);
        auto [ent, bounds] = gen.textWithBounds(text, (ui::Alignment)align);
        uint32_t k = 0;
        ent->forEachEntity([&](Entity& e) {
            if(!e.model) return;
            const Transform g = e.globalTransform();
            if(k < cap) { out_xy[2 * k] = g.position.x; out_xy[2 * k + 1] = g.position.y; }
            ++k;
        });
        if(bounds4) { bounds4[0] = bounds.lower.x; bounds4[1] = bounds.lower.y; bounds4[2] = bounds.upper.x; bounds4[3] = bounds.upper.y; }
        n = (int)k;
    });
    return n;
}

// ---- render::FadeIn (kind 0) / FadeTransition (kind 1) sampled at the given times (seconds since the fade was made); rgba receives
// 4 floats per sample, peak_index the first sample at which the transition callback fired (-1: never).
int rgh_fade_sample(int kind, double duration, const float* color3, const double* times, uint32_t n, float* rgba, int* peak_index, int* over_flags) {
    return guarded([&] {
        double now = 0.0;
        render::Clock clock = [&now] { return now; };
        int peak = -1, cur = 0;
        std::unique_ptr<render::Fade> f;
        const vec3 c(color3[0], color3[1], color3[2]);
        if(kind == 0) f = std::make_unique<render::FadeIn>(clock, duration, c);
        else f = std::make_unique<render::FadeTransition>(clock, duration, [&] { if(peak < 0) peak = cur; }, c);
        for(uint32_t k = 0; k < n; ++k) {
            now = times[k]; cur = (int)k;
            const vec4 v = f->curColor();
            rgba[4 * k] = v.x; rgba[4 * k + 1] = v.y; rgba[4 * k + 2] = v.z; rgba[4 * k + 3] = v.w;
            if(over_flags) over_flags[k] = f->over() ? 1 : 0;
        }
        if(peak_index) *peak_index = peak;
    });
}

// ---- the frame writers that stand in for the swapchain present
int rgh_write_image(const char* path, const uint8_t* rgba, uint32_t W, uint32_t H, int png) {
    return guarded([&] { if(png) render::RenderSystem::writeImagePNG(path, rgba, W, H); else render::RenderSystem::writeImagePPM(path, rgba, W, H); });
}
}  // extern "C"
