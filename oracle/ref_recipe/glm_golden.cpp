// TEST INFRASTRUCTURE (oracle/_ref recipe) -- not product code.
//
// Generates golden camera / transform matrices with the reference's OWN vendored
// GLM (vendor/glm, 0.9.9.5) configured exactly as raygun/pch.hpp:82-87 does
// (GLM_FORCE_DEPTH_ZERO_TO_ONE + the same gtc/gtx headers).  The expressions below
// restate, call for call, what the reference evaluates:
//   Transform(mat4)            raygun/transform.hpp:31-36   (glm::decompose)
//   Transform::toMat4          raygun/transform.hpp:38-46   (T * R * S)
//   operator*(Transform,..)    raygun/transform.hpp:99-106
//   Transform::lookAt          raygun/transform.hpp:82-86   (glm::quatLookAt)
//   Camera::updateProjection   raygun/camera.cpp:34-47      (perspective, [1][1] *= -1)
//   Camera::projInverse        raygun/camera.hpp:36         (glm::inverse)
//   instance 3x4               raygun/render/acceleration_structure.cpp:44-45
// Output: JSON on stdout; every float is written as its IEEE-754 bit pattern.
// Built only inside the authoring container (needs /root/reference); the JSON it
// prints is committed under tests/golden/ together with this file.
#define GLM_FORCE_DEPTH_ZERO_TO_ONE
#define GLM_ENABLE_EXPERIMENTAL
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
#include <glm/gtc/quaternion.hpp>
#include <glm/gtx/euler_angles.hpp>
#include <glm/gtx/matrix_decompose.hpp>
#include <glm/gtx/projection.hpp>
#include <glm/gtx/quaternion.hpp>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

using glm::mat4; using glm::quat; using glm::vec3; using glm::vec4;

struct Transform {
    vec3 position = glm::zero<vec3>();
    quat rotation = glm::identity<quat>();
    vec3 scaling = glm::one<vec3>();
    Transform() {}
    explicit Transform(const mat4& m) { vec3 skew; vec4 persp; glm::decompose(m, scaling, rotation, position, skew, persp); }
    mat4 toMat4() const {
        const auto id = glm::identity<mat4>();
        return glm::translate(id, position) * glm::toMat4(rotation) * glm::scale(id, scaling);
    }
};
static Transform mul(const Transform& x, const Transform& y) {
    Transform r;
    r.position = glm::rotate(x.rotation, x.scaling * y.position) + x.position;
    r.rotation = x.rotation * y.rotation;
    r.scaling = x.scaling * y.scaling;
    return r;
}

static uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static void printMat4ColMajor(const char* name, const mat4& m, bool last = false) {
    std::printf("  \"%s\": [", name);
    for(int c = 0; c < 4; ++c) for(int r = 0; r < 4; ++r) std::printf("%u%s", bits(m[c][r]), (c == 3 && r == 3) ? "" : ", ");
    std::printf("]%s\n", last ? "" : ",");
}
static void print3x4RowMajor(const char* name, const mat4& m) {
    // glm::transpose(m) reinterpret as float[3][4] == rows 0..2 of m, row-major
    const mat4 t = glm::transpose(m);
    const float* f = &t[0][0];
    std::printf("  \"%s\": [", name);
    for(int i = 0; i < 12; ++i) std::printf("%u%s", bits(f[i]), i == 11 ? "" : ", ");
    std::printf("],\n");
}
static mat4 rowMajor16(const float* f) {
    // Collada <matrix> is row-major; Assimp keeps row-major aiMatrix4x4; the reference
    // transposes it and reinterprets as glm::mat4 (raygun/utils/assimp_utils.hpp:29-33),
    // i.e. element (row r, col c) lands at glm m[c][r].
    mat4 m;
    for(int r = 0; r < 4; ++r) for(int c = 0; c < 4; ++c) m[c][r] = f[r * 4 + c];
    return m;
}

int main() {
    std::printf("{\n");
    // --- node transforms of room.dae (resources/models/room.dae:240,250,265) and ball.dae:75
    const float raygunM[16] = {7.5f, 0, 0, 3, 0, 7.5f, 0, 0, 0, 0, 7.5f, -21, 0, 0, 0, 1};
    const float ph3M[16] = {0.7071068f, 0, 0.7071068f, -9, 0, 1, 0, 0, -0.7071068f, 0, 0.7071068f, -21, 0, 0, 0, 1};
    const float roomM[16] = {1, 0, 0, -24, 0, 1, 0, -4, 0, 0, 1, -24, 0, 0, 0, 1};
    const Transform root, level;  // scene root and the loaded "room" entity: identity
    const Transform tR = mul(mul(root, level), Transform(rowMajor16(raygunM)));
    const Transform tP = mul(mul(root, level), Transform(rowMajor16(ph3M)));
    const Transform tO = mul(mul(root, level), Transform(rowMajor16(roomM)));
    Transform ball; ball.position = vec3(3.0f, 0.0f, -3.0f);  // example/example_scene.cpp:22
    const Transform tB = mul(root, ball);
    print3x4RowMajor("instance_Raygun", tR.toMat4());
    print3x4RowMajor("instance_ph3_games", tP.toMat4());
    print3x4RowMajor("instance_room", tO.toMat4());
    print3x4RowMajor("instance_Ball", tB.toMat4());

    // --- camera: example/example_scene.cpp:58-62, CAMERA_OFFSET example_scene.hpp:16
    Transform cam;
    cam.position = ball.position + vec3(5.0f, 10.0f, 10.0f);
    cam.rotation = glm::quatLookAt(glm::normalize(ball.position - cam.position), vec3(0, 1, 0));
    printMat4ColMajor("viewInverse", cam.toMat4());
    std::printf("  \"cam_quat_wxyz\": [%u, %u, %u, %u],\n", bits(cam.rotation.w), bits(cam.rotation.x), bits(cam.rotation.y), bits(cam.rotation.z));

    const int sizes[][2] = {{640, 360}, {1920, 1080}, {3840, 2160}, {64, 36}, {256, 144}, {100, 60}};
    for(auto& s: sizes) {
        float aspect = (float)s[0] / (float)s[1];
        mat4 proj = glm::perspective(glm::radians(45.f), aspect, 0.1f, 100.0f);
        proj[1][1] *= -1;
        char name[64]; std::snprintf(name, sizeof name, "projInverse_%dx%d", s[0], s[1]);
        printMat4ColMajor(name, glm::inverse(proj));
    }
    // lightDir: raygun/render/render_system.cpp:241
    const vec3 l = glm::normalize(vec3(.4f, -.6f, -.8f));
    std::printf("  \"lightDir\": [%u, %u, %u],\n", bits(l.x), bits(l.y), bits(l.z));

    // --- a second, non-trivial pose to exercise decompose / TRS composition / quat rotate
    Transform parent; parent.position = vec3(1.5f, -2.25f, 0.75f);
    parent.rotation = glm::rotate(glm::identity<quat>(), 0.7f, glm::normalize(vec3(1, 2, 3)));
    parent.scaling = vec3(2.0f, 2.0f, 2.0f);
    Transform child; child.position = vec3(-0.5f, 4.0f, 1.0f);
    child.rotation = quat(vec3(0.1f, -0.4f, 0.9f));  // euler ctor as in Transform::rotate(vec3)
    child.scaling = vec3(0.5f, 1.5f, 1.0f);
    const Transform pc = mul(parent, child);
    print3x4RowMajor("trs_compose_3x4", pc.toMat4());
    const Transform dec(pc.toMat4());
    std::printf("  \"decompose_pos\": [%u, %u, %u],\n", bits(dec.position.x), bits(dec.position.y), bits(dec.position.z));
    std::printf("  \"decompose_scale\": [%u, %u, %u],\n", bits(dec.scaling.x), bits(dec.scaling.y), bits(dec.scaling.z));
    std::printf("  \"decompose_quat_wxyz\": [%u, %u, %u, %u],\n", bits(dec.rotation.w), bits(dec.rotation.x), bits(dec.rotation.y), bits(dec.rotation.z));
    Transform cam2; cam2.position = vec3(35.f, 18.f, -20.f);
    cam2.rotation = glm::quatLookAt(glm::normalize(vec3(33.75f, 1.f, 33.75f) - cam2.position), vec3(0, 1, 0));
    printMat4ColMajor("viewInverse_c3", cam2.toMat4(), true);
    std::printf("}\n");
    return 0;
}
