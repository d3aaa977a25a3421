// rg_math.hpp -- the small slice of GLM the reference's host path uses (raygun/pch.hpp:82-87, 150-161), written out:
// vec3 / quat / mat4 (column-major), TRS helpers, RH zero-to-one perspective, quatLookAt, decompose, inverse.
// Checked against the reference's own vendored GLM through tests/golden/glm_golden.json (tests/test_host_shim.py).
#pragma once
#include <cmath>
#include <cstdint>

namespace raygun {

struct vec2 { float x = 0, y = 0; };
struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() = default;
    constexpr vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    explicit constexpr vec3(float s) : x(s), y(s), z(s) {}
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4 { float x = 0, y = 0, z = 0, w = 0; };
struct quat { float w = 1, x = 0, y = 0, z = 0; };
struct mat4 {  // column-major like GLM: m[c][r]
    float m[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    float* operator[](int c) { return m[c]; }
    const float* operator[](int c) const { return m[c]; }
};

constexpr vec3 UP = {0.0f, 1.0f, 0.0f};       // pch.hpp:159-161
constexpr vec3 RIGHT = {1.0f, 0.0f, 0.0f};
constexpr vec3 FORWARD = {0.0f, 0.0f, -1.0f};

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator/(float s, vec3 a) { return {s / a.x, s / a.y, s / a.z}; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline bool operator==(vec3 a, vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline vec3 lerp(vec3 a, vec3 b, float t) { return a + (b - a) * t; }

inline quat operator*(quat p, quat q) {
    return {p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z, p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y,
            p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z, p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x};
}
inline bool operator==(quat a, quat b) { return a.w == b.w && a.x == b.x && a.y == b.y && a.z == b.z; }
inline quat conjugate(quat q) { return {q.w, -q.x, -q.y, -q.z}; }
inline quat inverse(quat q) { const float d = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z; return {q.w / d, -q.x / d, -q.y / d, -q.z / d}; }
inline vec3 rotate(quat q, vec3 v) {  // glm::rotate(quat, vec3) == q * v
    const vec3 qv{q.x, q.y, q.z};
    const vec3 uv = cross(qv, v), uuv = cross(qv, uv);
    return v + ((uv * q.w) + uuv) * 2.0f;
}
inline quat angleAxis(float angle, vec3 axis) { const float s = std::sin(angle * 0.5f); return {std::cos(angle * 0.5f), axis.x * s, axis.y * s, axis.z * s}; }
inline quat rotate(quat q, float angle, vec3 axis) {  // glm::rotate(quat, angle, axis): axis is normalised first
    const float len = length(axis);
    if(std::fabs(len - 1.0f) > 0.001f) axis = axis * (1.0f / len);
    return q * angleAxis(angle, axis);
}
inline quat quatFromEuler(vec3 e) {  // glm::quat(vec3 eulerAngles)
    const vec3 c{std::cos(e.x * 0.5f), std::cos(e.y * 0.5f), std::cos(e.z * 0.5f)}, s{std::sin(e.x * 0.5f), std::sin(e.y * 0.5f), std::sin(e.z * 0.5f)};
    return {c.x * c.y * c.z + s.x * s.y * s.z, s.x * c.y * c.z - c.x * s.y * s.z, c.x * s.y * c.z + s.x * c.y * s.z, c.x * c.y * s.z - s.x * s.y * c.z};
}

inline mat4 toMat4(quat q) {  // glm::mat4_cast
    mat4 r;
    const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z, qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z, qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    r[0][0] = 1.0f - 2.0f * (qyy + qzz); r[0][1] = 2.0f * (qxy + qwz); r[0][2] = 2.0f * (qxz - qwy);
    r[1][0] = 2.0f * (qxy - qwz); r[1][1] = 1.0f - 2.0f * (qxx + qzz); r[1][2] = 2.0f * (qyz + qwx);
    r[2][0] = 2.0f * (qxz + qwy); r[2][1] = 2.0f * (qyz - qwx); r[2][2] = 1.0f - 2.0f * (qxx + qyy);
    return r;
}
inline quat quatCast(const vec3 c[3]) {  // glm::quat_cast(mat3), c = columns
    const float fx = c[0].x - c[1].y - c[2].z, fy = c[1].y - c[0].x - c[2].z, fz = c[2].z - c[0].x - c[1].y, fw = c[0].x + c[1].y + c[2].z;
    int big = 0; float best = fw;
    if(fx > best) { best = fx; big = 1; }
    if(fy > best) { best = fy; big = 2; }
    if(fz > best) { best = fz; big = 3; }
    const float bv = std::sqrt(best + 1.0f) * 0.5f, mult = 0.25f / bv;
    switch(big) {
    case 0: return {bv, (c[1].z - c[2].y) * mult, (c[2].x - c[0].z) * mult, (c[0].y - c[1].x) * mult};
    case 1: return {(c[1].z - c[2].y) * mult, bv, (c[0].y + c[1].x) * mult, (c[2].x + c[0].z) * mult};
    case 2: return {(c[2].x - c[0].z) * mult, (c[0].y + c[1].x) * mult, bv, (c[1].z + c[2].y) * mult};
    default: return {(c[0].y - c[1].x) * mult, (c[2].x + c[0].z) * mult, (c[1].z + c[2].y) * mult, bv};
    }
}
inline quat quatLookAt(vec3 direction, vec3 up) {  // glm::quatLookAtRH
    vec3 c[3];
    c[2] = -direction;
    c[0] = normalize(cross(up, c[2]));
    c[1] = cross(c[2], c[0]);
    return quatCast(c);
}

inline mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r;
    for(int c = 0; c < 4; ++c)
        for(int row = 0; row < 4; ++row) r[c][row] = a[0][row] * b[c][0] + a[1][row] * b[c][1] + a[2][row] * b[c][2] + a[3][row] * b[c][3];
    return r;
}
inline mat4 translate(vec3 p) { mat4 r; r[3][0] = p.x; r[3][1] = p.y; r[3][2] = p.z; return r; }
inline mat4 scale(vec3 s) { mat4 r; r[0][0] = s.x; r[1][1] = s.y; r[2][2] = s.z; return r; }
inline mat4 transpose(const mat4& a) { mat4 r; for(int c = 0; c < 4; ++c) for(int row = 0; row < 4; ++row) r[c][row] = a[row][c]; return r; }

inline mat4 perspectiveRH_ZO(float fovy, float aspect, float zNear, float zFar) {  // GLM_FORCE_DEPTH_ZERO_TO_ONE
    const float t = std::tan(fovy / 2.0f);
    mat4 r;
    for(int c = 0; c < 4; ++c) for(int row = 0; row < 4; ++row) r[c][row] = 0.0f;
    r[0][0] = 1.0f / (aspect * t);
    r[1][1] = 1.0f / t;
    r[2][2] = zFar / (zNear - zFar);
    r[2][3] = -1.0f;
    r[3][2] = -(zFar * zNear) / (zFar - zNear);
    return r;
}

inline mat4 inverse(const mat4& m) {  // cofactor expansion, as glm::inverse
    const float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    const float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    const float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    const float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    const float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    const float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    const float v0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, v1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
    const float v2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, v3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
    const float sa[4] = {+1, -1, +1, -1}, sb[4] = {-1, +1, -1, +1};
    mat4 inv;
    for(int i = 0; i < 4; ++i) {
        inv[0][i] = (v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i]) * sa[i];
        inv[1][i] = (v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i]) * sb[i];
        inv[2][i] = (v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i]) * sa[i];
        inv[3][i] = (v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i]) * sb[i];
    }
    const float det = (m[0][0] * inv[0][0] + m[0][1] * inv[1][0]) + (m[0][2] * inv[2][0] + m[0][3] * inv[3][0]);
    const float r = 1.0f / det;
    for(int c = 0; c < 4; ++c) for(int row = 0; row < 4; ++row) inv[c][row] *= r;
    return inv;
}

// glm::decompose restricted to what Transform(mat4) keeps (transform.hpp:31-36): translation, scale, rotation.
inline void decompose(const mat4& mat, vec3& scaling, quat& rotation, vec3& position) {
    position = {mat[3][0], mat[3][1], mat[3][2]};
    vec3 row[3] = {{mat[0][0], mat[0][1], mat[0][2]}, {mat[1][0], mat[1][1], mat[1][2]}, {mat[2][0], mat[2][1], mat[2][2]}};
    scaling.x = length(row[0]); row[0] = row[0] * (1.0f / scaling.x);
    float skewXY = dot(row[0], row[1]); row[1] = row[1] + row[0] * (-skewXY);
    scaling.y = length(row[1]); row[1] = row[1] * (1.0f / scaling.y);
    float skewXZ = dot(row[0], row[2]); row[2] = row[2] + row[0] * (-skewXZ);
    float skewYZ = dot(row[1], row[2]); row[2] = row[2] + row[1] * (-skewYZ);
    scaling.z = length(row[2]); row[2] = row[2] * (1.0f / scaling.z);
    if(dot(row[0], cross(row[1], row[2])) < 0) { scaling = scaling * -1.0f; for(auto& r: row) r = r * -1.0f; }
    rotation = quatCast(row);
}

}  // namespace raygun
