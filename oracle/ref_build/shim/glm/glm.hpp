// GLM-compatible shim (NOT GLM): the subset of the GLM 0.9.7/0.9.8 API that the
// RGKrt hot-path translation units use, written from GLM's published formulas
// (SURVEY.md Appendix B).  It exists only so that the reference's own sources
// under /root/reference/src can be compiled here into oracle/_ref/ -- the
// container has no GLM.  Test infrastructure; never linked into the product.
//
// Operation order matters for bit-level parity, so every function states the
// GLM formula it follows.  Default-constructed matrices/quaternions are identity
// (GLM <= 0.9.8 behaviour, which src/config.cpp:468,488 relies on).
#pragma once
#include <cmath>
#include <string>
#include <cstdio>
#include <algorithm>

namespace glm {

struct vec2 {
    union { float x, r, s; };
    union { float y, g, t; };
    vec2() : x(0), y(0) {}
    explicit vec2(float v) : x(v), y(v) {}
    template <class A, class B> vec2(A a, B b) : x((float)a), y((float)b) {}
    float& operator[](int i) { return i == 0 ? x : y; }
    const float& operator[](int i) const { return i == 0 ? x : y; }
};

struct vec4;
struct vec3 {
    union { float x, r, s; };
    union { float y, g, t; };
    union { float z, b, p; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float v) : x(v), y(v), z(v) {}
    explicit vec3(double v) : x((float)v), y((float)v), z((float)v) {}
    template <class A, class B, class C> vec3(A a, B b_, C c) : x((float)a), y((float)b_), z((float)c) {}
    template <class C> vec3(const vec2& v, C c) : x(v.x), y(v.y), z((float)c) {}
    explicit vec3(const vec4& v);
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    vec2 xy() const { return vec2(x, y); }
    vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    vec3& operator*=(float s_) { x *= s_; y *= s_; z *= s_; return *this; }
};

struct vec4 {
    union { float x, r; };
    union { float y, g; };
    union { float z, b; };
    union { float w, a; };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float v) : x(v), y(v), z(v), w(v) {}
    template <class A, class B, class C, class D> vec4(A a_, B b_, C c, D d) : x((float)a_), y((float)b_), z((float)c), w((float)d) {}
    template <class D> vec4(const vec3& v, D d) : x(v.x), y(v.y), z(v.z), w((float)d) {}
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    vec3 xyz() const { return vec3(x, y, z); }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

// ---- vec2 operators
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, const vec2& a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator/(const vec2& a, float s) { return vec2(a.x / s, a.y / s); }
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
// ---- vec3 operators
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
// ---- vec4 operators
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

// ---- scalar functions (GLM forwards to <cmath>)
template <class T> inline T pi() { return T(3.14159265358979323846264338327950288); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float sin(float x) { return std::sin(x); }
inline float cos(float x) { return std::cos(x); }
inline float tan(float x) { return std::tan(x); }
inline float asin(float x) { return std::asin(x); }
inline float acos(float x) { return std::acos(x); }
inline float atan(float y, float x) { return std::atan2(y, x); }
inline float exp(float x) { return std::exp(x); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline float abs(float x) { return std::fabs(x); }
inline bool isnan(float x) { return std::isnan(x); }
inline float degrees(float r) { return r * 57.295779513082320876798154814105f; }
// GLM: max(x,y) = (x < y) ? y : x ; min(x,y) = (y < x) ? y : x
template <class T> inline T max(T a, T b) { return (a < b) ? b : a; }
template <class T> inline T min(T a, T b) { return (b < a) ? b : a; }
// GLM: clamp = min(max(x, lo), hi)
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
// gtx/wrap.hpp: repeat(x) = fract(x) = x - floor(x)
inline float repeat(float x) { return x - std::floor(x); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }

inline vec3 abs(const vec3& v) { return vec3(std::fabs(v.x), std::fabs(v.y), std::fabs(v.z)); }

// ---- geometric (GLM compute_dot<vec3>: tmp = a*b; tmp.x + tmp.y + tmp.z)
inline float dot(const vec2& a, const vec2& b) { vec2 t(a.x * b.x, a.y * b.y); return t.x + t.y; }
inline float dot(const vec3& a, const vec3& b) { vec3 t(a * b); return t.x + t.y + t.z; }
inline float dot(const vec4& a, const vec4& b) { vec4 t(a * b); return (t.x + t.y) + (t.z + t.w); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline float length(const vec2& v) { return std::sqrt(dot(v, v)); }
inline float distance2(const vec3& a, const vec3& b) { vec3 d = a - b; return dot(d, d); } // gtx/norm: length2(a-b)
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
inline vec2 normalize(const vec2& v) { return v * inversesqrt(dot(v, v)); }
inline vec3 cross(const vec3& x, const vec3& y) {
    return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
// gtx/vector_angle.hpp: angle(x,y) = acos(clamp(dot(x,y), -1, 1))
inline float angle(const vec3& x, const vec3& y) { return std::acos(clamp(dot(x, y), -1.0f, 1.0f)); }

// ---- matrices (column major; m[c][r])
struct mat4;
struct mat3 {
    vec3 c[3];
    mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    template <class A> mat3(A x0, A y0, A z0, A x1, A y1, A z1, A x2, A y2, A z2) {
        c[0] = vec3(x0, y0, z0); c[1] = vec3(x1, y1, z1); c[2] = vec3(x2, y2, z2);
    }
    explicit mat3(const mat4& m);
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
    vec4 c[4];
    mat4() { c[0] = vec4(1, 0, 0, 0); c[1] = vec4(0, 1, 0, 0); c[2] = vec4(0, 0, 1, 0); c[3] = vec4(0, 0, 0, 1); }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline mat3::mat3(const mat4& m) { c[0] = vec3(m[0]); c[1] = vec3(m[1]); c[2] = vec3(m[2]); }

inline mat3 operator*(const mat3& m, float s) { return mat3(m[0] * s, m[1] * s, m[2] * s); }
inline mat3 operator+(const mat3& a, const mat3& b) { return mat3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline vec3 operator*(const mat3& m, const vec3& v) {
    return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
                m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
                m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}
inline mat3 operator*(const mat3& a, const mat3& b) {
    mat3 r;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++)
            r[j][i] = a[0][i] * b[j][0] + a[1][i] * b[j][1] + a[2][i] * b[j][2];
    return r;
}
// GLM 0.9.8 mat4*vec4: (m0*x + m1*y) + (m2*z + m3*w)
inline vec4 operator*(const mat4& m, const vec4& v) {
    vec4 Mul0 = m[0] * vec4(v[0]); vec4 Mul1 = m[1] * vec4(v[1]); vec4 Add0 = Mul0 + Mul1;
    vec4 Mul2 = m[2] * vec4(v[2]); vec4 Mul3 = m[3] * vec4(v[3]); vec4 Add1 = Mul2 + Mul3;
    return Add0 + Add1;
}
inline mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r;
    for (int j = 0; j < 4; j++) r[j] = a[0] * b[j][0] + a[1] * b[j][1] + a[2] * b[j][2] + a[3] * b[j][3];
    return r;
}
inline float determinant(const mat3& m) {
    return + m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
           - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
           + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
}
inline mat3 inverse(const mat3& m) {
    float OneOverDeterminant = 1.0f / (
        + m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
        - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
        + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
    mat3 I;
    I[0][0] = + (m[1][1] * m[2][2] - m[2][1] * m[1][2]) * OneOverDeterminant;
    I[1][0] = - (m[1][0] * m[2][2] - m[2][0] * m[1][2]) * OneOverDeterminant;
    I[2][0] = + (m[1][0] * m[2][1] - m[2][0] * m[1][1]) * OneOverDeterminant;
    I[0][1] = - (m[0][1] * m[2][2] - m[2][1] * m[0][2]) * OneOverDeterminant;
    I[1][1] = + (m[0][0] * m[2][2] - m[2][0] * m[0][2]) * OneOverDeterminant;
    I[2][1] = - (m[0][0] * m[2][1] - m[2][0] * m[0][1]) * OneOverDeterminant;
    I[0][2] = + (m[0][1] * m[1][2] - m[1][1] * m[0][2]) * OneOverDeterminant;
    I[1][2] = - (m[0][0] * m[1][2] - m[1][0] * m[0][2]) * OneOverDeterminant;
    I[2][2] = + (m[0][0] * m[1][1] - m[1][0] * m[0][1]) * OneOverDeterminant;
    return I;
}

// gtc/matrix_transform.hpp (applied to identity, as gtx/transform.hpp does)
inline mat4 translate(const vec3& v) {
    mat4 m, r;
    r[3] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3];
    return r;
}
inline mat4 scale(const vec3& v) {
    mat4 m, r;
    r[0] = m[0] * v[0]; r[1] = m[1] * v[1]; r[2] = m[2] * v[2]; r[3] = m[3];
    return r;
}
inline mat4 rotate(float angle_, const vec3& v) {
    mat4 m;
    const float a = angle_, c = std::cos(a), s = std::sin(a);
    vec3 axis(normalize(v));
    vec3 temp((1.0f - c) * axis);
    mat4 R;
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = temp[0] * axis[1] + s * axis[2];
    R[0][2] = temp[0] * axis[2] - s * axis[1];
    R[1][0] = temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = temp[1] * axis[2] + s * axis[0];
    R[2][0] = temp[2] * axis[0] + s * axis[1];
    R[2][1] = temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    mat4 r;
    r[0] = m[0] * R[0][0] + m[1] * R[0][1] + m[2] * R[0][2];
    r[1] = m[0] * R[1][0] + m[1] * R[1][1] + m[2] * R[1][2];
    r[2] = m[0] * R[2][0] + m[1] * R[2][1] + m[2] * R[2][2];
    r[3] = m[3];
    return r;
}
// gtx/rotate_vector.hpp: rotate(v, angle, normal) = mat3(glm::rotate(angle, normal)) * v
inline vec3 rotate(const vec3& v, float angle_, const vec3& normal) { return mat3(rotate(angle_, normal)) * v; }

// ---- quaternion (gtc/quaternion.hpp)
struct quat {
    float x, y, z, w;
    quat() : x(0), y(0), z(0), w(1) {}
    quat(float w_, float x_, float y_, float z_) : x(x_), y(y_), z(z_), w(w_) {}
};
inline float dot(const quat& a, const quat& b) {
    vec4 t(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
    return (t.x + t.y) + (t.z + t.w);
}
inline quat conjugate(const quat& q) { return quat(q.w, -q.x, -q.y, -q.z); }
inline quat operator/(const quat& q, float s) { return quat(q.w / s, q.x / s, q.y / s, q.z / s); }
inline quat inverse(const quat& q) { return conjugate(q) / dot(q, q); }
inline vec3 operator*(const quat& q, const vec3& v) {
    const vec3 QuatVector(q.x, q.y, q.z);
    const vec3 uv(cross(QuatVector, v));
    const vec3 uuv(cross(QuatVector, uv));
    return v + ((uv * q.w) + uuv) * 2.0f;
}
inline quat angleAxis(float angle_, const vec3& v) {
    const float a = angle_;
    const float s = std::sin(a * 0.5f);
    vec3 vs = v * s;
    return quat(std::cos(a * 0.5f), vs.x, vs.y, vs.z);
}

// gtx/string_cast.hpp (debug printing only)
inline std::string to_string(const vec3& v) {
    char b[128]; std::snprintf(b, sizeof b, "vec3(%f, %f, %f)", v.x, v.y, v.z); return b;
}
inline std::string to_string(const mat3& m) {
    return "mat3x3(" + to_string(m[0]) + ", " + to_string(m[1]) + ", " + to_string(m[2]) + ")";
}

} // namespace glm
