// TEST INFRASTRUCTURE — not product code.
//
// Minimal stand-in for the parts of glm the reference's simulation step uses, so that the
// reference sources (src/sph.cpp, src/neighborTable.cpp, src/SPHSystem.cpp, src/Geometry.cpp)
// can be compiled headless, unmodified, in a container that has no glm installed
// (the reference finds glm with find_package and pins no version, CMakeLists.txt:8).
//
// Written from scratch. Only the arithmetic order matters for parity, and it follows
// glm 0.9.9's scalar (non-SIMD) code paths:
//   dot(vec3)      = (a.x*b.x + a.y*b.y) + a.z*b.z      three rounded products, two rounded adds
//   length2(v)     = dot(v, v)                            (glm/gtx/norm.hpp)
//   inversesqrt(x) = 1 / sqrt(x)
//   normalize(v)   = v * inversesqrt(dot(v, v))           one divide, three multiplies
//   vec op scalar  = component-wise in float
//   vec *= U       = each component *= static_cast<float>(U)
//   glm::pow(float, int) is std::pow (pulled in with a using-declaration), i.e. double
//   mat4 is column-major, mat4 * mat4 sums the four column products left to right
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>

// The same header also serves the reference's CUDA file (src/kernels/sphGPU.cu, compiled only as a
// secondary same-box baseline): under nvcc every function is host + device.
#ifdef __CUDACC__
#define GLM_SHIM_HD __host__ __device__
#else
#define GLM_SHIM_HD
#endif

namespace glm {

using std::pow;
using std::sqrt;

template <typename T>
struct tvec3 {
    T x, y, z;

    tvec3() = default;
    template <typename S>
    GLM_SHIM_HD explicit tvec3(S s) : x(static_cast<T>(s)), y(static_cast<T>(s)), z(static_cast<T>(s)) {}
    template <typename A, typename B, typename C>
    GLM_SHIM_HD tvec3(A a, B b, C c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
    template <typename U>
    GLM_SHIM_HD tvec3(const tvec3<U> &v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}

    GLM_SHIM_HD T &operator[](int i) { return (&x)[i]; }
    GLM_SHIM_HD const T &operator[](int i) const { return (&x)[i]; }

    GLM_SHIM_HD tvec3 &operator+=(const tvec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLM_SHIM_HD tvec3 &operator-=(const tvec3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    template <typename U>
    GLM_SHIM_HD tvec3 &operator*=(U s) {
        x *= static_cast<T>(s); y *= static_cast<T>(s); z *= static_cast<T>(s);
        return *this;
    }
    template <typename U>
    GLM_SHIM_HD tvec3 &operator/=(U s) {
        x /= static_cast<T>(s); y /= static_cast<T>(s); z /= static_cast<T>(s);
        return *this;
    }
};

template <typename T> GLM_SHIM_HD inline tvec3<T> operator-(const tvec3<T> &v) { return tvec3<T>(-v.x, -v.y, -v.z); }
template <typename T> GLM_SHIM_HD inline tvec3<T> operator+(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> GLM_SHIM_HD inline tvec3<T> operator-(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> GLM_SHIM_HD inline tvec3<T> operator*(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T> GLM_SHIM_HD inline tvec3<T> operator*(const tvec3<T> &v, T s) { return tvec3<T>(v.x * s, v.y * s, v.z * s); }
template <typename T> GLM_SHIM_HD inline tvec3<T> operator*(T s, const tvec3<T> &v) { return tvec3<T>(s * v.x, s * v.y, s * v.z); }
template <typename T> GLM_SHIM_HD inline tvec3<T> operator/(const tvec3<T> &v, T s) { return tvec3<T>(v.x / s, v.y / s, v.z / s); }

typedef tvec3<float> vec3;
typedef tvec3<int> ivec3;
typedef tvec3<unsigned int> uvec3;

struct vec4 {
    float x, y, z, w;
    vec4() = default;
    GLM_SHIM_HD vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    GLM_SHIM_HD vec4(const vec3 &v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    GLM_SHIM_HD float &operator[](int i) { return (&x)[i]; }
    GLM_SHIM_HD const float &operator[](int i) const { return (&x)[i]; }
};
GLM_SHIM_HD inline vec4 operator*(const vec4 &v, float s) { return vec4(v.x * s, v.y * s, v.z * s, v.w * s); }
GLM_SHIM_HD inline vec4 operator+(const vec4 &a, const vec4 &b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// Column-major 4x4: col[c][r].
struct mat4 {
    vec4 col[4];
    mat4() = default;
    GLM_SHIM_HD explicit mat4(float d) {
        col[0] = vec4(d, 0, 0, 0); col[1] = vec4(0, d, 0, 0);
        col[2] = vec4(0, 0, d, 0); col[3] = vec4(0, 0, 0, d);
    }
    GLM_SHIM_HD vec4 &operator[](int i) { return col[i]; }
    GLM_SHIM_HD const vec4 &operator[](int i) const { return col[i]; }
};
GLM_SHIM_HD inline mat4 operator*(const mat4 &a, const mat4 &b) {
    mat4 r;
    for (int c = 0; c < 4; ++c)
        r[c] = a[0] * b[c][0] + a[1] * b[c][1] + a[2] * b[c][2] + a[3] * b[c][3];
    return r;
}

GLM_SHIM_HD inline float dot(const vec3 &a, const vec3 &b) {
    vec3 t(a * b);
    return t.x + t.y + t.z;
}
GLM_SHIM_HD inline float length(const vec3 &v) { return std::sqrt(dot(v, v)); }
GLM_SHIM_HD inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
GLM_SHIM_HD inline vec3 normalize(const vec3 &v) { return v * inversesqrt(dot(v, v)); }

} // namespace glm
