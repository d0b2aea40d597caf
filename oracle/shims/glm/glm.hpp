// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for glm 1.0.1 (pinned by the
// reference at cmake/glm.cmake:2,14), which is not vendored under /root/reference and
// cannot be fetched offline. Only what the reference's insert path uses is provided,
// with glm's GENERIC (scalar, non-SIMD) definitions -- the canonical arithmetic chosen
// in SURVEY.md section 8c:
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z        (glm/detail/func_geometric.inl compute_dot<vec<3>>)
//   length(v)     = sqrt(dot(v,v));  distance(a,b) = length(b - a)
//   normalize(v)  = v * inversesqrt(dot(v,v)),  inversesqrt(x) = T(1)/sqrt(x)
//   floor/abs     = per-component std::floor / std::abs;  sign(x) = (0<x)-(x<0)
// Unaligned (vec3/ivec3/dvec3) and aligned (aligned_vec3/aligned_ivec3) vectors are
// distinct types (the reference overloads on them, src/chad/tsdf.cpp:12-23) with implicit
// conversions in both directions, as in glm without GLM_FORCE_EXPLICIT_CTOR.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>

namespace glm {
template <typename T, bool Aligned>
struct tvec3 {
    T x, y, z;
    constexpr tvec3() : x(0), y(0), z(0) {}
    constexpr explicit tvec3(T s) : x(s), y(s), z(s) {}
    template <typename A, typename B, typename C>
    constexpr tvec3(A a, B b, C c) : x(T(a)), y(T(b)), z(T(c)) {}
    template <typename U, bool Q>
    constexpr tvec3(const tvec3<U, Q>& v) : x(T(v.x)), y(T(v.y)), z(T(v.z)) {}

    tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    tvec3& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
    tvec3 operator-() const { return tvec3(-x, -y, -z); }
};

using vec3 = tvec3<float, false>;
using dvec3 = tvec3<double, false>;
using ivec3 = tvec3<int32_t, false>;
using aligned_vec3 = tvec3<float, true>;
using aligned_ivec3 = tvec3<int32_t, true>;

// vec (op) vec: result takes the left operand's alignment qualifier
template <typename T, bool A, bool B>
inline tvec3<T, A> operator+(const tvec3<T, A>& a, const tvec3<T, B>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T, bool A, bool B>
inline tvec3<T, A> operator-(const tvec3<T, A>& a, const tvec3<T, B>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T, bool A, bool B>
inline tvec3<T, A> operator*(const tvec3<T, A>& a, const tvec3<T, B>& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
// vec (op) scalar, scalar (op) vec
template <typename T, bool A>
inline tvec3<T, A> operator*(const tvec3<T, A>& a, T s) { return {a.x * s, a.y * s, a.z * s}; }
template <typename T, bool A>
inline tvec3<T, A> operator*(T s, const tvec3<T, A>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <typename T, bool A>
inline tvec3<T, A> operator/(T s, const tvec3<T, A>& a) { return {s / a.x, s / a.y, s / a.z}; }
template <typename T, bool A>
inline tvec3<T, A> operator/(const tvec3<T, A>& a, T s) { return {a.x / s, a.y / s, a.z / s}; }

template <typename T, bool A, bool B>
inline T dot(const tvec3<T, A>& a, const tvec3<T, B>& b) {
    T tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return (tx + ty) + tz;
}
template <typename T, bool A>
inline T length(const tvec3<T, A>& v) { return std::sqrt(dot(v, v)); }
template <typename T, bool A, bool B>
inline T distance(const tvec3<T, A>& p0, const tvec3<T, B>& p1) { return length(p1 - p0); }
template <typename T, bool A>
inline tvec3<T, A> normalize(const tvec3<T, A>& v) { return v * (T(1) / std::sqrt(dot(v, v))); }
template <typename T, bool A>
inline tvec3<T, A> floor(const tvec3<T, A>& v) { return {std::floor(v.x), std::floor(v.y), std::floor(v.z)}; }
template <typename T, bool A>
inline tvec3<T, A> abs(const tvec3<T, A>& v) { return {std::abs(v.x), std::abs(v.y), std::abs(v.z)}; }
template <typename T, bool A>
inline tvec3<T, A> sign(const tvec3<T, A>& v) {
    return {T((T(0) < v.x) - (v.x < T(0))), T((T(0) < v.y) - (v.y < T(0))), T((T(0) < v.z) - (v.z < T(0)))};
}
}  // namespace glm
