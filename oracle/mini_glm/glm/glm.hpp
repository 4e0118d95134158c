// mini_glm -- the sliver of GLM that tinynurbs' curve evaluation needs (TEST INFRASTRUCTURE ONLY).
// GLM is not installed in this image and not vendored by the reference; tinynurbs (vendored under
// /root/reference/quadruped/extern/tinynurbs) only uses glm::vec<N,T> as a plain N-vector with component-wise
// + - and scalar * /.  Arithmetic is one IEEE operation per component, like GLM's scalar (non-SIMD) path.
#pragma once
#include <cmath>
#include <cstddef>

namespace glm {
template <int N, typename T>
struct vec_storage { T v[N]; };
template <typename T> struct vec_storage<1, T> { union { T v[1]; struct { T x; }; }; };
template <typename T> struct vec_storage<2, T> { union { T v[2]; struct { T x, y; }; }; };
template <typename T> struct vec_storage<3, T> { union { T v[3]; struct { T x, y, z; }; }; };
template <typename T> struct vec_storage<4, T> { union { T v[4]; struct { T x, y, z, w; }; }; };

template <int N, typename T>
struct vec : vec_storage<N, T> {
    using vec_storage<N, T>::v;
    vec() { for (int i = 0; i < N; ++i) v[i] = T(0); }
    explicit vec(T s) { for (int i = 0; i < N; ++i) v[i] = s; }
    vec(T a, T b) { static_assert(N == 2, "size"); v[0] = a; v[1] = b; }
    vec(T a, T b, T c) { static_assert(N == 3, "size"); v[0] = a; v[1] = b; v[2] = c; }
    vec(T a, T b, T c, T d) { static_assert(N == 4, "size"); v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
    template <int M> explicit vec(const vec<M, T>& o) { for (int i = 0; i < N; ++i) v[i] = i < M ? o.v[i] : T(0); }
    template <int M> vec(const vec<M, T>& o, T last) { for (int i = 0; i < N; ++i) v[i] = i < M ? o.v[i] : last; }
    static constexpr int length() { return N; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    vec& operator+=(const vec& o) { for (int i = 0; i < N; ++i) v[i] = v[i] + o.v[i]; return *this; }
    vec& operator-=(const vec& o) { for (int i = 0; i < N; ++i) v[i] = v[i] - o.v[i]; return *this; }
    vec& operator*=(T s) { for (int i = 0; i < N; ++i) v[i] = v[i] * s; return *this; }
    vec& operator/=(T s) { for (int i = 0; i < N; ++i) v[i] = v[i] / s; return *this; }
};
template <int N, typename T> vec<N, T> operator+(vec<N, T> a, const vec<N, T>& b) { a += b; return a; }
template <int N, typename T> vec<N, T> operator-(vec<N, T> a, const vec<N, T>& b) { a -= b; return a; }
template <int N, typename T> vec<N, T> operator-(vec<N, T> a) { for (int i = 0; i < N; ++i) a.v[i] = -a.v[i]; return a; }
template <int N, typename T> vec<N, T> operator*(T s, vec<N, T> a) { for (int i = 0; i < N; ++i) a.v[i] = s * a.v[i]; return a; }
template <int N, typename T> vec<N, T> operator*(vec<N, T> a, T s) { a *= s; return a; }
template <int N, typename T> vec<N, T> operator/(vec<N, T> a, T s) { a /= s; return a; }
template <int N, typename T> T length(const vec<N, T>& a) { T s = T(0); for (int i = 0; i < N; ++i) s += a.v[i] * a.v[i]; return std::sqrt(s); }
template <typename T> vec<3, T> cross(const vec<3, T>& a, const vec<3, T>& b) {
    return vec<3, T>(a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]);
}
typedef vec<3, float> vec3;
typedef vec<4, float> vec4;
typedef vec<2, float> vec2;
typedef vec<3, double> dvec3;
}  // namespace glm
