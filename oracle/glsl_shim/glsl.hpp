// glsl.hpp -- GLSL compute-shader compatibility layer for g++ (TEST INFRASTRUCTURE, oracle/_ref).
//
// Lets the reference's shader SOURCE TEXT (shaders/*.comp, read where it lies under /root/reference and passed
// through the mechanical rewrites of translate.py) compile as C++ and run on the CPU, so that the hand-written
// restatement in oracle/vkpbrt_oracle.c can be checked against the reference's own code.  This header plays the
// role of the Vulkan implementation and fixes, identically to the oracle, everything the Vulkan / GLSL specs leave
// implementation-defined: fp16 stores round to nearest even, unorm8 stores (uint8)(c*255+.5),
// bilinear REPEAT sampling with exact fp32 weights, out-of-bounds image loads return 0, subgroup size 32 with
// subgroupAdd as the xor-butterfly tree, sin / cos / pow as the fixed Cephes-style algorithms, no FMA contraction
// (build flag -ffp-contract=off).  Nothing here knows what the shaders compute.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <type_traits>

namespace glsl {

typedef unsigned int uint;

// ---- vectors with the swizzles the path's shaders use (.xy, .xyz; single components are plain members) ----------
template <class T, int N> struct vec;
template <class T, int N> struct Swz2 {
    T d[N];
    operator vec<T, 2>() const;
    Swz2& operator=(const Swz2& o) { d[0] = o.d[0]; d[1] = o.d[1]; return *this; }
    template <class V> Swz2& operator=(const V& v);
    template <class V> Swz2& operator+=(const V& v);
    template <class V> Swz2& operator-=(const V& v);
    template <class V> Swz2& operator*=(const V& v);
    template <class V> Swz2& operator/=(const V& v);
};
template <class T, int N> struct Swz3 {
    T d[N];
    operator vec<T, 3>() const;
    Swz3& operator=(const Swz3& o) { d[0] = o.d[0]; d[1] = o.d[1]; d[2] = o.d[2]; return *this; }
    template <class V> Swz3& operator=(const V& v);
};

template <class T> struct vec<T, 2> {
    union { struct { T x, y; }; T data[2]; Swz2<T, 2> xy; };
    vec() : x(), y() {}
    vec(const vec& o) : x(o.x), y(o.y) {}
    vec& operator=(const vec& o) { x = o.x; y = o.y; return *this; }
    template <class A, std::enable_if_t<std::is_arithmetic<A>::value, int> = 0> explicit vec(A s) : x((T)s), y((T)s) {}
    template <class A, class B, std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value, int> = 0>
    vec(A a, B b) : x((T)a), y((T)b) {}
    template <class U> vec(const vec<U, 2>& o) : x((T)o.x), y((T)o.y) {}                            // implicit int <-> float conversion
    template <class U, int M, std::enable_if_t<(M > 2), int> = 0> explicit vec(const vec<U, M>& o) : x((T)o.data[0]), y((T)o.data[1]) {}   // ivec2(uvec3): explicit truncation
    template <class U, int M> vec(const Swz2<U, M>& o) : x((T)o.d[0]), y((T)o.d[1]) {}
    T& operator[](int i) { return data[i]; }
    const T& operator[](int i) const { return data[i]; }
};
template <class T> struct vec<T, 3> {
    union { struct { T x, y, z; }; T data[3]; Swz2<T, 3> xy; Swz3<T, 3> xyz; };
    vec() : x(), y(), z() {}
    vec(const vec& o) : x(o.x), y(o.y), z(o.z) {}
    vec& operator=(const vec& o) { x = o.x; y = o.y; z = o.z; return *this; }
    template <class A, std::enable_if_t<std::is_arithmetic<A>::value, int> = 0> explicit vec(A s) : x((T)s), y((T)s), z((T)s) {}
    template <class A, class B, class C, std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<C>::value, int> = 0>
    vec(A a, B b, C c) : x((T)a), y((T)b), z((T)c) {}
    template <class U, class C, std::enable_if_t<std::is_arithmetic<C>::value, int> = 0> vec(const vec<U, 2>& a, C c) : x((T)a.x), y((T)a.y), z((T)c) {}
    template <class U, int M, class C, std::enable_if_t<std::is_arithmetic<C>::value, int> = 0> vec(const Swz2<U, M>& a, C c) : x((T)a.d[0]), y((T)a.d[1]), z((T)c) {}
    template <class U> vec(const vec<U, 3>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    template <class U> explicit vec(const vec<U, 4>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    template <class U, int M> vec(const Swz3<U, M>& o) : x((T)o.d[0]), y((T)o.d[1]), z((T)o.d[2]) {}
    T& operator[](int i) { return data[i]; }
    const T& operator[](int i) const { return data[i]; }
};
template <class T> struct vec<T, 4> {
    union { struct { T x, y, z, w; }; T data[4]; Swz2<T, 4> xy; Swz3<T, 4> xyz; };
    vec() : x(), y(), z(), w() {}
    vec(const vec& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec& operator=(const vec& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    template <class A, std::enable_if_t<std::is_arithmetic<A>::value, int> = 0> explicit vec(A s) : x((T)s), y((T)s), z((T)s), w((T)s) {}
    template <class A, class B, class C, class D, std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<D>::value, int> = 0>
    vec(A a, B b, C c, D d) : x((T)a), y((T)b), z((T)c), w((T)d) {}
    template <class U, class D, std::enable_if_t<std::is_arithmetic<D>::value, int> = 0> vec(const vec<U, 3>& a, D d) : x((T)a.x), y((T)a.y), z((T)a.z), w((T)d) {}
    template <class U, int M, class D, std::enable_if_t<std::is_arithmetic<D>::value, int> = 0> vec(const Swz3<U, M>& a, D d) : x((T)a.d[0]), y((T)a.d[1]), z((T)a.d[2]), w((T)d) {}
    template <class U, class C, class D, std::enable_if_t<std::is_arithmetic<D>::value, int> = 0> vec(const vec<U, 2>& a, C c, D d) : x((T)a.x), y((T)a.y), z((T)c), w((T)d) {}
    template <class U, int M, class C, class D, std::enable_if_t<std::is_arithmetic<D>::value, int> = 0> vec(const Swz2<U, M>& a, C c, D d) : x((T)a.d[0]), y((T)a.d[1]), z((T)c), w((T)d) {}
    template <class U> vec(const vec<U, 4>& o) : x((T)o.x), y((T)o.y), z((T)o.z), w((T)o.w) {}
    T& operator[](int i) { return data[i]; }
    const T& operator[](int i) const { return data[i]; }
};
typedef vec<float, 2> vec2; typedef vec<float, 3> vec3; typedef vec<float, 4> vec4;
typedef vec<int, 2> ivec2; typedef vec<int, 3> ivec3;
typedef vec<uint, 2> uvec2; typedef vec<uint, 3> uvec3;
typedef vec<bool, 2> bvec2; typedef vec<bool, 3> bvec3;

template <class T, int N> Swz2<T, N>::operator vec<T, 2>() const { return vec<T, 2>(d[0], d[1]); }
template <class T, int N> Swz3<T, N>::operator vec<T, 3>() const { return vec<T, 3>(d[0], d[1], d[2]); }

// unwrap: swizzle proxies decay to vectors, everything else passes through
template <class X> struct is_vec : std::false_type {};
template <class T, int N> struct is_vec<vec<T, N>> : std::true_type {};
template <class X> struct is_swz : std::false_type {};
template <class T, int N> struct is_swz<Swz2<T, N>> : std::true_type {};
template <class T, int N> struct is_swz<Swz3<T, N>> : std::true_type {};
template <class X> struct is_vecish : std::integral_constant<bool, is_vec<X>::value || is_swz<X>::value> {};
template <class T, int N> const vec<T, N>& unwrap(const vec<T, N>& v) { return v; }
template <class T, int N> vec<T, 2> unwrap(const Swz2<T, N>& v) { return vec<T, 2>(v); }
template <class T, int N> vec<T, 3> unwrap(const Swz3<T, N>& v) { return vec<T, 3>(v); }
template <class A, std::enable_if_t<std::is_arithmetic<A>::value, int> = 0> A unwrap(A a) { return a; }

template <class T, class U, int N, class F> auto zip(const vec<T, N>& a, const vec<U, N>& b, F f)
{
    vec<decltype(f(a.data[0], b.data[0])), N> r;
    for (int i = 0; i < N; ++i) r.data[i] = f(a.data[i], b.data[i]);
    return r;
}
template <class T, class S, int N, class F, std::enable_if_t<std::is_arithmetic<S>::value, int> = 0> auto zip(const vec<T, N>& a, S b, F f)
{
    vec<decltype(f(a.data[0], b)), N> r;
    for (int i = 0; i < N; ++i) r.data[i] = f(a.data[i], b);
    return r;
}
template <class S, class T, int N, class F, std::enable_if_t<std::is_arithmetic<S>::value, int> = 0> auto zip(S a, const vec<T, N>& b, F f)
{
    vec<decltype(f(a, b.data[0])), N> r;
    for (int i = 0; i < N; ++i) r.data[i] = f(a, b.data[i]);
    return r;
}
#define GLSL_BINOP(op)                                                                                          \
    template <class A, class B, std::enable_if_t<is_vecish<A>::value || is_vecish<B>::value, int> = 0>           \
    auto operator op(const A& a, const B& b) { return zip(unwrap(a), unwrap(b), [](auto x, auto y) { return x op y; }); }
GLSL_BINOP(+) GLSL_BINOP(-) GLSL_BINOP(*) GLSL_BINOP(/)
#undef GLSL_BINOP
template <class T, int N> vec<T, N> operator-(const vec<T, N>& a) { vec<T, N> r; for (int i = 0; i < N; ++i) r.data[i] = -a.data[i]; return r; }
#define GLSL_ASSIGNOP(op)                                                                                       \
    template <class T, int N, class B> vec<T, N>& operator op##=(vec<T, N>& a, const B& b) { a = vec<T, N>(a op b); return a; }
GLSL_ASSIGNOP(+) GLSL_ASSIGNOP(-) GLSL_ASSIGNOP(*) GLSL_ASSIGNOP(/)
#undef GLSL_ASSIGNOP
template <class T, class U, int N> bool operator==(const vec<T, N>& a, const vec<U, N>& b) { for (int i = 0; i < N; ++i) if (!(a.data[i] == b.data[i])) return false; return true; }
template <class T, class U, int N> bool operator!=(const vec<T, N>& a, const vec<U, N>& b) { return !(a == b); }

template <class T, int N> template <class V> Swz2<T, N>& Swz2<T, N>::operator=(const V& v) { vec<T, 2> t(unwrap(v)); d[0] = t.x; d[1] = t.y; return *this; }
template <class T, int N> template <class V> Swz2<T, N>& Swz2<T, N>::operator+=(const V& v) { return *this = vec<T, 2>(*this) + v; }
template <class T, int N> template <class V> Swz2<T, N>& Swz2<T, N>::operator-=(const V& v) { return *this = vec<T, 2>(*this) - v; }
template <class T, int N> template <class V> Swz2<T, N>& Swz2<T, N>::operator*=(const V& v) { return *this = vec<T, 2>(*this) * v; }
template <class T, int N> template <class V> Swz2<T, N>& Swz2<T, N>::operator/=(const V& v) { return *this = vec<T, 2>(*this) / v; }
template <class T, int N> template <class V> Swz3<T, N>& Swz3<T, N>::operator=(const V& v) { vec<T, 3> t(unwrap(v)); d[0] = t.x; d[1] = t.y; d[2] = t.z; return *this; }

// array constructor "x = float[N](a, b, ...)" (translate.py rewrites it to assign(x, a, b, ...))
template <class T, int N, class... A> void assign(T (&dst)[N], A... a)
{
    static_assert(sizeof...(A) == N, "array constructor size mismatch");
    T tmp[N] = {(T)a...};
    for (int i = 0; i < N; ++i) dst[i] = tmp[i];
}

// ---- scalar built-ins (GLSL definitions; NaN behaviour of min / max follows the spec's ternaries) ----------------
inline float abs(float x) { return fabsf(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float min(float x, float y) { return (y < x) ? y : x; }
inline float max(float x, float y) { return (x < y) ? y : x; }
inline int min(int x, int y) { return (y < x) ? y : x; }
inline float min(float x, int y) { return min(x, (float)y); }
inline float min(int x, float y) { return min((float)x, y); }
inline float max(float x, int y);
inline float max(int x, float y);
inline int max(int x, int y) { return (x < y) ? y : x; }
inline float max(float x, int y) { return max(x, (float)y); }
inline float max(int x, float y) { return max((float)x, y); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float floor(float x) { return floorf(x); }
inline float sqrt(float x) { return sqrtf(x); }
inline bool isnan(float x) { return x != x; }
inline bool isinf(float x) { return x == INFINITY || x == -INFINITY; }
// exp / pow(x, int): evaluated in double and rounded once (the oracle's definition of the BFR schedule terms)
inline float exp(float x) { return (float)::exp((double)x); }
inline float pow(float b, int e) { return (float)::pow((double)b, (double)e); }

inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// deterministic sin / cos / pow(x, y): same algorithms as oracle/vkpbrt_oracle.c (vk_sincos, vk_pow)
inline void sincos_det(float x, float* sn, float* cs)
{
    float ax = fabsf(x);
    if (!(ax < 8192.0f)) { *sn = *cs = (x - x) / (x - x); return; }
    uint32_t j = (uint32_t)(ax * 1.27323954473516f);
    float y = (float)j;
    if (j & 1u) { j += 1u; y += 1.0f; }
    float r = ((ax - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
    float z = r * r;
    float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z + -1.6666654611e-1f) * z * r + r;
    float pc = ((2.443315711809948e-5f * z + -1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z;
    pc = pc - 0.5f * z;
    pc = pc + 1.0f;
    uint32_t q = (j >> 1) & 3u;
    float s = (q == 0u) ? ps : ((q == 1u) ? pc : ((q == 2u) ? -ps : -pc));
    float c = (q == 0u) ? pc : ((q == 1u) ? -ps : ((q == 2u) ? -pc : ps));
    *sn = (x < 0.0f) ? -s : s;
    *cs = c;
}
inline float sin(float x) { float s, c; sincos_det(x, &s, &c); return s; }
inline float cos(float x) { float s, c; sincos_det(x, &s, &c); return c; }
inline float pow(float x, float y)
{
    if (!(x > 0.0f)) return (x == 0.0f) ? 0.0f : (x - x) / (x - x);
    if (x > 3.0e38f) return x;
    int e = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
    uint32_t bits = f2u(x);
    e += (int)((bits >> 23) & 255u) - 127;
    float m = u2f((bits & 0x007fffffu) | 0x3f800000u);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    float z = (m - 1.0f) / (m + 1.0f);
    float z2 = z * z;
    float p = ((((0.0909090909f * z2 + 0.1111111111f) * z2 + 0.1428571429f) * z2 + 0.2f) * z2 + 0.3333333333f) * z2;
    float lnm = 2.0f * (z + z * p);
    float lg = (float)e + lnm * 1.44269504089f;
    float t = y * lg;
    if (t > 127.99f) return u2f(0x7f800000u);
    if (t < -150.0f) return 0.0f;
    float n = rintf(t);
    float f = t - n;
    float px = ((((1.535336188319500e-4f * f + 1.339887440266574e-3f) * f + 9.618437357674640e-3f) * f + 5.550332471162809e-2f) * f + 2.402264791363012e-1f) * f + 6.931472028550421e-1f;
    float r = 1.0f + f * px;
    int ni = (int)n, n1 = ni / 2, n2 = ni - n1;
    return (r * u2f((uint32_t)(n1 + 127) << 23)) * u2f((uint32_t)(n2 + 127) << 23);
}

// ---- vector built-ins -----------------------------------------------------------------------------------------
#define GLSL_MAP1(fn) template <int N> vec<float, N> fn(const vec<float, N>& a) { vec<float, N> r; for (int i = 0; i < N; ++i) r.data[i] = fn(a.data[i]); return r; }
GLSL_MAP1(abs) GLSL_MAP1(sign) GLSL_MAP1(sqrt) GLSL_MAP1(floor)
#undef GLSL_MAP1
template <int N> vec<float, N> min(const vec<float, N>& a, const vec<float, N>& b) { return zip(a, b, [](float x, float y) { return min(x, y); }); }
template <int N> vec<float, N> max(const vec<float, N>& a, const vec<float, N>& b) { return zip(a, b, [](float x, float y) { return max(x, y); }); }
template <int N> vec<float, N> pow(const vec<float, N>& a, const vec<float, N>& b) { return zip(a, b, [](float x, float y) { return pow(x, y); }); }
// pow(abs(delta), ivec3(2)) in bfr.comp:132: an integer power of two is the square
template <int N> vec<float, N> pow(const vec<float, N>& a, const vec<int, N>& b) { vec<float, N> r; for (int i = 0; i < N; ++i) r.data[i] = (b.data[i] == 2) ? a.data[i] * a.data[i] : pow(a.data[i], (float)b.data[i]); return r; }
template <int N, class A, class B> vec<float, N> clamp(const vec<float, N>& x, A lo, B hi) { vec<float, N> r; for (int i = 0; i < N; ++i) r.data[i] = clamp(x.data[i], (float)lo, (float)hi); return r; }
template <int N> vec<float, N> clamp(const vec<float, N>& x, const vec<float, N>& lo, const vec<float, N>& hi) { vec<float, N> r; for (int i = 0; i < N; ++i) r.data[i] = clamp(x.data[i], lo.data[i], hi.data[i]); return r; }
template <int N> vec<float, N> mix(const vec<float, N>& x, const vec<float, N>& y, float a) { vec<float, N> r; for (int i = 0; i < N; ++i) r.data[i] = mix(x.data[i], y.data[i], a); return r; }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline float length(const vec3& a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(const vec3& a) { float inv = 1.0f / sqrtf(dot(a, a)); return vec3(a.x * inv, a.y * inv, a.z * inv); }
inline vec4 normalize(const vec4& a) { float inv = 1.0f / sqrtf(dot(a, a)); return vec4(a.x * inv, a.y * inv, a.z * inv, a.w * inv); }
inline bvec2 greaterThanEqual(const vec2& a, const vec2& b) { return bvec2(a.x >= b.x, a.y >= b.y); }
inline bvec2 lessThanEqual(const vec2& a, const vec2& b) { return bvec2(a.x <= b.x, a.y <= b.y); }
inline bvec2 equal(const ivec2& a, const ivec2& b) { return bvec2(a.x == b.x, a.y == b.y); }
inline bool all(const bvec2& b) { return b.x && b.y; }
inline bvec3 greaterThanEqual(const vec3& a, const vec3& b) { return bvec3(a.x >= b.x, a.y >= b.y, a.z >= b.z); }
inline bvec3 lessThanEqual(const vec3& a, const vec3& b) { return bvec3(a.x <= b.x, a.y <= b.y, a.z <= b.z); }
inline bool all(const bvec3& b) { return b.x && b.y && b.z; }

struct mat4 {
    vec4 c[4];
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4& m, const vec4& v)
{
    vec4 r;
    for (int i = 0; i < 4; ++i) r.data[i] = ((m.c[0].data[i] * v.x + m.c[1].data[i] * v.y) + m.c[2].data[i] * v.z) + m.c[3].data[i] * v.w;
    return r;
}
inline mat4 operator*(const mat4& a, const mat4& b) { mat4 r; for (int i = 0; i < 4; ++i) r.c[i] = a * b.c[i]; return r; }
inline mat4 inverse(const mat4& mm)   // cofactor expansion, the order shared with the oracle and the CUDA host side
{
    const float* m = &mm.c[0].x;
    float inv[16];
    float a00 = m[0], a01 = m[1], a02 = m[2], a03 = m[3], a10 = m[4], a11 = m[5], a12 = m[6], a13 = m[7];
    float a20 = m[8], a21 = m[9], a22 = m[10], a23 = m[11], a30 = m[12], a31 = m[13], a32 = m[14], a33 = m[15];
    float b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10, b02 = a00 * a13 - a03 * a10, b03 = a01 * a12 - a02 * a11;
    float b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12, b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30;
    float b08 = a20 * a33 - a23 * a30, b09 = a21 * a32 - a22 * a31, b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;
    float det = ((((b00 * b11 - b01 * b10) + b02 * b09) + b03 * b08) - b04 * b07) + b05 * b06;
    float id = 1.0f / det;
    inv[0] = ((a11 * b11 - a12 * b10) + a13 * b09) * id;  inv[1] = ((a02 * b10 - a01 * b11) - a03 * b09) * id;
    inv[2] = ((a31 * b05 - a32 * b04) + a33 * b03) * id;  inv[3] = ((a22 * b04 - a21 * b05) - a23 * b03) * id;
    inv[4] = ((a12 * b08 - a10 * b11) - a13 * b07) * id;  inv[5] = ((a00 * b11 - a02 * b08) + a03 * b07) * id;
    inv[6] = ((a32 * b02 - a30 * b05) - a33 * b01) * id;  inv[7] = ((a20 * b05 - a22 * b02) + a23 * b01) * id;
    inv[8] = ((a10 * b10 - a11 * b08) + a13 * b06) * id;  inv[9] = ((a01 * b08 - a00 * b10) - a03 * b06) * id;
    inv[10] = ((a30 * b04 - a31 * b02) + a33 * b00) * id; inv[11] = ((a21 * b02 - a20 * b04) - a23 * b00) * id;
    inv[12] = ((a11 * b07 - a10 * b09) - a12 * b06) * id; inv[13] = ((a00 * b09 - a01 * b07) + a02 * b06) * id;
    inv[14] = ((a31 * b01 - a30 * b03) - a32 * b00) * id; inv[15] = ((a20 * b03 - a21 * b01) + a22 * b00) * id;
    mat4 r;
    memcpy(&r.c[0].x, inv, sizeof inv);
    return r;
}

// ---- images and samplers ----------------------------------------------------------------------------------------
enum Format { F_R32F = 1, F_RG32F = 2, F_RGBA8 = 3, F_BGRA8 = 4, F_RG16F = 5, F_R8 = 6, F_RGBA16F = 7, F_RGBA32F = 8, F_R16F = 9 };
struct Binding { void* data; int width, height, layers, format; };
extern Binding g_bindings[32];

inline uint16_t f32_to_f16(float f)
{
    uint32_t x = f2u(f), sign = (x >> 16) & 0x8000u, ax = x & 0x7fffffffu;
    if (ax > 0x7f800000u) return (uint16_t)(sign | 0x7fffu);
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);
    if (ax < 0x38800000u) {
        if (ax <= 0x33000000u) return (uint16_t)sign;
        uint32_t e = ax >> 23, m = (ax & 0x7fffffu) | 0x800000u, s = 126u - e, h = m >> s, rem = m & ((1u << s) - 1u), half = 1u << (s - 1u);
        if (rem > half || (rem == half && (h & 1u))) h++;
        return (uint16_t)(sign | h);
    }
    uint32_t e = (ax >> 23) - 112u, m = ax & 0x7fffffu, h = (e << 10) | (m >> 13), rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return (uint16_t)(sign | h);
}
inline float f16_to_f32(uint16_t h)
{
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    if (e == 0) { float v = (float)m * 5.9604644775390625e-08f; return sign ? -v : v; }
    if (e == 31) return u2f(sign | 0x7f800000u | (m << 13));
    return u2f(sign | ((e + 112u) << 23) | (m << 13));
}
inline uint8_t f32_to_unorm8(float c)
{
    if (!(c == c)) return 0;
    if (c < 0.0f) c = 0.0f;
    if (c > 1.0f) c = 1.0f;
    return (uint8_t)(c * 255.0f + 0.5f);
}

inline vec4 load_texel(const Binding& b, int x, int y, int layer)
{
    if (x < 0 || y < 0 || x >= b.width || y >= b.height || layer < 0 || layer >= b.layers) return vec4(0.0f);
    size_t i = ((size_t)layer * b.height + y) * b.width + x;
    switch (b.format) {
    case F_R32F: return vec4(((const float*)b.data)[i], 0.0f, 0.0f, 1.0f);
    case F_RG32F: return vec4(((const float*)b.data)[2 * i], ((const float*)b.data)[2 * i + 1], 0.0f, 1.0f);
    case F_RGBA32F: { const float* p = (const float*)b.data + 4 * i; return vec4(p[0], p[1], p[2], p[3]); }
    case F_R16F: return vec4(f16_to_f32(((const uint16_t*)b.data)[i]), 0.0f, 0.0f, 1.0f);
    case F_RG16F: { const uint16_t* p = (const uint16_t*)b.data + 2 * i; return vec4(f16_to_f32(p[0]), f16_to_f32(p[1]), 0.0f, 1.0f); }
    case F_RGBA16F: { const uint16_t* p = (const uint16_t*)b.data + 4 * i; return vec4(f16_to_f32(p[0]), f16_to_f32(p[1]), f16_to_f32(p[2]), f16_to_f32(p[3])); }
    case F_R8: return vec4((float)((const uint8_t*)b.data)[i] / 255.0f, 0.0f, 0.0f, 1.0f);
    case F_RGBA8: { const uint8_t* p = (const uint8_t*)b.data + 4 * i; return vec4((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f); }
    case F_BGRA8: { const uint8_t* p = (const uint8_t*)b.data + 4 * i; return vec4((float)p[2] / 255.0f, (float)p[1] / 255.0f, (float)p[0] / 255.0f, (float)p[3] / 255.0f); }
    }
    return vec4(0.0f);
}
inline void store_texel(const Binding& b, int x, int y, int layer, const vec4& v)
{
    if (x < 0 || y < 0 || x >= b.width || y >= b.height || layer < 0 || layer >= b.layers) return;
    size_t i = ((size_t)layer * b.height + y) * b.width + x;
    switch (b.format) {
    case F_R32F: ((float*)b.data)[i] = v.x; break;
    case F_RG32F: ((float*)b.data)[2 * i] = v.x; ((float*)b.data)[2 * i + 1] = v.y; break;
    case F_RGBA32F: { float* p = (float*)b.data + 4 * i; p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; break; }
    case F_R16F: ((uint16_t*)b.data)[i] = f32_to_f16(v.x); break;
    case F_RG16F: { uint16_t* p = (uint16_t*)b.data + 2 * i; p[0] = f32_to_f16(v.x); p[1] = f32_to_f16(v.y); break; }
    case F_RGBA16F: { uint16_t* p = (uint16_t*)b.data + 4 * i; p[0] = f32_to_f16(v.x); p[1] = f32_to_f16(v.y); p[2] = f32_to_f16(v.z); p[3] = f32_to_f16(v.w); break; }
    case F_R8: ((uint8_t*)b.data)[i] = f32_to_unorm8(v.x); break;
    case F_RGBA8: { uint8_t* p = (uint8_t*)b.data + 4 * i; p[0] = f32_to_unorm8(v.x); p[1] = f32_to_unorm8(v.y); p[2] = f32_to_unorm8(v.z); p[3] = f32_to_unorm8(v.w); break; }
    case F_BGRA8: { uint8_t* p = (uint8_t*)b.data + 4 * i; p[2] = f32_to_unorm8(v.x); p[1] = f32_to_unorm8(v.y); p[0] = f32_to_unorm8(v.z); p[3] = f32_to_unorm8(v.w); break; }
    }
}
struct image2D { int binding; };
struct image2DArray { int binding; };
struct sampler2D { int binding; };
struct sampler2DArray { int binding; };
inline vec4 imageLoad(image2D i, const ivec2& p) { return load_texel(g_bindings[i.binding], p.x, p.y, 0); }
inline vec4 imageLoad(image2DArray i, const ivec3& p) { return load_texel(g_bindings[i.binding], p.x, p.y, p.z); }
inline void imageStore(image2D i, const ivec2& p, const vec4& v) { store_texel(g_bindings[i.binding], p.x, p.y, 0, v); }
inline void imageStore(image2DArray i, const ivec3& p, const vec4& v) { store_texel(g_bindings[i.binding], p.x, p.y, p.z, v); }
inline vec4 texelFetch(sampler2D s, const ivec2& p, int) { return load_texel(g_bindings[s.binding], p.x, p.y, 0); }
inline ivec2 textureSize(sampler2D s, int) { return ivec2(g_bindings[s.binding].width, g_bindings[s.binding].height); }
inline int wrapi(int i, int n) { i %= n; return i < 0 ? i + n : i; }
inline vec4 sample_bilinear(const Binding& b, float u, float v, int layer)
{
    // default vsg::Sampler: LINEAR, REPEAT, normalised coordinates (external/vsg/include/vsg/state/Sampler.h:29-43)
    float x = u * (float)b.width - 0.5f, y = v * (float)b.height - 0.5f;
    if (!(fabsf(x) < 1e9f) || !(fabsf(y) < 1e9f)) return vec4(0.0f);   // unusable coordinates: the shaders discard the value
    float fx0 = floorf(x), fy0 = floorf(y), a = x - fx0, bt = y - fy0;
    int x0 = wrapi((int)fx0, b.width), x1 = wrapi((int)fx0 + 1, b.width), y0 = wrapi((int)fy0, b.height), y1 = wrapi((int)fy0 + 1, b.height);
    float oma = 1.0f - a, omb = 1.0f - bt, w00 = oma * omb, w10 = a * omb, w01 = oma * bt, w11 = a * bt;
    vec4 t00 = load_texel(b, x0, y0, layer), t10 = load_texel(b, x1, y0, layer), t01 = load_texel(b, x0, y1, layer), t11 = load_texel(b, x1, y1, layer);
    vec4 r;
    for (int i = 0; i < 4; ++i) r.data[i] = ((w00 * t00.data[i] + w10 * t10.data[i]) + w01 * t01.data[i]) + w11 * t11.data[i];
    return r;
}
inline vec4 texture(sampler2D s, const vec2& uv) { return sample_bilinear(g_bindings[s.binding], uv.x, uv.y, 0); }
inline vec4 texture(sampler2DArray s, const vec3& uvw) { return sample_bilinear(g_bindings[s.binding], uvw.x, uvw.y, (int)floorf(uvw.z + 0.5f)); }

// ---- invocation state, barriers, subgroup operations (runner.cpp) -------------------------------------------------
struct Invocation {
    uvec3 global_id, local_id, workgroup_id, num_workgroups;
    uint local_index, subgroup_id, subgroup_invocation, num_subgroups;
};
Invocation& inv();
void barrier();
// all live lanes of the subgroup post up to 3 words; returns the posted words of every lane and the participation mask
void subgroup_gather(const uint32_t* mine, int nwords, uint32_t (*all)[3], uint32_t* mask);

template <class F> inline float subgroup_tree(float v, float identity, F f)
{
    uint32_t all[32][3], mask, w = f2u(v);
    subgroup_gather(&w, 1, all, &mask);
    float a[32];
    for (int l = 0; l < 32; ++l) a[l] = (mask >> l) & 1u ? u2f(all[l][0]) : identity;
    for (int off = 16; off >= 1; off >>= 1)
        for (int l = 0; l < off; ++l) a[l] = f(a[l], a[l + off]);     // xor-butterfly tree, as seen from lane 0
    return a[0];
}
inline float subgroupAdd(float v) { return subgroup_tree(v, 0.0f, [](float x, float y) { return x + y; }); }
inline float subgroupMin(float v) { return subgroup_tree(v, INFINITY, [](float x, float y) { return min(x, y); }); }
inline float subgroupMax(float v) { return subgroup_tree(v, -INFINITY, [](float x, float y) { return max(x, y); }); }
inline int subgroupAdd(int v)
{
    uint32_t all[32][3], mask, w = (uint32_t)v;
    subgroup_gather(&w, 1, all, &mask);
    int s = 0;
    for (int l = 0; l < 32; ++l) if ((mask >> l) & 1u) s += (int)all[l][0];
    return s;
}
inline vec3 subgroupAdd(const vec3& v)
{
    uint32_t all[32][3], mask, w[3] = {f2u(v.x), f2u(v.y), f2u(v.z)};
    subgroup_gather(w, 3, all, &mask);
    vec3 r;
    for (int c = 0; c < 3; ++c) {
        float a[32];
        for (int l = 0; l < 32; ++l) a[l] = (mask >> l) & 1u ? u2f(all[l][c]) : 0.0f;
        for (int off = 16; off >= 1; off >>= 1)
            for (int l = 0; l < off; ++l) a[l] = a[l] + a[l + off];
        r.data[c] = a[0];
    }
    return r;
}
inline bool subgroupElect() { return inv().subgroup_invocation == 0; }   // all call sites are in subgroup-uniform control flow

}  // namespace glsl

#define gl_GlobalInvocationID (glsl::inv().global_id)
#define gl_LocalInvocationID (glsl::inv().local_id)
#define gl_WorkGroupID (glsl::inv().workgroup_id)
#define gl_NumWorkGroups (glsl::inv().num_workgroups)
#define gl_LocalInvocationIndex (glsl::inv().local_index)
#define gl_SubgroupID (glsl::inv().subgroup_id)
#define gl_SubgroupInvocationID (glsl::inv().subgroup_invocation)
#define gl_NumSubgroups (glsl::inv().num_subgroups)
#define shared static thread_local

// registration of one compiled shader instantiation (emitted by translate.py at the end of every generated unit)
namespace glsl {
struct ShaderEntry {
    const char* name; int key0, key1, key2;
    void (*main)(); int local_x, local_y;
    int* image_width; int* image_height; int* filter_radius;
    void* push_constants; int push_size;
    ShaderEntry* next;
};
void register_shader(ShaderEntry* e);
}
