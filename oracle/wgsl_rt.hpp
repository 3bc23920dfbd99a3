// wgsl_rt.hpp — C++ runtime for WGSL shaders machine-translated by oracle/wgsl2cpp.py.
// TEST INFRASTRUCTURE (part of the oracle); never linked into the product.
//
// Provides WGSL's vector/matrix types, the builtin functions the four reference shaders call
// (renamed w_<name> by the translator) and CPU texture objects. Every builtin whose result the
// WGSL / Vulkan specs leave implementation-defined is FIXED here exactly as in oracle/vk_oracle.cpp
// (header comment there), so that the hand restatement and the translated reference can be compared
// bit for bit:
//   max/min/clamp -> fmaxf/fminf (non-NaN operand wins); out-of-range textureLoad -> 0; out-of-range
//   textureStore dropped; f32->i32 truncates/saturates; mat*vec sums columns left to right;
//   dot sums left to right; normalize = v / sqrt(dot(v,v)); smoothstep, mix per WGSL spec formulas;
//   pow/sin/cos = libm float versions; rgba16float stores round to nearest even;
//   linear filtering: fp32 weights, a + f*(b-a), x then y then z, clamp-to-edge.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>

namespace wgsl {

using f32 = float;
using i32 = int32_t;
using u32 = uint32_t;

inline i32 to_i32(f32 f) {
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (i32)f;
}
inline i32 to_i32(i32 v) { return v; }
inline i32 to_i32(u32 v) { return (i32)v; }
inline u32 to_u32(f32 f) {
    if (!(f == f) || f <= 0.0f) return 0u;
    if (f >= 4294967296.0f) return UINT32_MAX;
    return (u32)f;
}
inline u32 to_u32(u32 v) { return v; }
inline u32 to_u32(i32 v) { return (u32)v; }
inline f32 to_f32(f32 v) { return v; }
inline f32 to_f32(i32 v) { return (f32)v; }
inline f32 to_f32(u32 v) { return (f32)v; }

template <class To, class From> inline To conv(From v) {
    if constexpr (std::is_same_v<To, i32>) return to_i32(v);
    else if constexpr (std::is_same_v<To, u32>) return to_u32(v);
    else if constexpr (std::is_same_v<To, f32>) return to_f32(v);
    else return (To)v;
}

template <class T> struct vec2;
template <class T> struct vec3;
template <class T> struct vec4;

template <class T> struct vec2 {
    union { T x, r; };
    union { T y, g; };
    vec2() : x(T(0)), y(T(0)) {}
    explicit vec2(T s) : x(s), y(s) {}
    vec2(T x_, T y_) : x(x_), y(y_) {}
    template <class U, class = std::enable_if_t<!std::is_same_v<U, T>>> explicit vec2(const vec2<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)) {}
#include "swz2.inc"
};
template <class T> struct vec3 {
    union { T x, r; };
    union { T y, g; };
    union { T z, b; };
    vec3() : x(T(0)), y(T(0)), z(T(0)) {}
    explicit vec3(T s) : x(s), y(s), z(s) {}
    vec3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    vec3(vec2<T> a, T z_) : x(a.x), y(a.y), z(z_) {}
    vec3(T x_, vec2<T> a) : x(x_), y(a.x), z(a.y) {}
    template <class U, class = std::enable_if_t<!std::is_same_v<U, T>>> explicit vec3(const vec3<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)), z(conv<T>(o.z)) {}
#include "swz3.inc"
};
template <class T> struct vec4 {
    union { T x, r; };
    union { T y, g; };
    union { T z, b; };
    union { T w, a; };
    vec4() : x(T(0)), y(T(0)), z(T(0)), w(T(0)) {}
    explicit vec4(T s) : x(s), y(s), z(s), w(s) {}
    vec4(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {}
    vec4(vec3<T> v, T w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    vec4(T x_, vec3<T> v) : x(x_), y(v.x), z(v.y), w(v.z) {}
    vec4(vec2<T> p, vec2<T> q) : x(p.x), y(p.y), z(q.x), w(q.y) {}
    vec4(vec2<T> p, T z_, T w_) : x(p.x), y(p.y), z(z_), w(w_) {}
    template <class U, class = std::enable_if_t<!std::is_same_v<U, T>>> explicit vec4(const vec4<U>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)), z(conv<T>(o.z)), w(conv<T>(o.w)) {}
#include "swz4.inc"
};

using vec2f = vec2<f32>; using vec3f = vec3<f32>; using vec4f = vec4<f32>;
using vec2i = vec2<i32>; using vec3i = vec3<i32>; using vec4i = vec4<i32>;
using vec2u = vec2<u32>; using vec3u = vec3<u32>; using vec4u = vec4<u32>;
using vec2b = vec2<bool>; using vec3b = vec3<bool>; using vec4b = vec4<bool>;

template <class T> using ident = typename std::type_identity<T>::type;

#define WGSL_VEC_OPS(op)                                                                                              \
    template <class T> inline vec2<T> operator op(vec2<T> a, vec2<T> b) { return {a.x op b.x, a.y op b.y}; }           \
    template <class T> inline vec2<T> operator op(vec2<T> a, ident<T> b) { return {a.x op b, a.y op b}; }              \
    template <class T> inline vec2<T> operator op(ident<T> a, vec2<T> b) { return {a op b.x, a op b.y}; }              \
    template <class T> inline vec3<T> operator op(vec3<T> a, vec3<T> b) { return {a.x op b.x, a.y op b.y, a.z op b.z}; } \
    template <class T> inline vec3<T> operator op(vec3<T> a, ident<T> b) { return {a.x op b, a.y op b, a.z op b}; }    \
    template <class T> inline vec3<T> operator op(ident<T> a, vec3<T> b) { return {a op b.x, a op b.y, a op b.z}; }    \
    template <class T> inline vec4<T> operator op(vec4<T> a, vec4<T> b) { return {a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w}; } \
    template <class T> inline vec4<T> operator op(vec4<T> a, ident<T> b) { return {a.x op b, a.y op b, a.z op b, a.w op b}; } \
    template <class T> inline vec4<T> operator op(ident<T> a, vec4<T> b) { return {a op b.x, a op b.y, a op b.z, a op b.w}; }
WGSL_VEC_OPS(+)
WGSL_VEC_OPS(-)
WGSL_VEC_OPS(*)
WGSL_VEC_OPS(/)
#undef WGSL_VEC_OPS

#define WGSL_VEC_ASSIGN(op, bop)                                                                     \
    template <class T, class U> inline vec2<T>& operator op(vec2<T>& a, U b) { a = a bop b; return a; } \
    template <class T, class U> inline vec3<T>& operator op(vec3<T>& a, U b) { a = a bop b; return a; } \
    template <class T, class U> inline vec4<T>& operator op(vec4<T>& a, U b) { a = a bop b; return a; }
WGSL_VEC_ASSIGN(+=, +)
WGSL_VEC_ASSIGN(-=, -)
WGSL_VEC_ASSIGN(*=, *)
WGSL_VEC_ASSIGN(/=, /)
#undef WGSL_VEC_ASSIGN

template <class T> inline vec2<T> operator-(vec2<T> a) { return {-a.x, -a.y}; }
template <class T> inline vec3<T> operator-(vec3<T> a) { return {-a.x, -a.y, -a.z}; }
template <class T> inline vec4<T> operator-(vec4<T> a) { return {-a.x, -a.y, -a.z, -a.w}; }

#define WGSL_VEC_CMP(op)                                                                                                   \
    template <class T> inline vec2b operator op(vec2<T> a, vec2<T> b) { return {a.x op b.x, a.y op b.y}; }                  \
    template <class T> inline vec3b operator op(vec3<T> a, vec3<T> b) { return {a.x op b.x, a.y op b.y, a.z op b.z}; }      \
    template <class T> inline vec4b operator op(vec4<T> a, vec4<T> b) { return {a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w}; }
WGSL_VEC_CMP(<)
WGSL_VEC_CMP(>)
WGSL_VEC_CMP(<=)
WGSL_VEC_CMP(>=)
#undef WGSL_VEC_CMP

inline bool w_any(vec2b v) { return v.x || v.y; }
inline bool w_any(vec3b v) { return v.x || v.y || v.z; }
inline bool w_any(vec4b v) { return v.x || v.y || v.z || v.w; }
inline bool w_all(vec2b v) { return v.x && v.y; }
inline bool w_all(vec3b v) { return v.x && v.y && v.z; }

// ---- scalar builtins ----------------------------------------------------------------------
inline f32 w_min(f32 a, f32 b) { return std::fmin(a, b); }
inline f32 w_max(f32 a, f32 b) { return std::fmax(a, b); }
inline i32 w_min(i32 a, i32 b) { return a < b ? a : b; }
inline i32 w_max(i32 a, i32 b) { return a > b ? a : b; }
inline f32 w_clamp(f32 x, f32 lo, f32 hi) { return w_min(w_max(x, lo), hi); }
inline f32 w_smoothstep(f32 e0, f32 e1, f32 x) {
    f32 t = w_clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline f32 w_mix(f32 a, f32 b, f32 t) { return a * (1.0f - t) + b * t; }
inline f32 w_pow(f32 a, f32 b) { return std::pow(a, b); }
inline f32 w_abs(f32 a) { return std::fabs(a); }
inline f32 w_floor(f32 a) { return std::floor(a); }
inline f32 w_ceil(f32 a) { return std::ceil(a); }
inline f32 w_fract(f32 a) { return a - std::floor(a); }
inline f32 w_sin(f32 a) { return std::sin(a); }
inline f32 w_cos(f32 a) { return std::cos(a); }
inline f32 w_sqrt(f32 a) { return std::sqrt(a); }

#define WGSL_MAP1(name)                                                                                   \
    inline vec2f name(vec2f a) { return {name(a.x), name(a.y)}; }                                          \
    inline vec3f name(vec3f a) { return {name(a.x), name(a.y), name(a.z)}; }                               \
    inline vec4f name(vec4f a) { return {name(a.x), name(a.y), name(a.z), name(a.w)}; }
#define WGSL_MAP2(name)                                                                                   \
    inline vec2f name(vec2f a, vec2f b) { return {name(a.x, b.x), name(a.y, b.y)}; }                       \
    inline vec3f name(vec3f a, vec3f b) { return {name(a.x, b.x), name(a.y, b.y), name(a.z, b.z)}; }       \
    inline vec4f name(vec4f a, vec4f b) { return {name(a.x, b.x), name(a.y, b.y), name(a.z, b.z), name(a.w, b.w)}; }
#define WGSL_MAP3(name)                                                                                   \
    inline vec2f name(vec2f a, vec2f b, vec2f c) { return {name(a.x, b.x, c.x), name(a.y, b.y, c.y)}; }    \
    inline vec3f name(vec3f a, vec3f b, vec3f c) { return {name(a.x, b.x, c.x), name(a.y, b.y, c.y), name(a.z, b.z, c.z)}; } \
    inline vec4f name(vec4f a, vec4f b, vec4f c) { return {name(a.x, b.x, c.x), name(a.y, b.y, c.y), name(a.z, b.z, c.z), name(a.w, b.w, c.w)}; }
WGSL_MAP1(w_abs) WGSL_MAP1(w_floor) WGSL_MAP1(w_ceil) WGSL_MAP1(w_fract) WGSL_MAP1(w_sin) WGSL_MAP1(w_cos) WGSL_MAP1(w_sqrt)
WGSL_MAP2(w_min) WGSL_MAP2(w_max) WGSL_MAP2(w_pow)
WGSL_MAP3(w_clamp) WGSL_MAP3(w_smoothstep) WGSL_MAP3(w_mix)
// mix(vecN, vecN, f32)
inline vec2f w_mix(vec2f a, vec2f b, f32 t) { return {w_mix(a.x, b.x, t), w_mix(a.y, b.y, t)}; }
inline vec3f w_mix(vec3f a, vec3f b, f32 t) { return {w_mix(a.x, b.x, t), w_mix(a.y, b.y, t), w_mix(a.z, b.z, t)}; }
inline vec4f w_mix(vec4f a, vec4f b, f32 t) { return {w_mix(a.x, b.x, t), w_mix(a.y, b.y, t), w_mix(a.z, b.z, t), w_mix(a.w, b.w, t)}; }

inline f32 w_dot(vec2f a, vec2f b) { return a.x * b.x + a.y * b.y; }
inline f32 w_dot(vec3f a, vec3f b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline f32 w_dot(vec4f a, vec4f b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline f32 w_length(f32 a) { return std::fabs(a); }
inline f32 w_length(vec2f a) { return std::sqrt(w_dot(a, a)); }
inline f32 w_length(vec3f a) { return std::sqrt(w_dot(a, a)); }
inline f32 w_length(vec4f a) { return std::sqrt(w_dot(a, a)); }
inline vec2f w_normalize(vec2f a) { return a / w_length(a); }
inline vec3f w_normalize(vec3f a) { return a / w_length(a); }
inline vec4f w_normalize(vec4f a) { return a / w_length(a); }
inline vec3f w_cross(vec3f a, vec3f b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// ---- matrices (column-major, m[c] is a column) ---------------------------------------------
struct mat3f {
    vec3f c[3];
    mat3f() {}
    mat3f(vec3f a, vec3f b, vec3f d) { c[0] = a; c[1] = b; c[2] = d; }
    vec3f& operator[](int i) { return c[i]; }
    const vec3f& operator[](int i) const { return c[i]; }
};
inline mat3f operator-(const mat3f& a, const mat3f& b) { return {a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2]}; }
inline mat3f operator+(const mat3f& a, const mat3f& b) { return {a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]}; }
inline vec3f operator*(const mat3f& m, vec3f v) { return (m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z; }
struct mat4f {
    vec4f c[4];
    mat4f() {}
    vec4f& operator[](int i) { return c[i]; }
    const vec4f& operator[](int i) const { return c[i]; }
};
inline vec4f operator*(const mat4f& m, vec4f v) { return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3] * v.w; }

// ---- fp16 ----------------------------------------------------------------------------------
inline uint32_t f2u_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f_bits(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint16_t f32_to_f16(float f) {
    uint32_t x = f2u_bits(f), sign = (x >> 16) & 0x8000u, em = x & 0x7fffffffu;
    if (em >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (em > 0x7f800000u ? 0x0200u | ((em >> 13) & 0x3ffu) : 0u));
    if (em >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);
    if (em < 0x38800000u) {
        if (em < 0x33000000u) return (uint16_t)sign;
        int e = (int)(em >> 23);
        uint32_t m = (em & 0x7fffffu) | 0x800000u;
        int shift = 126 - e;
        uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (q & 1u))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t h = (((em >> 23) - 112u) << 10) | ((em >> 13) & 0x3ffu), rem = em & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return (uint16_t)(sign | h);
}
inline float f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    if (e == 0) {
        if (m == 0) return u2f_bits(sign);
        float v = (float)m * 5.9604644775390625e-08f;
        return sign ? -v : v;
    }
    if (e == 31) return u2f_bits(sign | 0x7f800000u | (m << 13));
    return u2f_bits(sign | ((e + 112u) << 23) | (m << 13));
}

// ---- textures ------------------------------------------------------------------------------
struct Sampler { int linear = 1; };

// rgba16float storage texture, 3-D, x fastest
struct StorageTex3D {
    uint16_t* data = nullptr;
    i32 nx = 0, ny = 0, nz = 0;
};
inline vec3i textureDimensions(const StorageTex3D& t) { return {t.nx, t.ny, t.nz}; }
template <class I> inline vec4f textureLoad(const StorageTex3D& t, vec3<I> p) {
    int64_t x = (int64_t)p.x, y = (int64_t)p.y, z = (int64_t)p.z;
    if (x < 0 || y < 0 || z < 0 || x >= t.nx || y >= t.ny || z >= t.nz) return vec4f(0.0f);
    const uint16_t* q = t.data + (((size_t)z * t.ny + y) * t.nx + x) * 4;
    return {f16_to_f32(q[0]), f16_to_f32(q[1]), f16_to_f32(q[2]), f16_to_f32(q[3])};
}
template <class I> inline void textureStore(StorageTex3D& t, vec3<I> p, vec4f v) {
    int64_t x = (int64_t)p.x, y = (int64_t)p.y, z = (int64_t)p.z;
    if (x < 0 || y < 0 || z < 0 || x >= t.nx || y >= t.ny || z >= t.nz) return;
    uint16_t* q = t.data + (((size_t)z * t.ny + y) * t.nx + x) * 4;
    q[0] = f32_to_f16(v.x); q[1] = f32_to_f16(v.y); q[2] = f32_to_f16(v.z); q[3] = f32_to_f16(v.w);
}
// rgba16float storage texture, 2-D
struct StorageTex2D {
    uint16_t* data = nullptr;
    i32 w = 0, h = 0;
};
inline vec2i textureDimensions(const StorageTex2D& t) { return {t.w, t.h}; }
template <class I> inline void textureStore(StorageTex2D& t, vec2<I> p, vec4f v) {
    int64_t x = (int64_t)p.x, y = (int64_t)p.y;
    if (x < 0 || y < 0 || x >= t.w || y >= t.h) return;
    uint16_t* q = t.data + ((size_t)y * t.w + x) * 4;
    q[0] = f32_to_f16(v.x); q[1] = f32_to_f16(v.y); q[2] = f32_to_f16(v.z); q[3] = f32_to_f16(v.w);
}
// sampled 3-D texture, R8Unorm (src/context/volume_texture.rs:41-47): (r, 0, 0, 1)
struct Tex3D {
    const uint8_t* data = nullptr;
    i32 nx = 0, ny = 0, nz = 0;
    f32 at(int x, int y, int z) const {
        x = x < 0 ? 0 : (x >= nx ? nx - 1 : x); y = y < 0 ? 0 : (y >= ny ? ny - 1 : y); z = z < 0 ? 0 : (z >= nz ? nz - 1 : z);
        return (f32)data[((size_t)z * ny + y) * nx + x] / 255.0f;
    }
};
inline vec3i textureDimensions(const Tex3D& t) { return {t.nx, t.ny, t.nz}; }
inline vec4f textureSampleLevel(const Tex3D& t, const Sampler&, vec3f p, f32) {
    f32 ux = p.x * (f32)t.nx - 0.5f, uy = p.y * (f32)t.ny - 0.5f, uz = p.z * (f32)t.nz - 0.5f;
    f32 flx = std::floor(ux), fly = std::floor(uy), flz = std::floor(uz);
    f32 fx = ux - flx, fy = uy - fly, fz = uz - flz;
    int ix = to_i32(flx), iy = to_i32(fly), iz = to_i32(flz);
    f32 c000 = t.at(ix, iy, iz), c100 = t.at(ix + 1, iy, iz), c010 = t.at(ix, iy + 1, iz), c110 = t.at(ix + 1, iy + 1, iz);
    f32 c001 = t.at(ix, iy, iz + 1), c101 = t.at(ix + 1, iy, iz + 1), c011 = t.at(ix, iy + 1, iz + 1), c111 = t.at(ix + 1, iy + 1, iz + 1);
    f32 c00 = c000 + fx * (c100 - c000), c10 = c010 + fx * (c110 - c010), c01 = c001 + fx * (c101 - c001), c11 = c011 + fx * (c111 - c011);
    f32 c0 = c00 + fy * (c10 - c00), c1 = c01 + fy * (c11 - c01);
    return {c0 + fz * (c1 - c0), 0.0f, 0.0f, 1.0f};
}
// sampled 2-D texture, rgba16float backbuffer (src/context/hdr_backbuffer.rs:10)
struct Tex2D {
    const uint16_t* data = nullptr;
    i32 w = 0, h = 0;
    vec4f at(int x, int y) const {
        x = x < 0 ? 0 : (x >= w ? w - 1 : x); y = y < 0 ? 0 : (y >= h ? h - 1 : y);
        const uint16_t* q = data + ((size_t)y * w + x) * 4;
        return {f16_to_f32(q[0]), f16_to_f32(q[1]), f16_to_f32(q[2]), f16_to_f32(q[3])};
    }
};
inline vec2i textureDimensions(const Tex2D& t) { return {t.w, t.h}; }
inline vec4f textureSample(const Tex2D& t, const Sampler&, vec2f uv) {
    f32 ux = uv.x * (f32)t.w - 0.5f, uy = uv.y * (f32)t.h - 0.5f;
    f32 flx = std::floor(ux), fly = std::floor(uy);
    f32 fx = ux - flx, fy = uy - fly;
    int ix = to_i32(flx), iy = to_i32(fly);
    vec4f a = t.at(ix, iy), b = t.at(ix + 1, iy), c = t.at(ix, iy + 1), d = t.at(ix + 1, iy + 1);
    vec4f ab = a + fx * (b - a), cd = c + fx * (d - c);
    return ab + fy * (cd - ab);
}

}  // namespace wgsl
