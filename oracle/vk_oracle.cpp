// vk_oracle.cpp — CPU ORACLE for the vokselis raycast path. TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A scalar fp32 restatement of the reference's algorithm (pudnax/vokselis), function by function,
// each citing the reference file:line it follows. Built with
//     g++ -O2 -std=c++17 -fopenmp -ffp-contract=off -fno-fast-math
// so that every fp32 operation is a separately rounded IEEE operation in the order written here.
//
// PARITY PIN STATUS: pinned against oracle/_ref (the reference WGSL machine-translated by
// oracle/wgsl2cpp.py), the golden vectors under tests/golden/ generated from it, and analytic known
// answers. The reference itself holds no tests/golden vectors for this path, so
// with respect to reference-owned vectors parity is UNPINNED; see DESIGN.md §3.
//
// Definitions this oracle FIXES where the reference leaves them to the Vulkan driver:
//   * max/min/clamp return the non-NaN operand (fmaxf/fminf)            [SURVEY F13/H4]
//   * an out-of-range textureLoad returns (0,0,0,0)                      [SURVEY H3]
//   * f32 -> i32 conversion truncates toward zero, saturates, NaN -> 0
//   * mat4*vec4 = ((c0*x + c1*y) + c2*z) + c3*w; dot3 = (x*x' + y*y') + z*z'
//   * normalize(v) = v / sqrt(dot(v,v)); smoothstep/mix per the WGSL spec formulas
//   * pow = powf, cos = cosf, sin = sinf (libm)
//   * textureStore to rgba16float rounds to nearest even
#include "vk_oracle.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// fp16 <-> fp32 (IEEE binary16, round-to-nearest-even), bit-level so the result does not depend
// on the host's F16C support.
inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

uint16_t f32_to_f16(float f) {
    uint32_t x = f2u(f);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t em = x & 0x7fffffffu;
    if (em >= 0x7f800000u) {  // inf / NaN
        return (uint16_t)(sign | 0x7c00u | (em > 0x7f800000u ? 0x0200u | ((em >> 13) & 0x3ffu) : 0u));
    }
    if (em >= 0x477ff000u) {  // >= 65520 rounds to inf
        return (uint16_t)(sign | 0x7c00u);
    }
    if (em < 0x38800000u) {  // subnormal half or zero (|f| < 2^-14)
        if (em < 0x33000000u) return (uint16_t)sign;  // < 2^-25 -> 0 (2^-25 exactly ties to even 0)
        int e = (int)(em >> 23);                      // biased fp32 exponent, 102..112
        uint32_t m = (em & 0x7fffffu) | 0x800000u;    // 24-bit significand
        int shift = 126 - e;                          // 14..24: value = m * 2^(e-150); half sub ulp = 2^-24
        uint32_t q = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1);
        if (rem > half || (rem == half && (q & 1u))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t e = ((em >> 23) - 112u) << 10;
    uint32_t m = (em >> 13) & 0x3ffu;
    uint32_t h = e | m;
    uint32_t rem = em & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;  // carry may roll into exponent: correct
    return (uint16_t)(sign | h);
}

float f16_to_f32_slow(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    if (e == 0) {
        if (m == 0) return u2f(sign);
        float v = (float)m * 5.9604644775390625e-08f;  // m * 2^-24, exact
        return sign ? -v : v;
    }
    if (e == 31) return u2f(sign | 0x7f800000u | (m << 13));
    return u2f(sign | ((e + 112u) << 23) | (m << 13));
}

struct HalfLut {
    float t[65536];
    HalfLut() { for (uint32_t i = 0; i < 65536; ++i) t[i] = f16_to_f32_slow((uint16_t)i); }
};
const HalfLut& lut() { static HalfLut l; return l; }
inline float h2f(uint16_t h) { return lut().t[h]; }

// ---------------------------------------------------------------------------------------------
// Minimal vector algebra with the evaluation orders fixed in the header comment.
struct v3 { float x, y, z; };
struct v4 { float x, y, z, w; };
inline v3 operator+(v3 a, v3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline v3 operator-(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline v3 operator*(v3 a, v3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline v3 operator*(float s, v3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline v3 operator*(v3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline v3 operator/(v3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline v3 normalize3(v3 a) { float l = std::sqrt(dot3(a, a)); return a / l; }
inline float wmax(float a, float b) { return std::fmax(a, b); }
inline float wmin(float a, float b) { return std::fmin(a, b); }
inline float wclamp(float x, float lo, float hi) { return wmin(wmax(x, lo), hi); }
inline float wsmoothstep(float e0, float e1, float x) {
    float t = wclamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline float wmix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline int32_t f2i(float f) {  // WGSL i32(f32): truncate; saturating; NaN -> 0 (CUDA cvt.rzi.s32.f32 semantics)
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
// column-major mat4 * vec4
inline v4 mat4_mul(const float* m, v4 v) {
    v4 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
    r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
    return r;
}

// ---------------------------------------------------------------------------------------------
// shaders/raycast_compute.wgsl:102-116 — ray generation in `render`.
//   coord = vec2<f32>(global_id) + offset            (no half-pixel offset)
//   sc = 2*coord/dims - 1;  sc.y *= -(dims.y/dims.x)
//   eye = (inv * (sc,0,1)).xyz/w ;  dir = normalize((inv * (sc,1,1)).xyz/w - eye)
inline void gen_ray(const float* inv, float gx, float gy, float offx, float offy, float W, float H,
                    v3& eye, v3& dir) {
    float cx = gx + offx, cy = gy + offy;
    float aspect_ratio = H / W;
    float sx = 2.0f * cx / W - 1.0f;
    float sy = 2.0f * cy / H - 1.0f;
    sy = sy * (-aspect_ratio);
    v4 sp = {sx, sy, 0.0f, 1.0f};
    v4 st = {sx + 0.0f, sy + 0.0f, 0.0f + 1.0f, 1.0f + 0.0f};
    v4 vp = mat4_mul(inv, sp);
    v4 vt = mat4_mul(inv, st);
    eye = v3{vp.x, vp.y, vp.z} / vp.w;
    dir = normalize3(v3{vt.x, vt.y, vt.z} / vt.w - eye);
}

// shaders/raycast_compute.wgsl:42-53 — slab test against [-1,1]^3.
inline void intersect_box(v3 o, v3 d, float& t0, float& t1) {
    v3 inv = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    v3 a = (v3{-1.0f, -1.0f, -1.0f} - o) * inv;
    v3 b = (v3{1.0f, 1.0f, 1.0f} - o) * inv;
    v3 mn = {wmin(a.x, b.x), wmin(a.y, b.y), wmin(a.z, b.z)};
    v3 mx = {wmax(a.x, b.x), wmax(a.y, b.y), wmax(a.z, b.z)};
    t0 = wmax(mn.x, wmax(mn.y, mn.z));
    t1 = wmin(mx.x, wmin(mx.y, mx.z));
}

// shaders/raycast_compute.wgsl:65-68 — step length. N = textureDimensions(volume).
inline float step_dt(const VkrtParams& P, v3 dir, float nx, float ny, float nz) {
    float dx = 1.0f / (nx * std::fabs(dir.x));
    float dy = 1.0f / (ny * std::fabs(dir.y));
    float dz = 1.0f / (nz * std::fabs(dir.z));
    return P.dt_scale * wmax(wmin(dx, wmin(dy, dz)), P.dt_floor);
}

struct Sampler {
    const VkoVolume* v;
    float fx, fy, fz;        // dims as float
    float hx, hy, hz;        // dims / 2
    bool has_brick = false;
    VkoBrick brick{};
};

inline bool in_brick(const Sampler& S, int ix, int iy, int iz) {
    if (!S.has_brick) return true;
    return (float)ix >= S.brick.lo[0] && (float)ix < S.brick.hi[0] && (float)iy >= S.brick.lo[1] &&
           (float)iy < S.brick.hi[1] && (float)iz >= S.brick.lo[2] && (float)iz < S.brick.hi[2];
}

// shaders/raycast_compute.wgsl:71-73 — `samp = vec3<i32>((p + 1.) * (block_size / 2.))`,
// textureLoad (nearest, integer coords). Out-of-range -> zero texel.
inline void m0_fetch(const Sampler& S, int ix, int iy, int iz, v4& c, v4& n) {
    const VkoVolume* v = S.v;
    if ((unsigned)ix >= (unsigned)v->nx || (unsigned)iy >= (unsigned)v->ny || (unsigned)iz >= (unsigned)v->nz) {
        c = {0, 0, 0, 0};
        n = {0, 0, 0, 0};
        return;
    }
    size_t i = (((size_t)iz * v->ny + iy) * v->nx + ix) * 4;
    c = {h2f(v->color[i]), h2f(v->color[i + 1]), h2f(v->color[i + 2]), h2f(v->color[i + 3])};
    n = {h2f(v->normal[i]), h2f(v->normal[i + 1]), h2f(v->normal[i + 2]), h2f(v->normal[i + 3])};
}

inline float scalar_at(const VkoVolume* v, int ix, int iy, int iz) {
    // clamp-to-edge: src/context/volume_texture.rs:61-66 (default address mode)
    ix = ix < 0 ? 0 : (ix >= v->nx ? v->nx - 1 : ix);
    iy = iy < 0 ? 0 : (iy >= v->ny ? v->ny - 1 : iy);
    iz = iz < 0 ? 0 : (iz >= v->nz ? v->nz - 1 : iz);
    size_t i = ((size_t)iz * v->ny + iy) * v->nx + ix;
    switch (v->dtype) {
        case VKRT_U8: return (float)((const uint8_t*)v->scalar)[i] / 255.0f;  // R8Unorm
        case VKRT_F16: return h2f(((const uint16_t*)v->scalar)[i]);
        default: return ((const float*)v->scalar)[i];
    }
}

// Linear-filter sample (src/context/volume_texture.rs:61-66 sampler; use site
// shaders/raycast_naive.wgsl:102). u = unnormalised texel-space coordinate minus 0.5, i.e. texel
// centres sit at integer u. Weights in fp32; lerp(a,b,f) = a + f*(b-a); x, then y, then z.
inline float trilinear(const VkoVolume* v, float ux, float uy, float uz) {
    float flx = std::floor(ux), fly = std::floor(uy), flz = std::floor(uz);
    float fx = ux - flx, fy = uy - fly, fz = uz - flz;
    int ix = f2i(flx), iy = f2i(fly), iz = f2i(flz);
    float c000 = scalar_at(v, ix, iy, iz), c100 = scalar_at(v, ix + 1, iy, iz);
    float c010 = scalar_at(v, ix, iy + 1, iz), c110 = scalar_at(v, ix + 1, iy + 1, iz);
    float c001 = scalar_at(v, ix, iy, iz + 1), c101 = scalar_at(v, ix + 1, iy, iz + 1);
    float c011 = scalar_at(v, ix, iy + 1, iz + 1), c111 = scalar_at(v, ix + 1, iy + 1, iz + 1);
    float c00 = c000 + fx * (c100 - c000), c10 = c010 + fx * (c110 - c010);
    float c01 = c001 + fx * (c101 - c001), c11 = c011 + fx * (c111 - c011);
    float c0 = c00 + fy * (c10 - c00), c1 = c01 + fy * (c11 - c01);
    return c0 + fz * (c1 - c0);
}

// shaders/raycast_naive.wgsl:63-68
inline float linear_to_srgb_naive(float x) {
    if (x <= 0.0031308f) return 12.92f * x;
    return 1.055f * std::pow(x, 1.0f / 2.4f) - 0.055f;
}

struct MarchOut {
    float r, g, b, a;
    uint32_t iters;    // loop iterations executed (reference semantics)
    uint32_t fetched;  // iterations that fetched texels (== iters here unless a brick restricts)
};

// shaders/raycast_compute.wgsl:62-97 — `get_col2`, one iteration per sample, literal.
// (r,g,b,a) is the running colour on entry.
inline void march_m0(const Sampler& S, const VkrtParams& P, v3 eye, v3 dir, float tmin, float tmax, MarchOut& o) {
    const float cr = P.clear_color[0], cg = P.clear_color[1], cb = P.clear_color[2], ca = P.clear_color[3];
    const v3 light = {0.0f, -1.0f, 0.0f};
    const v3 ldir = normalize3(v3{-2.0f, -2.0f, -1.0f});
    const v3 pdir = normalize3(v3{1.0f, 1.0f, -1.0f});
    const v3 dcol = 3.0f * v3{1.0f, 0.1f, 0.13f};
    const float dt = step_dt(P, dir, S.fx, S.fy, S.fz);
    float r = o.r, g = o.g, b = o.b, a = o.a;
    uint32_t iters = 0, fetched = 0;
    for (float t = tmin; t < tmax; t = t + dt) {
        ++iters;
        v3 p = eye + t * dir;
        int ix = f2i((p.x + 1.0f) * S.hx), iy = f2i((p.y + 1.0f) * S.hy), iz = f2i((p.z + 1.0f) * S.hz);
        if (!in_brick(S, ix, iy, iz)) continue;
        ++fetched;
        v4 c, n;
        m0_fetch(S, ix, iy, iz, c, n);
        v3 nn = {n.x, n.y, n.z};
        float shade_s = wmax(0.0f, dot3(light, nn));
        float vol_alpha = std::pow(c.w, 3.0f);
        vol_alpha = wsmoothstep(0.0f, 0.7f, vol_alpha);
        v3 directional = dcol * wmax(dot3(nn, ldir), 0.0f);
        directional = directional * wsmoothstep(0.3f, 1.5f, dot3(p, pdir));
        v3 vol_color = v3{c.x, c.y, c.z} + directional;
        float bottom_light = 0.9f * wclamp(0.5f - 0.5f * n.y, 0.0f, 1.0f);
        v3 bl = bottom_light * v3{0.0f, 0.0f, 0.6f};
        v3 shade = {wmix(shade_s, bl.x, 0.2f), wmix(shade_s, bl.y, 0.2f), wmix(shade_s, bl.z, 0.2f)};
        float w = (1.0f - a) * vol_alpha;
        float tr = r + w * vol_color.x * shade.x;
        float tg = g + w * vol_color.y * shade.y;
        float tb = b + w * vol_color.z * shade.z;
        float k = 1.0f - vol_alpha;
        r = tr + cr * ca * k;
        g = tg + cg * ca * k;
        b = tb + cb * ca * k;
        a = a + (1.0f - a) * vol_alpha * (1.0f - ca);
        if (a >= P.alpha_threshold) break;
    }
    o.r = r; o.g = g; o.b = b; o.a = a;
    o.iters = iters; o.fetched = fetched;
}

// M1 — march body of shaders/raycast_naive.wgsl:96-119 on M0's rays and [-1,1]^3 box:
//   s = textureSampleLevel(volume, linear clamp sampler, p01)          (:102)
//   val = clamp(vec3(0.4), vec3(0.9), s.rgb) = min(0.9, s.r)            (:106; e=0.4, low=0.9, high=val)
//   val = smoothstep(0.10, 1.2, val)                                    (:107)
//   val_color = (vertigo(val.r), val.r)                                 (:108, :70-81)
//   rgb += (1-a)*val_color.a*val_color.rgb + background.rgb*background.a*(1-pow(s.a,2))  (:112; s.a==1 for R8 -> +0)
//   a += (1-a)*val_color.a ; break at threshold                         (:114-117)
// p = eye + t*dir as in M0 (the naive shader accumulates p += dir*dt; M1 is defined on M0's form).
inline void march_m1(const Sampler& S, const VkrtParams& P, v3 eye, v3 dir, float tmin, float tmax, MarchOut& o) {
    const float TAU = 6.28318f;
    const float dt = step_dt(P, dir, S.fx, S.fy, S.fz);
    float r = o.r, g = o.g, b = o.b, a = o.a;
    uint32_t iters = 0, fetched = 0;
    for (float t = tmin; t < tmax; t = t + dt) {
        ++iters;
        v3 p = eye + t * dir;
        float qx = (p.x + 1.0f) * S.hx, qy = (p.y + 1.0f) * S.hy, qz = (p.z + 1.0f) * S.hz;
        if (S.has_brick && !in_brick(S, f2i(qx), f2i(qy), f2i(qz))) continue;
        ++fetched;
        float s = trilinear(S.v, qx - 0.5f, qy - 0.5f, qz - 0.5f);
        float val = wmin(wmax(0.4f, 0.9f), s);
        val = wsmoothstep(0.10f, 1.2f, val);
        float pr = 0.5f + 0.5f * std::cos(TAU * (1.0f * val + 0.0f));
        float pg = 0.5f + 0.5f * std::cos(TAU * (1.7f * val + 0.15f));
        float pb = 0.5f + 0.5f * std::cos(TAU * (0.4f * val + 0.20f));
        float w = (1.0f - a) * val;
        r = r + w * pr;
        g = g + w * pg;
        b = b + w * pb;
        a = a + (1.0f - a) * val;
        if (a >= P.alpha_threshold) break;
    }
    o.r = r; o.g = g; o.b = b; o.a = a;
    o.iters = iters; o.fetched = fetched;
}

inline void store_px(uint16_t* frame, int W, int x, int y, float r, float g, float b, float a) {
    uint16_t* px = frame + ((size_t)y * W + x) * 4;
    px[0] = f32_to_f16(r); px[1] = f32_to_f16(g); px[2] = f32_to_f16(b); px[3] = f32_to_f16(a);
}

Sampler make_sampler(const VkoVolume* vol) {
    Sampler S;
    S.v = vol;
    S.fx = (float)vol->nx; S.fy = (float)vol->ny; S.fz = (float)vol->nz;
    S.hx = S.fx / 2.0f; S.hy = S.fy / 2.0f; S.hz = S.fz / 2.0f;
    return S;
}

// shaders/raycast_compute.wgsl:99-131 — `render` for one invocation.
// gx,gy = global_invocation_id.xy; (offx,offy) = dyn_offset for `tile`, 0 for `single`.
inline void render_pixel(const Sampler& S, const VkrtParams& P, const float* inv, uint32_t gx, uint32_t gy,
                         float offx, float offy, int W, int H, float out[4], uint32_t& aux, MarchOut& mo) {
    v3 eye, dir;
    gen_ray(inv, (float)gx, (float)gy, offx, offy, (float)W, (float)H, eye, dir);
    aux = 0;
    mo = MarchOut{};
    // `if (any(vec2<f32>(global_id.xy) < dims))` (:121): false only when BOTH ids are >= dims.
    if (!((float)gx < (float)W || (float)gy < (float)H)) {
        out[0] = out[1] = out[2] = out[3] = 0.0f;
        return;
    }
    float t0, t1;
    intersect_box(eye, dir, t0, t1);
    if (t0 < t1) {
        t0 = wmax(t0, 0.0f);
        mo.r = P.clear_color[0]; mo.g = P.clear_color[1]; mo.b = P.clear_color[2]; mo.a = P.initial_alpha;
        if (P.mode == VKRT_MODE_M0) {
            march_m0(S, P, eye, dir, t0, t1, mo);
        } else {
            mo.r = mo.g = mo.b = 0.0f;  // raycast_naive.wgsl:96 `var color = vec4(0.0)`
            march_m1(S, P, eye, dir, t0, t1, mo);
            if (P.m1_srgb) {
                mo.r = linear_to_srgb_naive(mo.r); mo.g = linear_to_srgb_naive(mo.g); mo.b = linear_to_srgb_naive(mo.b);
            }
        }
        out[0] = mo.r; out[1] = mo.g; out[2] = mo.b; out[3] = 1.0f;
        aux = 0x80000000u | (mo.iters & 0x7fffffffu);
    } else {
        out[0] = P.clear_color[0]; out[1] = P.clear_color[1]; out[2] = P.clear_color[2]; out[3] = 1.0f;
    }
}

int resolve_threads(int nthreads) {
#ifdef _OPENMP
    return nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    (void)nthreads;
    return 1;
#endif
}

bool volume_ok(const VkoVolume* v, const VkrtParams* P) {
    if (!v || v->nx <= 0 || v->ny <= 0 || v->nz <= 0) return false;
    if (P->mode == VKRT_MODE_M0) return v->dtype == -1 && v->color && v->normal;
    return v->dtype >= VKRT_U8 && v->dtype <= VKRT_F32 && v->scalar;
}

}  // namespace

extern "C" {

int vko_num_threads(void) { return resolve_threads(0); }
uint16_t vko_f32_to_f16(float f) { return f32_to_f16(f); }
float vko_f16_to_f32(uint16_t h) { return f16_to_f32_slow(h); }

int vko_render(const VkoVolume* vol, const VkrtParams* params, const VkrtCameraUniform* cam, int W, int H,
               const VkrtOffset* offsets, int n_offsets, uint16_t* frame, uint32_t* aux, VkrtStats* stats,
               int nthreads) {
    if (!params || !cam || !frame || W <= 0 || H <= 0 || !volume_ok(vol, params)) return VKRT_ERR_INVALID;
    const VkrtParams P = *params;
    const Sampler S = make_sampler(vol);
    const float* inv = cam->inv_proj;
    uint64_t hits = 0, its = 0, fet = 0;
    const int nt = resolve_threads(nthreads);
    (void)nt;
    if (!offsets || n_offsets <= 0) {
        // `single` (:133-137): one invocation per pixel; ceil-div groups overshoot, stores out of range are dropped.
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt) reduction(+ : hits, its, fet)
        for (int y = 0; y < H; ++y) {
            for (int x = 0; x < W; ++x) {
                float o[4]; uint32_t a; MarchOut mo;
                render_pixel(S, P, inv, (uint32_t)x, (uint32_t)y, 0.0f, 0.0f, W, H, o, a, mo);
                store_px(frame, W, x, y, o[0], o[1], o[2], o[3]);
                if (aux) aux[(size_t)y * W + x] = a;
                if (a & 0x80000000u) { ++hits; its += mo.iters; fet += mo.fetched; }
            }
        }
    } else {
        // `tile` (:139-144): global_id in [0,tile_size)^2; store at global_id + vec2<u32>(offset).
        const int ts = P.tile_size;
        for (int k = 0; k < n_offsets; ++k) {
            const float offx = offsets[k].x, offy = offsets[k].y;
            // vec2<u32>(vec2<f32>): truncates and SATURATES (negative / NaN -> 0), like oracle/wgsl_rt.hpp to_u32
            auto sat_u32 = [](float f) -> uint32_t { return (!(f == f) || f <= 0.0f) ? 0u : (f >= 4294967296.0f ? 0xffffffffu : (uint32_t)f); };
            const uint32_t ox = sat_u32(offx), oy = sat_u32(offy);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt) reduction(+ : hits, its, fet)
            for (int gy = 0; gy < ts; ++gy) {
                for (int gx = 0; gx < ts; ++gx) {
                    uint32_t px = (uint32_t)gx + ox, py = (uint32_t)gy + oy;
                    if (px >= (uint32_t)W || py >= (uint32_t)H) continue;  // textureStore out of range: dropped
                    float o[4]; uint32_t a; MarchOut mo;
                    render_pixel(S, P, inv, (uint32_t)gx, (uint32_t)gy, offx, offy, W, H, o, a, mo);
                    store_px(frame, W, (int)px, (int)py, o[0], o[1], o[2], o[3]);
                    if (aux) aux[(size_t)py * W + px] = a;
                    if (a & 0x80000000u) { ++hits; its += mo.iters; fet += mo.fetched; }
                }
            }
        }
    }
    if (stats) {
        stats->rays_hit = hits; stats->samples_reference = its; stats->samples_fetched = fet;
        stats->last_render_ms = 0.0f; stats->_pad = 0.0f;
    }
    return VKRT_OK;
}

int vko_render_partial(const VkoVolume* vol, const VkrtParams* params, const VkrtCameraUniform* cam, int W, int H,
                       const VkoBrick* brick, const float* a_in, float* partial, int nthreads) {
    if (!params || !cam || !partial || !brick || W <= 0 || H <= 0 || !volume_ok(vol, params)) return VKRT_ERR_INVALID;
    const VkrtParams P = *params;
    Sampler S = make_sampler(vol);
    S.has_brick = true;
    S.brick = *brick;
    const float* inv = cam->inv_proj;
    const int nt = resolve_threads(nthreads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float* o = partial + ((size_t)y * W + x) * 4;
            float a0 = a_in ? a_in[(size_t)y * W + x] : 0.0f;
            o[0] = o[1] = o[2] = 0.0f; o[3] = a0;
            v3 eye, dir;
            gen_ray(inv, (float)x, (float)y, 0.0f, 0.0f, (float)W, (float)H, eye, dir);
            float t0, t1;
            intersect_box(eye, dir, t0, t1);
            if (!(t0 < t1) || a0 >= P.alpha_threshold) continue;
            t0 = wmax(t0, 0.0f);
            MarchOut mo{};
            mo.a = a0;
            if (P.mode == VKRT_MODE_M0) march_m0(S, P, eye, dir, t0, t1, mo);
            else march_m1(S, P, eye, dir, t0, t1, mo);
            o[0] = mo.r; o[1] = mo.g; o[2] = mo.b; o[3] = mo.a;
        }
    }
    return VKRT_OK;
}

int vko_rays(const VkrtCameraUniform* cam, int W, int H, float offx, float offy, float* out8) {
    if (!cam || !out8 || W <= 0 || H <= 0) return VKRT_ERR_INVALID;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            v3 eye, dir;
            gen_ray(cam->inv_proj, (float)x, (float)y, offx, offy, (float)W, (float)H, eye, dir);
            float t0, t1;
            intersect_box(eye, dir, t0, t1);
            float* o = out8 + ((size_t)y * W + x) * 8;
            o[0] = eye.x; o[1] = eye.y; o[2] = eye.z; o[3] = dir.x; o[4] = dir.y; o[5] = dir.z; o[6] = t0; o[7] = t1;
        }
    }
    return VKRT_OK;
}

// shaders/raycast_naive.wgsl:83-125 — `fs_main`, literal (box [0,1]^3, p accumulated by += dir*dt,
// dt from the literal 256, in-shader sRGB). Input per fragment: what vs_main hands over (:40-48),
// transformed_eye and ray_dir. This is M1's semantic source; M1 itself (march_m1) runs on the
// compute boundary's rays. out4 = fragment colour.
int vko_naive_fs(const uint8_t* vol, int nx, int ny, int nz, int count, const float* eye3, const float* dir3, float* out4,
                 int nthreads) {
    if (!vol || !eye3 || !dir3 || !out4) return VKRT_ERR_INVALID;
    VkoVolume V{};
    V.nx = nx; V.ny = ny; V.nz = nz; V.dtype = VKRT_U8; V.scalar = vol;
    const int nt = resolve_threads(nthreads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
    for (int i = 0; i < count; ++i) {
        v3 eye = {eye3[i * 3], eye3[i * 3 + 1], eye3[i * 3 + 2]};
        v3 ray_dir = normalize3(v3{dir3[i * 3], dir3[i * 3 + 1], dir3[i * 3 + 2]});
        float* o = out4 + i * 4;
        // intersect_box (:50-61) against [0,1]^3
        v3 inv = {1.0f / ray_dir.x, 1.0f / ray_dir.y, 1.0f / ray_dir.z};
        v3 a = (v3{0.0f, 0.0f, 0.0f} - eye) * inv, b = (v3{1.0f, 1.0f, 1.0f} - eye) * inv;
        float t0 = wmax(wmin(a.x, b.x), wmax(wmin(a.y, b.y), wmin(a.z, b.z)));
        float t1 = wmin(wmax(a.x, b.x), wmin(wmax(a.y, b.y), wmax(a.z, b.z)));
        if (t0 > t1) { o[0] = o[1] = o[2] = 0.0f; o[3] = 1.0f; continue; }
        t0 = wmax(t0, 0.0f);
        float r = 0.0f, g = 0.0f, bl = 0.0f, al = 0.0f;
        const float dx = 1.0f / (256.0f * std::fabs(ray_dir.x)), dy = 1.0f / (256.0f * std::fabs(ray_dir.y)),
                    dz = 1.0f / (256.0f * std::fabs(ray_dir.z));
        const float dt = 1.0f * wmin(dx, wmin(dy, dz));
        v3 p = eye + t0 * ray_dir;
        for (float t = t0; t < t1; t = t + dt) {
            float s = trilinear(&V, p.x * (float)nx - 0.5f, p.y * (float)ny - 0.5f, p.z * (float)nz - 0.5f);
            // tex_content = (s, 0, 0, 1): val = .rgb, val_alpha = pow(.a, 2)
            float val_alpha = std::pow(1.0f, 2.0f);
            float val = wmin(wmax(0.4f, 0.9f), s);
            val = wsmoothstep(0.10f, 1.2f, val);
            const float TAU = 6.28318f;
            float pr = 0.5f + 0.5f * std::cos(TAU * (1.0f * val + 0.0f));
            float pg = 0.5f + 0.5f * std::cos(TAU * (1.7f * val + 0.15f));
            float pb = 0.5f + 0.5f * std::cos(TAU * (0.4f * val + 0.20f));
            float w = (1.0f - al) * val;
            float k = 1.0f - val_alpha;
            r = r + w * pr + 0.1f * 0.01f * k;
            g = g + w * pg + 0.2f * 0.01f * k;
            bl = bl + w * pb + 0.3f * 0.01f * k;
            al = al + (1.0f - al) * val;
            if (al >= 0.95f) break;
            p = p + ray_dir * dt;
        }
        o[0] = linear_to_srgb_naive(r); o[1] = linear_to_srgb_naive(g); o[2] = linear_to_srgb_naive(bl); o[3] = 1.0f;
    }
    return VKRT_OK;
}

// ---------------------------------------------------------------------------------------------
// shaders/xor.wgsl — the volume generator.
namespace {
inline float wfract(float x) { return x - std::floor(x); }
// :18-20
inline float hash1(float h) { return wfract(std::sin(h) * 43758.5453123f); }
// :22-33
inline float noise3(v3 x) {
    v3 p = {std::floor(x.x), std::floor(x.y), std::floor(x.z)};
    v3 f = {wfract(x.x), wfract(x.y), wfract(x.z)};
    f = {f.x * f.x * (3.0f - 2.0f * f.x), f.y * f.y * (3.0f - 2.0f * f.y), f.z * f.z * (3.0f - 2.0f * f.z)};
    float n = p.x + p.y * 157.0f + 113.0f * p.z;
    return wmix(
        wmix(wmix(hash1(n + 0.0f), hash1(n + 1.0f), f.x), wmix(hash1(n + 157.0f), hash1(n + 158.0f), f.x), f.y),
        wmix(wmix(hash1(n + 113.0f), hash1(n + 114.0f), f.x), wmix(hash1(n + 270.0f), hash1(n + 271.0f), f.x), f.y),
        f.z);
}
// :35-44
inline float fbm(v3 p) {
    float f = 0.5000f * noise3(p);
    p = p * 2.01f;
    f += 0.2500f * noise3(p);
    p = p * 2.02f;
    f += 0.1250f * noise3(p);
    return f;
}
inline float length3(v3 a) { return std::sqrt(dot3(a, a)); }
// :55-61 noise_volume ; :46-53 volume (dead code in the reference, kept as generator `which=1`)
inline v4 gen_volume(v3 coord, float time, int which) {
    v3 pos = (coord + v3{1.0f, std::sin(time * 1.0f) * 0.1f, 21.0f}) * 32.0f;
    if (which == 0) {
        float val = fbm(pos);
        float alpha = val * wsmoothstep(0.5f, 0.25f, length3(coord));
        return {val, val, val, alpha};
    }
    const float res = 25.0f;
    float val = (float)(f2i(pos.x * res) & f2i(pos.y * res) & f2i(pos.z * res)) / res;
    float alpha = val * wsmoothstep(0.7f, 0.0f, length3(coord));
    return {val, val, val, alpha};
}
// :63-67 — the reference's gradient() always differentiates noise_volume (even if cs_main were
// switched to volume()); `which` is threaded through so generator 1 is self-consistent.
inline v3 gradient(v3 pos, float eps, float time, int which) {
    v3 k0 = pos - v3{eps, 0.0f, 0.0f}, k1 = pos - v3{0.0f, eps, 0.0f}, k2 = pos - v3{0.0f, 0.0f, eps};
    float a = gen_volume(pos, time, which).w;
    v3 d = v3{a, a, a} - v3{gen_volume(k0, time, which).w, gen_volume(k1, time, which).w, gen_volume(k2, time, which).w};
    return normalize3(d);
}
}  // namespace

// :69-78 cs_main
int vko_generate_xor(int n, float time, int which, uint16_t* color, uint16_t* normal, int nthreads) {
    if (n <= 0 || !color || !normal || which < 0 || which > 1) return VKRT_ERR_INVALID;
    const int nt = resolve_threads(nthreads);
    (void)nt;
    const float dims = (float)n;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int z = 0; z < n; ++z) {
        for (int y = 0; y < n; ++y) {
            for (int x = 0; x < n; ++x) {
                v3 coord = (v3{(float)x, (float)y, (float)z} - v3{dims / 2.0f, dims / 2.0f, dims / 2.0f}) / dims;
                v4 vol = gen_volume(coord, time, which);
                v3 nrm = gradient(coord, 0.0001f, time, which);
                size_t i = (((size_t)z * n + y) * n + x) * 4;
                color[i] = f32_to_f16(vol.x / 2.0f); color[i + 1] = f32_to_f16(vol.y / 2.0f);
                color[i + 2] = f32_to_f16(vol.z / 2.0f); color[i + 3] = f32_to_f16(vol.w);
                normal[i] = f32_to_f16(nrm.x); normal[i + 1] = f32_to_f16(nrm.y); normal[i + 2] = f32_to_f16(nrm.z);
                normal[i + 3] = f32_to_f16(length3(nrm));
            }
        }
    }
    return VKRT_OK;
}

// "Next" row N3 — scalar grid -> rgba16f pair, the construction of shaders/xor.wgsl cs_main (:69-78) and
// gradient (:63-67) on a sampled field: colour = (a/2, a/2, a/2, a); normal = normalize(a(p) - (a(p-ex),
// a(p-ey), a(p-ez))) with one-voxel backward differences clamped at the border; normal.w = length(normal).
int vko_scalar_to_rgba16f(const void* scalar, int dtype, int nx, int ny, int nz, uint16_t* color, uint16_t* normal) {
    if (!scalar || !color || !normal || nx <= 0 || ny <= 0 || nz <= 0 || dtype < VKRT_U8 || dtype > VKRT_F32) return VKRT_ERR_INVALID;
    VkoVolume V{};
    V.nx = nx; V.ny = ny; V.nz = nz; V.dtype = dtype; V.scalar = scalar;
#pragma omp parallel for schedule(static)
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const float a = scalar_at(&V, x, y, z);
                v3 d = v3{a, a, a} - v3{scalar_at(&V, x - 1, y, z), scalar_at(&V, x, y - 1, z), scalar_at(&V, x, y, z - 1)};
                v3 n = normalize3(d);
                const size_t i = (((size_t)z * ny + y) * nx + x) * 4;
                color[i] = f32_to_f16(a / 2.0f); color[i + 1] = f32_to_f16(a / 2.0f); color[i + 2] = f32_to_f16(a / 2.0f); color[i + 3] = f32_to_f16(a);
                normal[i] = f32_to_f16(n.x); normal[i + 1] = f32_to_f16(n.y); normal[i + 2] = f32_to_f16(n.z); normal[i + 3] = f32_to_f16(length3(n));
            }
    return VKRT_OK;
}

// ---------------------------------------------------------------------------------------------
// shaders/present.wgsl:23-35,111-119 — ACES then sRGB, written to an Rgba8Unorm target.
// 1:1 (backbuffer size == target size): the bilinear `textureSample` at pixel centres returns the
// texel itself. Unorm conversion: round(clamp(x,0,1)*255) (Vulkan float->unorm, RNE on the product).
int vko_present(const uint16_t* frame, int W, int H, uint8_t* rgba8) {
    if (!frame || !rgba8 || W <= 0 || H <= 0) return VKRT_ERR_INVALID;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < W * H; ++i) {
        float c[4] = {h2f(frame[i * 4]), h2f(frame[i * 4 + 1]), h2f(frame[i * 4 + 2]), h2f(frame[i * 4 + 3])};
        for (int k = 0; k < 3; ++k) {
            float x = c[k];
            // ACESFilm (:33-35)
            float v = wclamp((x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f), 0.0f, 1.0f);
            // linear_to_srgb (:23-30): selector = ceil(v - 0.0031308); mix(under, over, selector)
            float sel = std::ceil(v - 0.0031308f);
            float under = 12.92f * v;
            float over = 1.055f * std::pow(v, 0.41666f) - 0.055f;
            c[k] = wmix(under, over, sel);
        }
        for (int k = 0; k < 4; ++k) {
            float q = wclamp(c[k], 0.0f, 1.0f) * 255.0f;
            rgba8[i * 4 + k] = (uint8_t)std::nearbyint(q);
        }
    }
    return VKRT_OK;
}

// The same pass onto a target of another size (window != backbuffer): the full-screen triangle's interpolated uv is
// the fragment centre over the target size (shaders/present.wgsl:98-104), `textureSample` is the bilinear,
// clamp-to-edge sampler of src/context/present_pipeline.rs:110-118 (fp32 weights, a + f (b - a), x then y).
int vko_present_scaled(const uint16_t* frame, int W, int H, int outW, int outH, uint8_t* rgba8) {
    if (!frame || !rgba8 || W <= 0 || H <= 0 || outW <= 0 || outH <= 0) return VKRT_ERR_INVALID;
    auto at = [&](int x, int y, int k) {
        x = x < 0 ? 0 : (x >= W ? W - 1 : x);
        y = y < 0 ? 0 : (y >= H ? H - 1 : y);
        return h2f(frame[((size_t)y * W + x) * 4 + k]);
    };
#pragma omp parallel for schedule(static)
    for (int oy = 0; oy < outH; ++oy) {
        for (int ox = 0; ox < outW; ++ox) {
            const float u = ((float)ox + 0.5f) / (float)outW, v = ((float)oy + 0.5f) / (float)outH;
            const float ux = u * (float)W - 0.5f, uy = v * (float)H - 0.5f;
            const float flx = std::floor(ux), fly = std::floor(uy);
            const float fx = ux - flx, fy = uy - fly;
            const int ix = f2i(flx), iy = f2i(fly);
            float c[4];
            for (int k = 0; k < 4; ++k) {
                const float a = at(ix, iy, k), b = at(ix + 1, iy, k), cc = at(ix, iy + 1, k), d = at(ix + 1, iy + 1, k);
                const float ab = a + fx * (b - a), cd = cc + fx * (d - cc);
                c[k] = ab + fy * (cd - ab);
            }
            for (int k = 0; k < 3; ++k) {
                const float x = c[k];
                const float t = wclamp((x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f), 0.0f, 1.0f);
                const float sel = std::ceil(t - 0.0031308f);
                c[k] = wmix(12.92f * t, 1.055f * std::pow(t, 0.41666f) - 0.055f, sel);
            }
            for (int k = 0; k < 4; ++k)
                rgba8[((size_t)oy * outW + ox) * 4 + k] = (uint8_t)std::nearbyint(wclamp(c[k], 0.0f, 1.0f) * 255.0f);
        }
    }
    return VKRT_OK;
}

// ---------------------------------------------------------------------------------------------
// src/camera.rs:93-113,148-171 with glam 0.20.5 semantics (crates.io; not vendored in the reference):
//   Mat4::look_at_rh(eye, center, up) = look_to_rh(eye, center - eye, up):
//       f = normalize(dir); s = normalize(cross(f, up)); u = cross(s, f);
//       cols: (s.x,u.x,-f.x,0) (s.y,u.y,-f.y,0) (s.z,u.z,-f.z,0) (-dot(s,eye), -dot(u,eye), dot(f,eye), 1)
//   Mat4::perspective_rh(fovy, aspect, near, far)  [depth 0..1]:
//       h = cos(fovy/2)/sin(fovy/2); w = h/aspect; r = far/(near-far);
//       cols: (w,0,0,0) (0,h,0,0) (0,0,r,-1) (0,0,r*near,0)
//   Mat4::inverse: cofactor expansion / determinant.
// glam normalises with v * (1/length); last-ulp differences from glam's SSE2 paths are possible and
// irrelevant to kernel parity because the 144-B uniform is an INPUT of the raycast boundary.
namespace {
inline v3 cross3(v3 a, v3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline v3 glam_normalize(v3 a) { float rl = 1.0f / std::sqrt(dot3(a, a)); return a * rl; }
void mat4_mul_mat4(const float* a, const float* b, float* o) {  // o = a*b, column-major
    for (int c = 0; c < 4; ++c) {
        v4 col = mat4_mul(a, v4{b[c * 4], b[c * 4 + 1], b[c * 4 + 2], b[c * 4 + 3]});
        o[c * 4] = col.x; o[c * 4 + 1] = col.y; o[c * 4 + 2] = col.z; o[c * 4 + 3] = col.w;
    }
}
bool mat4_inverse(const float* m, float* inv) {
    float t[16];
    t[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    t[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    t[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    t[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    t[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    t[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    t[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    t[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    t[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    t[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    t[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    t[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    t[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    t[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    t[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    t[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    float det = m[0] * t[0] + m[1] * t[4] + m[2] * t[8] + m[3] * t[12];
    if (det == 0.0f) return false;
    float rdet = 1.0f / det;
    for (int i = 0; i < 16; ++i) inv[i] = t[i] * rdet;
    return true;
}
}  // namespace

int vko_camera_uniform(float zoom, float pitch, float yaw, const float target[3], float aspect, VkrtCameraUniform* out) {
    if (!target || !out) return VKRT_ERR_INVALID;
    const float ZFAR = 100.0f, ZNEAR = 0.1f, FOVY = 3.14159265358979323846f / 2.0f;  // camera.rs:88-90
    // fix_eye (camera.rs:148-157)
    float pitch_cos = std::cos(pitch);
    v3 tgt = {target[0], target[1], target[2]};
    v3 eye = tgt - zoom * v3{std::sin(yaw) * pitch_cos, std::sin(pitch), std::cos(yaw) * pitch_cos};
    // look_at_rh
    v3 f = glam_normalize(tgt - eye);
    v3 s = glam_normalize(cross3(f, v3{0.0f, 1.0f, 0.0f}));
    v3 u = cross3(s, f);
    float view[16] = {s.x, u.x, -f.x, 0.0f, s.y, u.y, -f.y, 0.0f, s.z, u.z, -f.z, 0.0f, -dot3(s, eye), -dot3(u, eye), dot3(f, eye), 1.0f};
    // perspective_rh
    float sf = std::sin(0.5f * FOVY), cf = std::cos(0.5f * FOVY);
    float h = cf / sf, w = h / aspect, r = ZFAR / (ZNEAR - ZFAR);
    float proj[16] = {w, 0, 0, 0, 0, h, 0, 0, 0, 0, r, -1.0f, 0, 0, r * ZNEAR, 0};
    float pv[16], inv[16];
    mat4_mul_mat4(proj, view, pv);  // camera.rs:112 `proj * view`
    if (!mat4_inverse(pv, inv)) return VKRT_ERR_INVALID;
    out->view_position[0] = eye.x; out->view_position[1] = eye.y; out->view_position[2] = eye.z; out->view_position[3] = 1.0f;
    std::memcpy(out->proj_view, pv, sizeof pv);
    std::memcpy(out->inv_proj, inv, sizeof inv);
    return VKRT_OK;
}

}  // extern "C"
