// vkrt_device.cuh — device-side building blocks shared by the raycast kernels (sm_100a).
//
// Arithmetic policy. Everything that decides WHICH texel a sample reads or WHETHER a
// ray hits — ray generation, the slab test, the t sequence, the sample position and the voxel
// index — is written with explicitly rounded intrinsics (__fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn),
// which nvcc never contracts into FMAs, in exactly the operation order the oracle fixes. Hit masks
// and voxel indices are therefore bit-identical to the oracle's. Shading and compositing use
// ordinary fp32 (FMA contraction allowed) and are checked against the stated tolerance.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vokselis_rt.h"

namespace vkrt {

struct f3 {
    float x, y, z;
};

// ---- exact helpers ------------------------------------------------------------------------
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xdot3(f3 a, f3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }

// column-major mat4 * (x, y, z, w) with the oracle's summation order ((c0*x + c1*y) + c2*z) + c3*w
__device__ __forceinline__ float xrow(const float* m, int r, float x, float y, float z, float w) {
    return xadd(xadd(xadd(xmul(m[r], x), xmul(m[4 + r], y)), xmul(m[8 + r], z)), xmul(m[12 + r], w));
}

// shaders/raycast_compute.wgsl:102-116 — ray generation (`render`). gx, gy = global_invocation_id. aspect_ratio = the
// IEEE quotient H / W (:105), the same for every pixel: the raycast kernel takes it from the host (RenderArgs::aspect_hw).
__device__ __forceinline__ void gen_ray(const float* inv, float gx, float gy, float offx, float offy, float W, float H, float aspect_ratio,
                                        f3& eye, f3& dir) {
    const float cx = xadd(gx, offx), cy = xadd(gy, offy);
    const float sx = xsub(xdiv(xmul(2.0f, cx), W), 1.0f);
    float sy = xsub(xdiv(xmul(2.0f, cy), H), 1.0f);
    sy = xmul(sy, -aspect_ratio);
    // screen_point = (sx, sy, 0, 1); screen_tangent = screen_point + (0,0,1,0) = (sx, sy, 1, 1)
    f3 vp, vt;
    vp.x = xrow(inv, 0, sx, sy, 0.0f, 1.0f); vp.y = xrow(inv, 1, sx, sy, 0.0f, 1.0f); vp.z = xrow(inv, 2, sx, sy, 0.0f, 1.0f);
    const float vpw = xrow(inv, 3, sx, sy, 0.0f, 1.0f);
    vt.x = xrow(inv, 0, sx, sy, 1.0f, 1.0f); vt.y = xrow(inv, 1, sx, sy, 1.0f, 1.0f); vt.z = xrow(inv, 2, sx, sy, 1.0f, 1.0f);
    const float vtw = xrow(inv, 3, sx, sy, 1.0f, 1.0f);
    eye.x = xdiv(vp.x, vpw); eye.y = xdiv(vp.y, vpw); eye.z = xdiv(vp.z, vpw);
    f3 d = {xsub(xdiv(vt.x, vtw), eye.x), xsub(xdiv(vt.y, vtw), eye.y), xsub(xdiv(vt.z, vtw), eye.z)};
    const float len = __fsqrt_rn(xdot3(d, d));
    dir.x = xdiv(d.x, len); dir.y = xdiv(d.y, len); dir.z = xdiv(d.z, len);
}

__device__ __forceinline__ void gen_ray(const float* inv, float gx, float gy, float offx, float offy, float W, float H, f3& eye, f3& dir) {
    gen_ray(inv, gx, gy, offx, offy, W, H, xdiv(H, W), eye, dir);
}

// shaders/raycast_compute.wgsl:42-53 — slab test against [-1,1]^3 (fminf/fmaxf = WGSL min/max on NVIDIA).
__device__ __forceinline__ void intersect_box(f3 o, f3 d, float& t0, float& t1, f3& inv) {
    const float ix = xdiv(1.0f, d.x), iy = xdiv(1.0f, d.y), iz = xdiv(1.0f, d.z);
    inv.x = ix; inv.y = iy; inv.z = iz;
    const float ax = xmul(xsub(-1.0f, o.x), ix), ay = xmul(xsub(-1.0f, o.y), iy), az = xmul(xsub(-1.0f, o.z), iz);
    const float bx = xmul(xsub(1.0f, o.x), ix), by = xmul(xsub(1.0f, o.y), iy), bz = xmul(xsub(1.0f, o.z), iz);
    t0 = fmaxf(fminf(ax, bx), fmaxf(fminf(ay, by), fminf(az, bz)));
    t1 = fminf(fmaxf(ax, bx), fminf(fmaxf(ay, by), fmaxf(az, bz)));
}

__device__ __forceinline__ void intersect_box(f3 o, f3 d, float& t0, float& t1) {
    f3 inv;
    intersect_box(o, d, t0, t1, inv);
}

// The same slab test against an arbitrary box (used to clip rays to the occupied bounds; approximate
// arithmetic is fine there, the box carries a margin).
__device__ __forceinline__ void slab_box(f3 o, f3 inv, const float* lo, const float* hi, float& t0, float& t1) {
    const float ax = (lo[0] - o.x) * inv.x, ay = (lo[1] - o.y) * inv.y, az = (lo[2] - o.z) * inv.z;
    const float bx = (hi[0] - o.x) * inv.x, by = (hi[1] - o.y) * inv.y, bz = (hi[2] - o.z) * inv.z;
    t0 = fmaxf(fminf(ax, bx), fmaxf(fminf(ay, by), fminf(az, bz)));
    t1 = fminf(fmaxf(ax, bx), fminf(fmaxf(ay, by), fmaxf(az, bz)));
}

// shaders/raycast_compute.wgsl:65-68 — step length: dt_scale * max(min_i 1 / (n_i |dir_i|), dt_floor). Correctly rounded
// division is monotone, so the minimum of the three quotients IS the quotient by the largest divisor, bit for bit — one
// IEEE division instead of three (fminf / fmaxf both drop a NaN operand; an axis with |dir_i| = 0 gives +inf either way).
__device__ __forceinline__ float step_dt(f3 dir, float nx, float ny, float nz, float dt_scale, float dt_floor) {
    const float a = fmaxf(xmul(nx, fabsf(dir.x)), fmaxf(xmul(ny, fabsf(dir.y)), xmul(nz, fabsf(dir.z))));
    return xmul(dt_scale, fmaxf(xdiv(1.0f, a), dt_floor));
}

// Exact result of `n` repeated additions t = fl(t + dt) (round-to-nearest-even), without executing them
// one by one. Inside one binade every addition moves the 24-bit significand of t by the same integer
// (dt expressed in ulps of t, rounded once), so the whole run is one integer multiply-add on the bit
// pattern; binade crossings and the rare exact-tie case fall back to a real FADD for that step.
// Requires t >= 0, dt > 0 (finite). Used to leap over empty space / other ranks' bricks while landing
// on exactly the floats the reference's `t = t + dt` loop visits.
__device__ __forceinline__ float advance_t(float t, float dt, int n) {
    const uint32_t db = __float_as_uint(dt);
    const int e_d = (int)(db >> 23);
    const uint32_t M = (db & 0x7fffffu) | 0x800000u;
    while (n > 0) {
        const uint32_t tb = __float_as_uint(t);
        const int e = (int)(tb >> 23);
        const int shift = e - e_d;
        // t in a lower binade than dt (incl. t == 0 / subnormal), or dt subnormal: plain step
        if (shift < 0 || e_d == 0 || e == 0) {
            t = xadd(t, dt);
            --n;
            continue;
        }
        if (shift > 24) return t;  // dt < ulp(t)/2: the reference loop would not advance either
        uint32_t inc;
        if (shift == 0) {
            inc = M;
        } else {
            const uint32_t m = M >> shift, rem = M & ((1u << shift) - 1u), half = 1u << (shift - 1);
            if (rem == half) {  // exact tie: result depends on the parity of t; take this step for real
                t = xadd(t, dt);
                --n;
                continue;
            }
            inc = m + (rem > half ? 1u : 0u);
        }
        const uint32_t T = (tb & 0x7fffffu) | 0x800000u;
        const uint32_t room = inc ? (0xffffffu - T) / inc : 0u;  // steps that stay inside this binade
        const uint32_t k = min((uint32_t)n, room);
        if (k == 0u) {  // the next step crosses into the next binade
            t = xadd(t, dt);
            --n;
            continue;
        }
        t = __uint_as_float(tb + k * inc);
        n -= (int)k;
    }
    return t;
}

// leap_steps: the same result — t after n additions t = fl(t + dt), bit for bit — for the raycast loop, where a ray
// leaps many times (short leaps mostly) and crosses one or two binades of t inside the volume:
//  * the per-binade increment is cached in registers across the leaps of a ray (LeapCache) when it is plain (no tie)
//    and below 2^19, so that n * inc with n <= kLeapFastMax cannot leave 31 bits;
//  * a leap that stays inside the cached binade is then ONE 32-bit multiply-add on the bit pattern, one LOP3
//    ((t ^ base) | (result ^ base): both in the binade?) and a compare;
//  * everything else goes to leap_steps_slow: a leap that reaches the end of the binade takes as many steps as provably
//    stay inside (quotient by a float reciprocal, one below the estimate, verified in integers — no integer division),
//    then real additions across the boundary, then goes on in the next binade; an exact tie (dt = (m + 1/2) ulp(t))
//    rounds to even: from an even significand every step adds the even one of {m, m+1}, so after at most one real
//    addition the closed form applies there too; t below dt's binade, subnormals: real additions.
// binade_inc: < 2^24 = increment; bit 31 set = tie binade (low bits: the even increment); 0xffffffff = no closed form.
constexpr int kLeapFastMax = 4095;
struct LeapCache {
    uint32_t eb, inc;  // eb = e << 23 for the binade `inc` belongs to (0xffffffff = none cached); inc < 2^19, plain
};
__device__ __forceinline__ uint32_t binade_inc(uint32_t e, float dt) {
    const uint32_t db = __float_as_uint(dt), ed = db >> 23;
    const int shift = (int)e - (int)ed;
    if (shift < 0 || shift > 24 || ed == 0u || e == 0u) return 0xffffffffu;
    const uint32_t M = (db & 0x7fffffu) | 0x800000u;
    if (shift == 0) return M;
    const uint32_t m = M >> shift, rem = M & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem == half) return 0x80000000u | (m + (m & 1u));
    return m + (rem > half ? 1u : 0u);  // >= 1 for shift <= 24
}
__device__ __forceinline__ float leap_steps_slow(float t, float dt, int n, LeapCache& c) {
    for (;;) {
        const uint32_t tb = __float_as_uint(t), e = tb >> 23;  // t >= 0
        uint32_t inc;
        if ((e << 23) == c.eb) {
            inc = c.inc;
        } else {
            inc = binade_inc(e, dt);
            if (inc < (1u << 19)) {
                c.eb = e << 23;
                c.inc = inc;
            }
        }
        const bool closed = inc != 0xffffffffu && !((inc >> 31) && (tb & 1u));  // tie binade: only from an even significand
        if (closed) {
            inc &= 0x7fffffffu;
            const unsigned long long nb = (unsigned long long)tb + (unsigned long long)(uint32_t)n * inc;
            if ((nb >> 23) == (unsigned long long)e) return __uint_as_float((uint32_t)nb);
            // the run leaves this binade: steps that provably stay inside (one below the float quotient, verified)
            const uint32_t room = (((e + 1u) << 23) - 1u) - tb;
            const int m = __float2int_rz(__fdividef((float)room, (float)inc)) - 1;
            if (m >= 1 && m < n && (unsigned long long)(uint32_t)m * inc <= (unsigned long long)room) {
                t = __uint_as_float(tb + (uint32_t)m * inc);
                n -= m;
            }
        }
        t = xadd(t, dt);  // a real addition: towards / across the boundary, off an odd tie significand, below dt's binade
        if (--n == 0) return t;
    }
}
// requires 1 <= n <= kLeapFastMax, t >= 0
__device__ __forceinline__ float leap_steps(float t, float dt, int n, LeapCache& c) {
    const uint32_t tb = __float_as_uint(t);
    const uint32_t nb = tb + (uint32_t)n * c.inc;  // < 2^32: tb < 2^31, n * inc < 2^12 * 2^19
    if ((((tb ^ c.eb) | (nb ^ c.eb)) >> 23) == 0u) return __uint_as_float(nb);  // t and the result both in the cached binade
    return leap_steps_slow(t, dt, n, c);
}

// ---- shading (tolerance-checked, FMA allowed) ----------------------------------------------
// smoothstep with compile-time edges: the division by (e1 - e0) becomes a multiplication by its
// reciprocal (an IEEE fdiv is ~10 instructions plus a slow path for zero numerators; profiles/).
// __saturatef maps NaN -> 0 like fmin(fmax(NaN, 0), 1).
#define VKRT_SMOOTHSTEP(e0, e1, x) vkrt::smoothstep_r((e0), 1.0f / ((e1) - (e0)), (x))
// NOTE on determinism: the shading below spells out every multiply-add (fmaf / __fmul_rn / __fadd_rn).
// Left to the compiler, FMA contraction differs between template instantiations of the kernel (SKIP vs
// not, DBG vs not), which showed up as one pixel in 10^6 differing by one fp16 ulp between the skipping and
// the non-skipping kernel. With explicit operations all instantiations and layouts produce identical bits.
__device__ __forceinline__ float smoothstep_r(float e0, float inv_range, float x) {
    const float t = __saturatef(__fmul_rn(__fsub_rn(x, e0), inv_range));
    return __fmul_rn(__fmul_rn(t, t), fmaf(-2.0f, t, 3.0f));
}

struct Rgba {
    float r, g, b, a;
};

// shaders/raycast_compute.wgsl:74-91 — one sample of `get_col2`. c = volume texel, n = normal texel
// (n.w unused), p = sample position.
__device__ __forceinline__ float m0_alpha(float ca) {
    // pow(a, 3.0) then smoothstep(0, 0.7, .): x*x*x is within 1 ulp of the exact cube (CUDA powf: 4 ulp).
    const float a3 = __fmul_rn(__fmul_rn(ca, ca), ca);
    return VKRT_SMOOTHSTEP(0.0f, 0.7f, a3);
}

// A texel a sample may be skipped over (exact empty-space skipping): per-sample alpha exactly 0 AND
// nothing non-finite that 0 * (.) would turn into NaN in the full march — colour inf/NaN, normal +-inf (NaN normals
// are harmless: every use goes through fmaxf / __saturatef, and the generator fills empty space with them).
__device__ __forceinline__ bool m0_texel_skippable(uint2 c, uint2 n) {
    const float a = __half2float(__ushort_as_half((unsigned short)(c.y >> 16)));
    if (m0_alpha(a) != 0.0f) return false;
    const uint32_t cx = c.x & 0x7C00u, cy = (c.x >> 16) & 0x7C00u, cz = c.y & 0x7C00u;
    if (cx == 0x7C00u || cy == 0x7C00u || cz == 0x7C00u) return false;
    const uint32_t nx = n.x & 0x7FFFu, ny = (n.x >> 16) & 0x7FFFu, nz = n.y & 0x7FFFu;
    return !(nx == 0x7C00u || ny == 0x7C00u || nz == 0x7C00u);
}

__device__ __forceinline__ void m0_shade(Rgba& col, float4 c, float4 n, f3 p, const float* clear) {
    const float kL = 0.33333334f;  // normalize(-2,-2,-1) = (-2/3, -2/3, -1/3)
    const float kP = 0.57735026f;  // normalize(1,1,-1) = (1,1,-1)/sqrt(3)
    // dot((0,-1,0), n) summed like the oracle, (0*n.x + -1*n.y) + 0*n.z: a non-finite n.x or n.z makes it NaN, which
    // fmaxf drops (raycast_compute.wgsl:74 on golden G2's inf normals)
    const float shade_s = fmaxf(0.0f, __fadd_rn(__fsub_rn(__fmul_rn(0.0f, n.x), n.y), __fmul_rn(0.0f, n.z)));
    const float vol_alpha = m0_alpha(c.w);
    const float ndl = fmaxf(fmaf(-kL, n.z, fmaf(-2.0f * kL, n.y, __fmul_rn(-2.0f * kL, n.x))), 0.0f);
    const float pd = VKRT_SMOOTHSTEP(0.3f, 1.5f, fmaf(-kP, p.z, fmaf(kP, p.y, __fmul_rn(kP, p.x))));
    const float dsc = __fmul_rn(ndl, pd);
    const float vr = fmaf(3.0f, dsc, c.x), vg = fmaf(0.3f, dsc, c.y), vb = fmaf(0.39f, dsc, c.z);
    const float bottom = __fmul_rn(0.9f, __saturatef(fmaf(-0.5f, n.y, 0.5f)));
    const float sh_rg = __fmul_rn(shade_s, 0.8f);                                   // mix(shade, 0, 0.2)
    const float sh_b = fmaf(__fmul_rn(bottom, 0.6f), 0.2f, __fmul_rn(shade_s, 0.8f));  // mix(shade, bottom*0.6, 0.2)
    const float w = __fmul_rn(__fsub_rn(1.0f, col.a), vol_alpha);
    const float k = __fmul_rn(clear[3], __fsub_rn(1.0f, vol_alpha));
    col.r = fmaf(clear[0], k, fmaf(__fmul_rn(w, vr), sh_rg, col.r));
    col.g = fmaf(clear[1], k, fmaf(__fmul_rn(w, vg), sh_rg, col.g));
    col.b = fmaf(clear[2], k, fmaf(__fmul_rn(w, vb), sh_b, col.b));
    col.a = fmaf(w, __fsub_rn(1.0f, clear[3]), col.a);
}

// shaders/raycast_naive.wgsl:70-81,106-117 — transfer function + composite of one scalar sample.
__device__ __forceinline__ float m1_alpha(float s) { return VKRT_SMOOTHSTEP(0.10f, 1.2f, fminf(0.9f, s)); }

__device__ __forceinline__ void m1_shade(Rgba& col, float s) {
    const float TAU = 6.28318f;
    const float v = m1_alpha(s);
    // v == 0 (s <= 0.1) gives w = 0, and the palette is finite: the sample leaves colour and alpha bit-identical without a
    // branch. (Round 1 returned early here to save the three cosines; with 2-voxel occupancy bricks few fetched samples are
    // transparent and a warp skips the palette only when ALL its lanes are: the branch cost more than it saved, -1 %.)
    // vertigo palette 0.5 + 0.5 cos(TAU (c v + d)): TAU folded into the constants (colour is tolerance-checked)
    const float pr = fmaf(0.5f, __cosf(__fmul_rn(TAU, v)), 0.5f);
    const float pg = fmaf(0.5f, __cosf(fmaf(TAU * 1.7f, v, TAU * 0.15f)), 0.5f);
    const float pb = fmaf(0.5f, __cosf(fmaf(TAU * 0.4f, v, TAU * 0.20f)), 0.5f);
    const float w = __fmul_rn(__fsub_rn(1.0f, col.a), v);
    col.r = fmaf(w, pr, col.r);
    col.g = fmaf(w, pg, col.g);
    col.b = fmaf(w, pb, col.b);
    col.a = __fadd_rn(col.a, w);
}

// The same sample for the raycast loop, with the palette's constant half taken out of the loop: the sum over the samples of
// w (0.5 + 0.5 cos) is 0.5 (sum of w) + 0.5 (sum of w cos), and the sum of w is the alpha gained, so the loop accumulates
// only w cos(.) per channel (one FFMA instead of two, no 0.5 held in a register: four issue slots fewer per sample) and
// m1_finish adds the constant half once per ray. Alpha — what termination and the iteration count depend on — is computed
// exactly as in m1_shade; colours move by ~1e-6 (summation order; the alpha gained carries alpha's own rounding), far
// inside the 2/255 tolerance, and a transparent sample (w = 0) still leaves every accumulator bit-identical.
// green / blue phase = TAU (c v + d) as pairs in the constant bank: one FFMA2 on two LDC.64 operands instead of two FFMAs and
// two constants materialised in registers per sample
static __constant__ float2 kPalScale = {6.28318f * 1.7f, 6.28318f * 0.4f}, kPalPhase = {6.28318f * 0.15f, 6.28318f * 0.20f};
__device__ __forceinline__ void m1_shade_acc(Rgba& col, float s) {
    const float TAU = 6.28318f;
    const float v = m1_alpha(s);
    const float cr = __cosf(__fmul_rn(TAU, v));
    float2 gb;  // green and blue phases in one FFMA2 (the same two IEEE fused multiply-adds)
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %2}; mov.b64 rb, {%3, %4}; mov.b64 rc, {%5, %6}; fma.rn.f32x2 rd, ra, rb, rc; "
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(gb.x), "=f"(gb.y)
        : "f"(v), "f"(kPalScale.x), "f"(kPalScale.y), "f"(kPalPhase.x), "f"(kPalPhase.y));
    const float cg = __cosf(gb.x), cb = __cosf(gb.y);
    const float w = __fmul_rn(__fsub_rn(1.0f, col.a), v);
    {   // red and green in one packed FFMA2 (the same two IEEE fused multiply-adds)
        float2 rg;
        asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %4}; mov.b64 rc, {%5, %6}; fma.rn.f32x2 rd, ra, rb, rc; "
            "mov.b64 {%0, %1}, rd;}"
            : "=f"(rg.x), "=f"(rg.y)
            : "f"(cr), "f"(cg), "f"(w), "f"(col.r), "f"(col.g));
        col.r = rg.x;
        col.g = rg.y;
    }
    col.b = fmaf(w, cb, col.b);
    col.a = __fadd_rn(col.a, w);
}
__device__ __forceinline__ void m1_finish(Rgba& col, float initial_alpha) {
    const float half_gain = __fmul_rn(0.5f, __fsub_rn(col.a, initial_alpha));
    col.r = fmaf(0.5f, col.r, half_gain);
    col.g = fmaf(0.5f, col.g, half_gain);
    col.b = fmaf(0.5f, col.b, half_gain);
}

// shaders/raycast_naive.wgsl:63-68
__device__ __forceinline__ float linear_to_srgb_naive(float x) {
    return x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
}

__device__ __forceinline__ uint2 pack_rgba16f(float r, float g, float b, float a) {
    __half2 lo = __floats2half2_rn(r, g), hi = __floats2half2_rn(b, a);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    return o;
}
__device__ __forceinline__ float4 unpack_rgba16f(uint2 v) {
    const __half2 lo = *reinterpret_cast<const __half2*>(&v.x), hi = *reinterpret_cast<const __half2*>(&v.y);
    const float2 a = __half22float2(lo), b = __half22float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}

// ---- present (shaders/present.wgsl:23-35,111-119): ACESFilm + linear_to_srgb + unorm8 pack --------------
// Every multiply-add is spelled out so that the stand-alone present pass (present.cu) and the raycast kernel's
// fused epilogue produce identical bytes.
__device__ __forceinline__ float present_channel(float x) {
    // ACESFilm; clamp via fminf/fmaxf so NaN -> 0 like the oracle
    const float num = __fmul_rn(x, fmaf(2.51f, x, 0.03f)), den = fmaf(x, fmaf(2.43f, x, 0.59f), 0.14f);
    const float v = fminf(fmaxf(__fdiv_rn(num, den), 0.0f), 1.0f);
    // linear_to_srgb: selector = ceil(v - 0.0031308) in {0,1}; mix(under, over, selector)
    const float sel = ceilf(__fsub_rn(v, 0.0031308f));
    const float under = __fmul_rn(12.92f, v);
    const float over = fmaf(1.055f, powf(v, 0.41666f), -0.055f);
    return fmaf(over, sel, __fmul_rn(under, __fsub_rn(1.0f, sel)));
}
__device__ __forceinline__ uint32_t unorm8(float x) { return (uint32_t)__float2int_rn(__fmul_rn(__saturatef(x), 255.0f)); }
// packed = one rgba16f texel of the frame (what textureStore wrote; the present pass samples that, not the fp32 colour)
__device__ __forceinline__ uint32_t present_pixel(uint2 packed) {
    const __half2 lo = *reinterpret_cast<const __half2*>(&packed.x), hi = *reinterpret_cast<const __half2*>(&packed.y);
    const float2 a = __half22float2(lo), b = __half22float2(hi);
    return unorm8(present_channel(a.x)) | unorm8(present_channel(a.y)) << 8 | unorm8(present_channel(b.x)) << 16 | unorm8(b.y) << 24;
}

// ---- bricked layout address --------------------------------------------------
// M0 interleaved texel = 16 B (colour rgba16f + normal rgba16f). 2x2x2 texels fill one 128-B line;
// 4x4x4 lines (8^3 voxels, 8 KB) form a brick; bricks are x-fastest. nbx, nby = bricks per axis.
__device__ __host__ __forceinline__ uint32_t bricked_index(int ix, int iy, int iz, int nbx, int nby) {
    const uint32_t brick = ((uint32_t)(iz >> 3) * (uint32_t)nby + (uint32_t)(iy >> 3)) * (uint32_t)nbx + (uint32_t)(ix >> 3);
    const uint32_t line = (((uint32_t)iz >> 1) & 3u) << 4 | (((uint32_t)iy >> 1) & 3u) << 2 | (((uint32_t)ix >> 1) & 3u);
    const uint32_t in_line = ((uint32_t)iz & 1u) << 2 | ((uint32_t)iy & 1u) << 1 | ((uint32_t)ix & 1u);
    return (brick << 9) | (line << 3) | in_line;
}

}  // namespace vkrt
