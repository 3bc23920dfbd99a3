// raycast.cu — the hot path: one thread per pixel, warp-coherent 8x4 pixel tiles (sm_100a).
//
// Replaces shaders/raycast_compute.wgsl entry points `single` (:133-137) and `tile` (:139-144):
// ray generation + slab test (`render` :99-131, `intersect_box` :42-53), fixed-step march with
// shading, front-to-back compositing and early ray termination (`get_col2` :62-97). Mode M1 swaps
// the march body for the scalar/trilinear/transfer-function body of shaders/raycast_naive.wgsl:96-119.
//
// B200 mapping: a block is 8x16 pixels = four warps of 8x4 pixels, so the 32 rays of a
// warp walk through neighbouring voxels and their texel requests fall into a handful of 32-B sectors
// (L1 hit rate 94-98 % on the 256^3 u8 volume). All tiles of a frame go out in ONE launch (grid.z = tile), or the
// cameras of up to VKRT_MAX_BATCH consecutive frames of a sweep (grid.y = frame): one 1080p frame cannot fill 148 SMs.
// Not a contraction: no tensor cores. Measured limiters (profiles/): instruction issue when skipping is on
// (issue active 83 %, SM throughput 82 % of peak at 16-30 frames per launch), the texel path in the dense case (l1tex 98 %).
#include <cstdio>
#include <cstdlib>

#include "raycast.cuh"
#include "vkrt_device.cuh"

namespace vkrt {

namespace {

// Scalar taps of the LINEAR layout, per storage type.
template <int DTYPE> struct Scalar;
template <> struct Scalar<VKRT_U8> {
    // R8Unorm: value = b / 255. The byte is turned into a float on the FMA pipe (0x4B000000 | b is
    // 2^23 + b; no I2F on the XU pipe, no per-tap IEEE division); the 1/255 is applied once, after
    // the interpolation (differs from the oracle's per-tap b/255 only in the last ulp).
    static constexpr float kScale = 1.0f / 255.0f;
    static __device__ __forceinline__ float load(const void* p, size_t i) {
        return __uint_as_float(0x4B000000u | (uint32_t)__ldg((const uint8_t*)p + i)) - 8388608.0f;
    }
};
template <> struct Scalar<VKRT_F16> {
    static constexpr float kScale = 1.0f;
    static __device__ __forceinline__ float load(const void* p, size_t i) {
        return __half2float(__ushort_as_half(__ldg((const unsigned short*)p + i)));
    }
};
template <> struct Scalar<VKRT_F32> {
    static constexpr float kScale = 1.0f;
    static __device__ __forceinline__ float load(const void* p, size_t i) { return __ldg((const float*)p + i); }
};

// Distance field over the 8^3-voxel bricks (A.dist): 0 = some sample in the brick can be non-transparent;
// d >= 1 = every brick within Chebyshev radius d-1 (this one included) is empty.

// ---- texel fetch, M0 (nearest, two rgba16f texels at one integer coordinate) -----------------
template <int LAYOUT>
__device__ __forceinline__ void m0_fetch(const RenderArgs& A, int ix, int iy, int iz, bool inb, float4& c, float4& n) {
    if (LAYOUT == VKRT_LAYOUT_TEXTURE) {
        // border addressing returns 0 outside: the out-of-range textureLoad definition.
        // (tex3DLod with level 0, here and below: TEX.LZ, no level-of-detail register to set up per fetch — tex3D compiles
        // to TEX.LL with a register holding the level; the arrays have one level, the texels are the same)
        c = tex3DLod<float4>(A.tex_a, (float)ix + 0.5f, (float)iy + 0.5f, (float)iz + 0.5f, 0.0f);
        n = tex3DLod<float4>(A.tex_b, (float)ix + 0.5f, (float)iy + 0.5f, (float)iz + 0.5f, 0.0f);
        return;
    }
    if (!inb) {
        c = make_float4(0.f, 0.f, 0.f, 0.f);
        n = c;
        return;
    }
    if (LAYOUT == VKRT_LAYOUT_BRICKED) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(A.vol_a) + bricked_index(ix, iy, iz, A.nbx, A.nby));
        c = unpack_rgba16f(make_uint2(v.x, v.y));
        n = unpack_rgba16f(make_uint2(v.z, v.w));
    } else {
        const size_t i = ((size_t)iz * A.ny + iy) * A.nx + ix;
        c = unpack_rgba16f(__ldg(reinterpret_cast<const uint2*>(A.vol_a) + i));
        n = unpack_rgba16f(__ldg(reinterpret_cast<const uint2*>(A.vol_b) + i));
    }
}

// tld4 on a 2-D layered texture: the four texels of the bilinear footprint around (x, y) of `layer`,
// returned as (x0,y1), (x1,y1), (x1,y0), (x0,y0). CUDA C++ only exposes gather for plain 2-D textures.
__device__ __forceinline__ float4 tld4_layer(cudaTextureObject_t tex, int layer, float x, float y) {
    float4 r;
    asm("tld4.r.a2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(tex), "r"(layer), "f"(x), "f"(y));
    return r;
}

// Packed fp32x2 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2): two IEEE round-to-nearest operations per issue slot, each
// component rounded exactly like the scalar __fadd_rn / __fmul_rn / __fmaf_rn, never contracted. The kernel is bound by
// instruction issue, so the x/y halves of the position arithmetic and the pairs of the trilinear
// interpolation go through these; ptxas broadcasts a scalar or an immediate to both halves for free.
// (Inline PTX, not the __fadd2_rn / __fmul2_rn intrinsics: nvcc contracts those two into one FFMA2, which would change
// the bits of p = eye + t*dir; `.rn` PTX instructions are never fused.)
#define VKRT_F32X2_OP2(name, op)                                                                                         \
    __device__ __forceinline__ float2 name(float2 a, float2 b) {                                                         \
        float2 r;                                                                                                        \
        asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; " op " rc, ra, rb; mov.b64 {%0, %1}, rc;}" \
            : "=f"(r.x), "=f"(r.y)                                                                                       \
            : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                                                                   \
        return r;                                                                                                        \
    }
VKRT_F32X2_OP2(add2, "add.rn.f32x2")
VKRT_F32X2_OP2(mul2, "mul.rn.f32x2")
#undef VKRT_F32X2_OP2
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return add2(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; "
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
// the same, rounded towards -infinity (FFMA2.RM)
__device__ __forceinline__ float2 fma2_rm(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rm.f32x2 rd, ra, rb, rc; "
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }
// 1.5 * 2^23: RD(x + kMagic) = kMagic + floor(x) for |x| < 2^22, an integer-valued float whose bit pattern is
// 0x4B400000 + floor(x) (two's complement for negative floors) — floor and float-to-int on the FMA pipe, no F2I / I2F
constexpr float kMagic = 12582912.0f;

// a + f*(b - a), every operation spelled out (see the determinism note in vkrt_device.cuh). Order: z, then y, then x —
// the order in which the pre-gathered quads pair up for the packed form; every layout uses it, so all layouts
// produce identical bits (the oracle interpolates x, y, z: same value up to rounding, tolerance-checked).
__device__ __forceinline__ float lerp1(float a, float b, float f) { return fmaf(f, __fsub_rn(b, a), a); }
__device__ __forceinline__ float2 lerp2(float2 a, float2 b, float2 f) { return fma2(f, sub2(b, a), a); }
__device__ __forceinline__ float lerp3(float c000, float c100, float c010, float c110, float c001, float c101, float c011, float c111,
                                       float fx, float fy, float fz) {
    const float c00 = lerp1(c000, c001, fz), c10 = lerp1(c100, c101, fz), c01 = lerp1(c010, c011, fz), c11 = lerp1(c110, c111, fz);
    return lerp1(lerp1(c00, c01, fy), lerp1(c10, c11, fy), fx);
}
// the same on two texels of the QUAD layout: g = (v(x0,y0), v(x1,y0), v(x0,y1), v(x1,y1)) at z0 and z1
__device__ __forceinline__ float lerp3_quads(float4 g0, float4 g1, float fx, float fy, float fz) {
    const float2 fz2 = dup2(fz);
    const float2 a = lerp2(make_float2(g0.x, g0.y), make_float2(g1.x, g1.y), fz2);  // (c00, c10)
    const float2 b = lerp2(make_float2(g0.z, g0.w), make_float2(g1.z, g1.w), fz2);  // (c01, c11)
    const float2 c = lerp2(a, b, dup2(fy));                                         // (c0, c1)
    return lerp1(c.x, c.y, fx);
}

// ---- scalar sample, M1 (linear filter, clamp-to-edge) ----------------------------------------
template <int LAYOUT, int DTYPE>
__device__ __forceinline__ float m1_sample(const RenderArgs& A, float2 qxy, float qz) {
    if (LAYOUT == VKRT_LAYOUT_TEXTURE) {
        return tex3DLod<float>(A.tex_a, qxy.x, qxy.y, qz, 0.0f);  // hardware trilinear, 8-bit weights
    }
    // q - 0.5 as RN(q * 1 + -0.5): qxy comes out of a packed multiplication, and ptxas would contract a plain packed
    // addition with it (see RenderArgs::one), rounding u once instead of twice
    const float2 uxy = fma2(qxy, dup2(A.one), dup2(-0.5f));
    const float uz = __fsub_rn(qz, 0.5f);
    // (floorf is FRND on the XU pipe; the FMA-pipe form — add 1.5*2^23 rounding down, subtract it again, FADD2.RM — was
    // measured 0.5 % slower: one more issue slot per sample, and the XU pipe is not the limiter)
    const float2 flxy = make_float2(floorf(uxy.x), floorf(uxy.y));
    const float flz = floorf(uz);
    const float2 fxy = sub2(uxy, flxy);
    const float fx = fxy.x, fy = fxy.y, fz = __fsub_rn(uz, flz);
    const float flx = flxy.x, fly = flxy.y;
    if (LAYOUT == VKRT_LAYOUT_GATHER) {
        // (Measured, ncu: in the dense case this path is bound by the texture DATA pipe — l1tex throughput
        // 97 %, data_pipe_tex_wavefronts 83 %, ~35 sectors per warp-level tld4 because the lanes of a warp sit
        // in different layers. Packing (z, z+1) pairs into two-channel texels so that both gathers share a
        // footprint was tried and changed nothing: profiles/r01_gather_notes.md.)
        // The 8 taps in two texture instructions; the gather point sits exactly between the four texels,
        // far from any footprint-selection boundary. x/y clamp-to-edge is the texture's address mode,
        // z is clamped here. Weights stay fp32 (unlike tex3D's 8-bit weights): same result as LINEAR.
        const int z0 = (int)flz;
        const int za = min(max(z0, 0), A.nz - 1), zb = min(max(z0 + 1, 0), A.nz - 1);
        const float4 g0 = tld4_layer(A.tex_a, za, flx + 1.0f, fly + 1.0f);
        const float4 g1 = tld4_layer(A.tex_a, zb, flx + 1.0f, fly + 1.0f);
        return lerp3(g0.w, g0.z, g0.x, g0.y, g1.w, g1.z, g1.x, g1.y, fx, fy, fz);
    }
    if (LAYOUT == VKRT_LAYOUT_QUAD) {
        // Texel (x0+1, y0+1, z) holds the pre-gathered 2x2 xy footprint v(x0..x0+1, y0..y0+1, z) with clamp-to-edge baked
        // in (volume.cu pregather_quads_kernel); z clamps through the texture's address mode. Two POINT fetches of a 3-D
        // texture return the 8 taps; no gather footprint, 3-D tiled locality across layers.
        // Both fetches read ONE coordinate triple; the second adds the texel offset (0, 0, 1) inside the texture unit
        // (TEX.LZ.AOFFI; the offset is applied before the address mode, so z + 1 clamps to the edge like a coordinate would):
        // one packed constant instead of a second coordinate triple (FADD2 + FADD), and level 0 without a level register
        // (TEX.LZ). CUDA C++ has no texel-offset fetch for 3-D textures, hence the PTX. Same texels, same bits.
        const float2 cxy = add2(flxy, dup2(1.5f));
        const float cz = flz + 0.5f;
        float4 g0, g1;
        asm("tex.level.3d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}], 0f00000000;"
            : "=f"(g0.x), "=f"(g0.y), "=f"(g0.z), "=f"(g0.w) : "l"(A.tex_a), "f"(cxy.x), "f"(cxy.y), "f"(cz));
        asm("tex.level.3d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}], 0f00000000, {0, 0, 1, 0};"
            : "=f"(g1.x), "=f"(g1.y), "=f"(g1.z), "=f"(g1.w) : "l"(A.tex_a), "f"(cxy.x), "f"(cxy.y), "f"(cz));
        return lerp3_quads(g0, g1, fx, fy, fz);
    }
    const int x0 = (int)flx, y0 = (int)fly, z0 = (int)flz;
    // clamp-to-edge on both taps, like the oracle's scalar_at()
    const int xa2 = min(max(x0, 0), A.nx - 1), xb2 = min(max(x0 + 1, 0), A.nx - 1);
    const int ya2 = min(max(y0, 0), A.ny - 1), yb2 = min(max(y0 + 1, 0), A.ny - 1);
    const int za2 = min(max(z0, 0), A.nz - 1), zb2 = min(max(z0 + 1, 0), A.nz - 1);
    const size_t sy = (size_t)A.nx, sz = (size_t)A.nx * A.ny;
    const size_t r00 = za2 * sz + ya2 * sy, r10 = za2 * sz + yb2 * sy, r01 = zb2 * sz + ya2 * sy, r11 = zb2 * sz + yb2 * sy;
    using S = Scalar<DTYPE>;
    const float c000 = S::load(A.vol_a, r00 + xa2), c100 = S::load(A.vol_a, r00 + xb2);
    const float c010 = S::load(A.vol_a, r10 + xa2), c110 = S::load(A.vol_a, r10 + xb2);
    const float c001 = S::load(A.vol_a, r01 + xa2), c101 = S::load(A.vol_a, r01 + xb2);
    const float c011 = S::load(A.vol_a, r11 + xa2), c111 = S::load(A.vol_a, r11 + xb2);
    return __fmul_rn(lerp3(c000, c100, c010, c110, c001, c101, c011, c111, fx, fy, fz), S::kScale);
}

// Per-ray constants of the leap length model: reciprocal of the voxel-space advance per step on each axis (sign =
// direction of travel; 1e30 for an axis the ray does not move along) and 1 - the drift allowance. (The direction as +-1
// used to sit in three more registers; leap_count now takes the sign bit of rq: the kernel runs at its 40-register cap,
// and the three registers ended the spills: +1 %.)
struct LeapRay {
    float rqx, rqy, rqz, keep;
};

// How many consecutive samples, this one included, provably stay inside the empty region around the sample's
// brick (d >= 1 from the distance field). Region = bricks [c-(d-1), c+d) per axis, shrunk by A.leap_eps voxels
// (covers the rounding of p and q). s = steps until the ray leaves it (approximate line model); the margin
// covers the drift of the replayed additions: each rounds by <= ulp(t)/2, i.e. <= drift_per_step of a step,
// so after n <= s steps the model is off by at most s * drift_per_step steps (+ a fixed 0.02). The result is
// >= 1: the current sample's emptiness was read from its exact index.
//
// Written around the brick centre 8b + 4 (b = brick coordinate as a float, already computed for the distance lookup): the exit
// plane on an axis is centre + sg * R with R = 8d - 4 - eps, so the distance to it is w = 8b + (sg*R + (4 - q)) and
// s = w * rq. (Not u*rq + R*|rq|: for a ray almost parallel to a face rq is huge and that sum cancels.) x and y go
// through the packed FADD2 / FFMA2 / FMUL2: 11 floating-point instructions for the three axes. In M1 the region is
// also clipped to the grid when a partial last brick sticks out of it (A.leap_clip; clamp-to-edge sampling: outside is
// NOT empty; the distance field's border is "occupied", so whole bricks never stick out).
template <int MODE, bool CLIP>
__device__ __forceinline__ int leap_count(const RenderArgs& A, const LeapRay& L, uint32_t d, float2 bxy, float bz, float2 qxy, float qz) {
    const float R = fmaf(A.brick, (float)d, A.leap_r0);  // B d - (B / 2 + eps)
    // sg * R = R with the sign of rq (R > 0): one LOP3 per axis instead of a register per axis for sg; RN(sg*R + c) = RN(+-R + c)
    const uint32_t Rb = __float_as_uint(R);
    const float Rx = __uint_as_float(Rb | (__float_as_uint(L.rqx) & 0x80000000u)), Ry = __uint_as_float(Rb | (__float_as_uint(L.rqy) & 0x80000000u));
    const float Rz = __uint_as_float(Rb | (__float_as_uint(L.rqz) & 0x80000000u));
    const float2 hxy = add2(make_float2(Rx, Ry), add2(make_float2(-qxy.x, -qxy.y), dup2(A.half_brick)));
    const float hz = __fadd_rn(Rz, __fsub_rn(A.half_brick, qz));
    const float2 wxy = fma2(dup2(A.brick), bxy, hxy);
    const float wz = fmaf(A.brick, bz, hz);
    const float2 sxy = mul2(wxy, make_float2(L.rqx, L.rqy));
    float sx = sxy.x, sy = sxy.y, sz = __fmul_rn(wz, L.rqz);
    if (MODE == VKRT_MODE_M1 && CLIP && A.leap_clip) {  // uniform; only grids whose dims are not multiples of the brick edge
        if (L.rqx > 0.0f) sx = fminf(sx, (A.leap_lim[0] - qxy.x) * L.rqx);
        if (L.rqy > 0.0f) sy = fminf(sy, (A.leap_lim[1] - qxy.y) * L.rqy);
        if (L.rqz > 0.0f) sz = fminf(sz, (A.leap_lim[2] - qz) * L.rqz);
    }
    const float sm = fminf(fminf(sx, sy), fminf(sz, (float)(kLeapFastMax - 1)));
    // samples j = 0 .. floor(s - margin) (this one is j = 0) lie inside the region: floor(s - (s*drift + 0.02)) + 1
    return max(__float2int_rz(fmaf(sm, L.keep, 0.98f)), 1);
}

// CLIP = false: the host vouches for A.leap_clip == 0 (launch3), and the leap drops the uniform test of it (three issue
// slots per leap: LDCU + UISETP + BRA.U)
template <int MODE, int LAYOUT, int DTYPE, bool SKIP, bool DBG, bool CLIP>
#ifndef VKRT_M1_BLOCKS
#define VKRT_M1_BLOCKS 12
#endif
__global__ void __launch_bounds__(128, MODE == VKRT_MODE_M0 ? 10 : (LAYOUT != VKRT_LAYOUT_LINEAR ? VKRT_M1_BLOCKS : 8)) raycast_kernel(const __grid_constant__ RenderArgs A) {
    // ---- which pixel -------------------------------------------------------------------------
    // `single`: grid = (block columns, frames of the batch, block rows) — the frame index varies FASTER than the block row, so
    // the launch works through all its frames top to bottom together and the last blocks it schedules are the (culled, cheap)
    // bottom rows of every frame; with the frame slowest the launch ended on the last frame's longest rays (launch_raycast).
    // `tile`: grid = (block columns, block rows, frame * n_tiles + tile).
    const uint32_t gx = blockIdx.x * blockDim.x + threadIdx.x, gy = (A.n_tiles > 0 ? blockIdx.y : blockIdx.z) * blockDim.y + threadIdx.y;
    float offx = 0.0f, offy = 0.0f;
    uint32_t px = gx, py = gy;
    bool valid = true;
    const uint32_t fr = A.n_tiles > 0 ? blockIdx.z / (uint32_t)A.n_tiles : blockIdx.y;
    if (A.n_tiles > 0) {  // `tile` entry: coord = gid + offset; store at gid + u32(offset)
        const VkrtOffset o = A.offsets[blockIdx.z - fr * (uint32_t)A.n_tiles];
        offx = o.x;
        offy = o.y;
        px = gx + __float2uint_rz(offx);  // vec2<u32>(f32) saturates: a negative offset stores at gid + 0
        py = gy + __float2uint_rz(offy);
        valid = gx < (uint32_t)A.tile_size && gy < (uint32_t)A.tile_size;
    }
    valid = valid && px < (uint32_t)A.W && py < (uint32_t)A.H;  // out-of-range textureStore is dropped

    // Four fifths of a frame's blocks lie entirely outside the cull rectangle: they store the clear colour (packed on the
    // host, A.clear_texel) and leave — no per-pixel cull test, no colour bookkeeping (`single` entry; block-uniform branch).
    if (!DBG && A.n_tiles == 0) {
        const float* cull = A.cull[fr];
        const float bx0 = (float)(blockIdx.x * blockDim.x), by0 = (float)(blockIdx.z * blockDim.y);
        if (bx0 > cull[2] || bx0 + (float)(blockDim.x - 1) < cull[0] || by0 > cull[3] || by0 + (float)(blockDim.y - 1) < cull[1]) {
            if (valid) {
                const size_t o = ((size_t)fr * A.H + py) * A.W + px;
                A.frame[o] = A.clear_texel;
                if (A.rgba8) A.rgba8[o] = present_pixel(A.clear_texel);
            }
            return;
        }
    }

    // ---- ray ---------------------------------------------------------------------------------
    // Pixels outside A.cull (the screen rectangle around the projected box, computed on the host with a
    // margin of pixels; the whole plane when the projection is not trustworthy) cannot hit: no ray is built.
    f3 eye = {0.f, 0.f, 0.f}, dir = {0.f, 0.f, 1.f}, inv_dir = {0.f, 0.f, 1.f};
    float t0 = 0.0f, t1 = -1.0f;
    {
        const float cx = (float)gx + offx, cy = (float)gy + offy;
        const float* cull = A.cull[fr];
        bool may_hit = valid && cx >= cull[0] && cy >= cull[1] && cx <= cull[2] && cy <= cull[3];
        if (may_hit) {  // inside the rectangle: inside the silhouette's convex hull as well? (half of the rectangle is not)
            const float(*hp)[3] = A.hull[fr];
            const float d0 = fmaf(hp[0][0], cx, fmaf(hp[0][1], cy, hp[0][2])), d1 = fmaf(hp[1][0], cx, fmaf(hp[1][1], cy, hp[1][2]));
            const float d2 = fmaf(hp[2][0], cx, fmaf(hp[2][1], cy, hp[2][2])), d3 = fmaf(hp[3][0], cx, fmaf(hp[3][1], cy, hp[3][2]));
            const float d4 = fmaf(hp[4][0], cx, fmaf(hp[4][1], cy, hp[4][2])), d5 = fmaf(hp[5][0], cx, fmaf(hp[5][1], cy, hp[5][2]));
            may_hit = fminf(fminf(fminf(d0, d1), fminf(d2, d3)), fminf(d4, d5)) >= 0.0f;
        }
        if (may_hit) {
            gen_ray(A.inv[fr], (float)gx, (float)gy, offx, offy, (float)A.W, (float)A.H, A.aspect_hw, eye, dir);
            intersect_box(eye, dir, t0, t1, inv_dir);
        }
    }
    const bool hit = valid && (t0 < t1);
    t0 = fmaxf(t0, 0.0f);

    Rgba col;
    col.r = MODE == VKRT_MODE_M0 ? A.clear[0] : 0.0f;
    col.g = MODE == VKRT_MODE_M0 ? A.clear[1] : 0.0f;
    col.b = MODE == VKRT_MODE_M0 ? A.clear[2] : 0.0f;
    col.a = A.initial_alpha;
    uint32_t iters = 0, fetched = 0;

    if (hit) {
        const float dt = step_dt(dir, A.fx, A.fy, A.fz, A.dt_scale, A.dt_floor);
        // Leap geometry (SKIP only; approximate on purpose, see leap_count).
        LeapRay L = {};
        LeapCache lc = {0xffffffffu, 0u};
        float drift_per_step = 0.0f;
        if (SKIP) {
            const float dqx = dir.x * A.hx * dt, dqy = dir.y * A.hy * dt, dqz = dir.z * A.hz * dt;  // voxels per step
            // (approximate reciprocals, MUFU.RCP: the leap length model carries a margin of 0.02 steps + the drift, a relative
            // error of 2^-22 in rq moves a 4095-step leap by 0.001)
            L.rqx = fabsf(dqx) > 1e-12f ? __fdividef(1.0f, dqx) : 1e30f;
            L.rqy = fabsf(dqy) > 1e-12f ? __fdividef(1.0f, dqy) : 1e30f;
            L.rqz = fabsf(dqz) > 1e-12f ? __fdividef(1.0f, dqz) : 1e30f;
            // one replayed addition moves t off the exact line by <= ulp(t)/2 <= t1 * 2^-24; in units of a step:
            drift_per_step = __fdividef(t1 * 5.9604645e-08f, dt) * 2.0f;  // x2 safety
            L.keep = 1.0f - drift_per_step;
        }
        // The ray's octant selects its directional distance field (RenderArgs::dist); the signs are the ones the leap
        // length model uses. M1 folds the table into the z brick coordinate (kMagic + octant * slabs per table + bz).
        const uint32_t oct = SKIP ? ((L.rqx > 0.0f ? 1u : 0u) | (L.rqy > 0.0f ? 2u : 0u) | (L.rqz > 0.0f ? 4u : 0u)) : 0u;
        const float magic_z = fmaf((float)oct, A.dsz_f, kMagic);
        const uint32_t oct_off = oct * A.dist_tab;
        // (Two traversal restructurings were measured and rejected on B200. While-while — every lane first
        // advances to its next non-empty sample on its own, then the warp shades together: 1.9x slower,
        // profiles/r01_whilewhile_ab.md. Warp-uniform leaps — leap only when every live lane sits in empty
        // space, all by the warp minimum: 13 % slower, the 20 % extra samples taken by lanes that could have
        // leapt outweigh the leaps saved, profiles/r01_uniform_ab.md.)
        float t = t0, t_end = t1;
        bool terminated = false;  // early ray termination: alpha reached the threshold
        if (SKIP) {
            // Clip the march to the bounding box of the occupied bricks (grown by one voxel, A.bb_*): every
            // sample outside it lies in an empty brick, i.e. is a bit-exact no-op, so the loop may stop at the
            // box's exit (no trailing leaps to the far face), does not run at all for rays that miss it, and
            // reaches its entry with ONE leap — landing on the floats the reference's t = t + dt visits, like
            // every leap. The exit test is on the sample's own t, so drift does not matter there; the entry
            // leap keeps the margin of leap_count (2 steps + the drift of the replayed additions), on top of
            // the voxel of margin in the box.
            float tb0, tb1;
            slab_box(eye, inv_dir, A.bb_lo, A.bb_hi, tb0, tb1);
            // a ray that misses the box (or an empty volume) does not march at all; otherwise tb0 < t_end <= t1,
            // which also bounds the entry leap (a ray almost parallel to a face it passes outside of has an
            // "entry" thousands of steps away)
            t_end = (tb0 < tb1 && A.bb_lo[0] <= A.bb_hi[0]) ? fminf(t1, tb1) : -1.0f;
            if (tb0 > t0 && tb0 < t_end) {
                const float s = fminf((tb0 - t0) * __frcp_rn(dt), 1.0e6f);
                const int n0 = __float2int_rz(s - fmaf(s, drift_per_step, 2.0f));
                if (n0 >= 1) {
                    if (DBG) {
                        for (int j = 0; j < n0 && t < t1; ++j) {
                            ++iters;
                            t = xadd(t, dt);
                        }
                    } else {
                        t = leap_steps_slow(t, dt, n0, lc);  // up to 10^6 steps: the 64-bit form
                    }
                }
            }
        }
        if (SKIP && !DBG && lc.eb == 0xffffffffu && t < t_end) {
            // the increment of the binade the march starts in, computed here at full lane occupancy rather than by the
            // first leap's slow path (a ray without an entry leap has nothing cached yet)
            const uint32_t e = __float_as_uint(t) >> 23, inc = binade_inc(e, dt);
            if (inc < (1u << 19)) {
                lc.eb = e << 23;
                lc.inc = inc;
            }
        }
        const float2 exy = make_float2(eye.x, eye.y), dxy = make_float2(dir.x, dir.y), hxy = make_float2(A.hx, A.hy), one2 = dup2(A.one);
        // Early ray termination is part of the loop condition, not a `break` out of the divergent sample branch: a `break`
        // costs a reliable-reconvergence pair (BSSY.RELIABLE / BSYNC) in EVERY iteration plus BREAK + BRA in the sample path;
        // the flag lives in a predicate and rides on the loop's own branch (`@P0 BRA P2`).
        while (!terminated && t < t_end) {
            // p = eye + t*dir ; q = (p + 1) * (N/2) — exact (IEEE round-to-nearest per component, the oracle's operation
            // order), decides the texel; x and y go through the packed FMUL2 / FADD2
            const float2 pxy = fma2(mul2(dup2(t), dxy), one2, exy);  // RN(m * 1 + e) = RN(m + e), see A.one
            const float pz = xadd(eye.z, xmul(t, dir.z));
            const float2 qxy = mul2(add2(pxy, dup2(1.0f)), hxy);
            const float qx = qxy.x, qy = qxy.y, qz = xmul(xadd(pz, 1.0f), A.hz);
            const int ix = __float2int_rz(qx), iy = __float2int_rz(qy), iz = __float2int_rz(qz);
            // M1 needs no bounds test: the fetch clamps to the edge, and the distance field it consults is padded (below)
            const bool inb = MODE != VKRT_MODE_M0 || ((unsigned)ix < (unsigned)A.nx && (unsigned)iy < (unsigned)A.ny && (unsigned)iz < (unsigned)A.nz);
            if (SKIP) {
                // Exact empty-space skipping. A sample in an empty brick (or, in M0,
                // outside the grid) leaves colour and alpha bit-identical, so only the t sequence
                // advances — landing on exactly the floats the reference's `t = t + dt` visits. n = how
                // many consecutive samples (this one included) provably stay inside the empty region.
                int n = 0;
                if (MODE == VKRT_MODE_M1) {
                    // brick coordinates floor(q / 8) as magic floats (FFMA.RM): their bit patterns index the padded tables
                    // directly. q == N (p rounded onto the box face) lands on the pad layer, a slightly negative q on the
                    // pad of the previous row / slab / table or past the end (clamped): all "occupied", i.e. the sample is
                    // evaluated. (floor, not the truncation of the voxel index: differs for -1 < q < 0 only, towards the pad)
                    const float2 bxy = fma2_rm(qxy, dup2(A.inv_brick), dup2(kMagic));
                    const float bzm = __fmaf_rd(qz, A.inv_brick, magic_z);
                    uint32_t cell = (__float_as_uint(bzm) * A.dsy + __float_as_uint(bxy.y)) * A.dsx + __float_as_uint(bxy.x) - A.dist_bias;
                    cell = min(cell, A.dist_last);
                    const uint32_t d = __ldg(A.dist + cell);
                    if (d != 0u) n = leap_count<MODE, CLIP>(A, L, d, add2(bxy, dup2(-kMagic)), __fsub_rn(bzm, magic_z), qxy, qz);
                } else if (!inb) {
                    n = 1;
                } else {
                    const int bx = ix >> A.obs, by = iy >> A.obs, bz = iz >> A.obs;
                    const uint32_t cell = ((uint32_t)bz * A.dsy + (uint32_t)by) * A.dsx + (uint32_t)bx + oct_off;
                    const uint32_t d = __ldg(A.dist + cell);
                    if (d != 0u) n = leap_count<MODE, CLIP>(A, L, d, make_float2((float)bx, (float)by), (float)bz, qxy, qz);
                }
                if (n > 0) {
                    if (DBG) {
                        for (int j = 0; j < n && t < t1; ++j) {
                            ++iters;
                            t = xadd(t, dt);
                        }
                    } else {
                        t = leap_steps(t, dt, n, lc);  // closed form, bit-identical to n additions
                    }
                    continue;
                }
            }
            if (DBG) {
                ++iters;
                ++fetched;
            }
            if (MODE == VKRT_MODE_M0) {
                float4 c, n;
                m0_fetch<LAYOUT>(A, ix, iy, iz, inb, c, n);
                const f3 p = {pxy.x, pxy.y, pz};
                m0_shade(col, c, n, p, A.clear);
            } else {
                m1_shade_acc(col, m1_sample<LAYOUT, DTYPE>(A, qxy, qz));  // the palette's constant half is added after the loop
            }
            if (DBG) {  // the counting kernel keeps t at the terminating sample (the tail below counts from it)
                if (col.a >= A.alpha_threshold) {
                    terminated = true;
                    break;
                }
                t = xadd(t, dt);
            } else {
                t = xadd(t, dt);
                terminated = col.a >= A.alpha_threshold;
            }
        }
        if (DBG && SKIP && !terminated) {  // the reference's loop runs on to the far face: count those iterations
            while (t < t1) {
                ++iters;
                t = xadd(t, dt);
            }
        }
    }
    if (hit) {
        if (MODE == VKRT_MODE_M1) m1_finish(col, A.initial_alpha);
        if (MODE == VKRT_MODE_M1 && A.m1_srgb) {
            col.r = linear_to_srgb_naive(col.r);
            col.g = linear_to_srgb_naive(col.g);
            col.b = linear_to_srgb_naive(col.b);
        }
    } else {
        col.r = A.clear[0];
        col.g = A.clear[1];
        col.b = A.clear[2];
    }

    if (valid) {
        const size_t o = ((size_t)fr * A.H + py) * A.W + px;
        const uint2 texel = pack_rgba16f(col.r, col.g, col.b, 1.0f);
        A.frame[o] = texel;
        if (A.rgba8) A.rgba8[o] = present_pixel(texel);  // fused present pass (host-frame paths): no second kernel, no re-read
        if (DBG && A.aux) A.aux[o] = hit ? (0x80000000u | iters) : 0u;
    }
    if (DBG && A.counters) {
        const unsigned h = __reduce_add_sync(0xffffffffu, hit ? 1u : 0u);
        const unsigned it = __reduce_add_sync(0xffffffffu, hit ? iters : 0u);
        const unsigned fe = __reduce_add_sync(0xffffffffu, hit ? fetched : 0u);
        if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0) {
            atomicAdd(A.counters + 0, (unsigned long long)h);
            atomicAdd(A.counters + 1, (unsigned long long)it);
            atomicAdd(A.counters + 2, (unsigned long long)fe);
        }
    }
}

template <int MODE, int LAYOUT, int DTYPE>
cudaError_t launch3(const RenderArgs& A, dim3 grid, dim3 block, cudaStream_t s, bool skip, bool dbg) {
    if (skip) {
        if (dbg) raycast_kernel<MODE, LAYOUT, DTYPE, true, true, true><<<grid, block, 0, s>>>(A);
        // M1 on a grid whose dims are multiples of the occupancy brick: the instantiation without the clip test (M0 never
        // clips: for it this names the same instantiation as the line below)
        else if (MODE == VKRT_MODE_M1 && !A.leap_clip) raycast_kernel<MODE, LAYOUT, DTYPE, true, false, MODE != VKRT_MODE_M1><<<grid, block, 0, s>>>(A);
        else raycast_kernel<MODE, LAYOUT, DTYPE, true, false, true><<<grid, block, 0, s>>>(A);
    } else {
        if (dbg) raycast_kernel<MODE, LAYOUT, DTYPE, false, true, true><<<grid, block, 0, s>>>(A);
        else raycast_kernel<MODE, LAYOUT, DTYPE, false, false, true><<<grid, block, 0, s>>>(A);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_raycast(const RenderArgs& A, int mode, int layout, int dtype, bool skip, bool dbg, cudaStream_t s) {
    // Block = 8x16 pixels: blockDim.x = 8 makes every warp an 8x4-pixel tile; measured best on B200
    // (profiles/r01_blockshape.md).
    const int bw = 8, bh = 16;
    const dim3 block((unsigned)bw, (unsigned)bh, 1);
    dim3 grid;
    if (A.n_tiles > 0) {
        grid = dim3((unsigned)((A.tile_size + bw - 1) / bw), (unsigned)((A.tile_size + bh - 1) / bh),
                    (unsigned)A.n_tiles * (unsigned)(A.n_frames > 0 ? A.n_frames : 1));
    } else {
        grid = dim3((unsigned)((A.W + bw - 1) / bw), (unsigned)(A.n_frames > 0 ? A.n_frames : 1), (unsigned)((A.H + bh - 1) / bh));
    }
    if (mode == VKRT_MODE_M0) {
        switch (layout) {
            case VKRT_LAYOUT_LINEAR: return launch3<VKRT_MODE_M0, VKRT_LAYOUT_LINEAR, 0>(A, grid, block, s, skip, dbg);
            case VKRT_LAYOUT_BRICKED: return launch3<VKRT_MODE_M0, VKRT_LAYOUT_BRICKED, 0>(A, grid, block, s, skip, dbg);
            case VKRT_LAYOUT_TEXTURE: return launch3<VKRT_MODE_M0, VKRT_LAYOUT_TEXTURE, 0>(A, grid, block, s, skip, dbg);
        }
    } else {
        if (layout == VKRT_LAYOUT_TEXTURE) return launch3<VKRT_MODE_M1, VKRT_LAYOUT_TEXTURE, 0>(A, grid, block, s, skip, dbg);
        if (layout == VKRT_LAYOUT_GATHER) return launch3<VKRT_MODE_M1, VKRT_LAYOUT_GATHER, 0>(A, grid, block, s, skip, dbg);
        if (layout == VKRT_LAYOUT_QUAD) return launch3<VKRT_MODE_M1, VKRT_LAYOUT_QUAD, 0>(A, grid, block, s, skip, dbg);
        if (layout == VKRT_LAYOUT_LINEAR) {
            switch (dtype) {
                case VKRT_U8: return launch3<VKRT_MODE_M1, VKRT_LAYOUT_LINEAR, VKRT_U8>(A, grid, block, s, skip, dbg);
                case VKRT_F16: return launch3<VKRT_MODE_M1, VKRT_LAYOUT_LINEAR, VKRT_F16>(A, grid, block, s, skip, dbg);
                case VKRT_F32: return launch3<VKRT_MODE_M1, VKRT_LAYOUT_LINEAR, VKRT_F32>(A, grid, block, s, skip, dbg);
            }
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace vkrt
