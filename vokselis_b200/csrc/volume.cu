// volume.cu — volume staging kernels: bricked re-layout, occupancy grids, and the procedural generator.
//
//  * interleave_bricked: builds the B200 layout of DESIGN.md §4.1 from the upload layout of
//    examples/xor/xor_compute.rs:93-118 (two rgba16f textures, x fastest).
//  * occupancy_m0 / occupancy_m1 + distance_step: one byte per 8^3-voxel brick — 0 when some sample falling
//    in the brick can change the running colour, else the Chebyshev distance (in bricks) to the nearest such
//    brick (exactness argument: DESIGN.md §4.2). The M1 bound 0.0999999 sits just below the transfer
//    function's 0.1f threshold so that fp32 rounding inside any trilinear formula cannot cross it.
//  * generate_xor: shaders/xor.wgsl `cs_main` (:69-78) with `noise_volume` (:55-61) or the dead
//    bit-pattern `volume` (:46-53), writing both rgba16f volumes on the device ("next" row N1).
#include <algorithm>

#include "raycast.cuh"
#include "vkrt_device.cuh"

namespace vkrt {

namespace {

__global__ void __launch_bounds__(256) interleave_bricked_kernel(const uint2* __restrict__ color, const uint2* __restrict__ normal,
                                                                 uint4* __restrict__ out, int nx, int ny, int nz, int nbx, int nby,
                                                                 size_t total) {
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    // invert bricked_index(): o = brick<<9 | line<<3 | in_line
    const uint32_t in_line = (uint32_t)o & 7u, line = ((uint32_t)o >> 3) & 63u;
    const size_t brick = o >> 9;
    const int bx = (int)(brick % (size_t)nbx), by = (int)((brick / (size_t)nbx) % (size_t)nby), bz = (int)(brick / ((size_t)nbx * nby));
    const int ix = bx * 8 + (int)((line & 3u) << 1 | (in_line & 1u));
    const int iy = by * 8 + (int)(((line >> 2) & 3u) << 1 | ((in_line >> 1) & 1u));
    const int iz = bz * 8 + (int)(((line >> 4) & 3u) << 1 | ((in_line >> 2) & 1u));
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ix < nx && iy < ny && iz < nz) {
        const size_t i = ((size_t)iz * ny + iy) * nx + ix;
        const uint2 c = __ldg(color + i), n = __ldg(normal + i);
        v = make_uint4(c.x, c.y, n.x, n.y);
    }
    out[o] = v;
}

// one warp per brick (edge 2^bs voxels)
__global__ void __launch_bounds__(256) occupancy_m0_kernel(const uint2* __restrict__ color, const uint2* __restrict__ normal, int nx, int ny, int nz, int nbx, int nby,
                                                           int nbz, int bs, uint8_t* __restrict__ dist) {
    const size_t cell = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31u;
    if (cell >= (size_t)nbx * nby * nbz) return;
    const int bx = (int)(cell % (size_t)nbx), by = (int)((cell / (size_t)nbx) % (size_t)nby), bz = (int)(cell / ((size_t)nbx * nby));
    const int m = (1 << bs) - 1;
    bool any = false;
    for (int k = (int)lane; k < (1 << (3 * bs)); k += 32) {
        const int ix = (bx << bs) + (k & m), iy = (by << bs) + ((k >> bs) & m), iz = (bz << bs) + (k >> (2 * bs));
        if (ix < nx && iy < ny && iz < nz) {
            const size_t i = ((size_t)iz * ny + iy) * nx + ix;
            any = any || !m0_texel_skippable(__ldg(color + i), __ldg(normal + i));  // alpha through the very function the march uses
        }
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) dist[cell] = any ? 0 : 255;
}

template <int DTYPE> __device__ __forceinline__ float load_scalar(const void* p, size_t i);
template <> __device__ __forceinline__ float load_scalar<VKRT_U8>(const void* p, size_t i) { return (float)__ldg((const uint8_t*)p + i) / 255.0f; }
template <> __device__ __forceinline__ float load_scalar<VKRT_F16>(const void* p, size_t i) {
    return __half2float(__ushort_as_half(__ldg((const unsigned short*)p + i)));
}
template <> __device__ __forceinline__ float load_scalar<VKRT_F32>(const void* p, size_t i) { return __ldg((const float*)p + i); }

// A sample whose index floor((p+1)*N/2) lies in brick b (edge B = 2^bs voxels) touches voxels [B b - 1, B b + B] per axis
// (clamped to the grid). The brick is empty iff all of them are <= M1_EMPTY_MAX.
#define M1_EMPTY_MAX 0.0999999f
template <int DTYPE>
__global__ void __launch_bounds__(256) occupancy_m1_kernel(const void* __restrict__ vol, int nx, int ny, int nz, int nbx, int nby, int nbz, int bs,
                                                           uint8_t* __restrict__ dist) {
    const size_t cell = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31u;
    if (cell >= (size_t)nbx * nby * nbz) return;
    const int bx = (int)(cell % (size_t)nbx), by = (int)((cell / (size_t)nbx) % (size_t)nby), bz = (int)(cell / ((size_t)nbx * nby));
    const int B = 1 << bs;
    const int x0 = max(bx * B - 1, 0), x1 = min(bx * B + B, nx - 1);
    const int y0 = max(by * B - 1, 0), y1 = min(by * B + B, ny - 1);
    const int z0 = max(bz * B - 1, 0), z1 = min(bz * B + B, nz - 1);
    const int wy = y1 - y0 + 1, wz = z1 - z0 + 1;
    bool any = false;
    for (int r = (int)lane; r < wy * wz; r += 32) {
        const int iy = y0 + r % wy, iz = z0 + r / wy;
        const size_t row = ((size_t)iz * ny + iy) * nx;
        for (int ix = x0; ix <= x1; ++ix) {
            const float v = load_scalar<DTYPE>(vol, row + ix);
            any = any || !(v <= M1_EMPTY_MAX);  // NaN counts as occupied
        }
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) dist[cell] = any ? 0 : 255;
}

// One relaxation step of the Chebyshev distance transform over bricks: d = min(d, 1 + min over the 26
// neighbours). `border` is the value of bricks outside the grid: 255 (empty, M0: out-of-range texels
// are zero) or 0 (occupied, M1: clamp-to-edge sampling reads edge voxels).
__global__ void __launch_bounds__(256) distance_step_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int nbx, int nby,
                                                            int nbz, int border, int cap) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (size_t)nbx * nby * nbz) return;
    const int bx = (int)(c % (size_t)nbx), by = (int)((c / (size_t)nbx) % (size_t)nby), bz = (int)(c / ((size_t)nbx * nby));
    int d = in[c];
    if (d != 0) {
        int m = 255;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int x = bx + dx, y = by + dy, z = bz + dz;
                    const int v = (x < 0 || y < 0 || z < 0 || x >= nbx || y >= nby || z >= nbz) ? border : (int)in[((size_t)z * nby + y) * nbx + x];
                    m = min(m, v);
                }
        d = min(min(d, m + 1), cap);
    }
    out[c] = (uint8_t)d;
}

// Directional distance fields, one per ray octant k (bit 0 / 1 / 2 set = the ray travels towards +x / +y / +z): table k holds,
// per brick b, the largest d such that the d^3 bricks [b, b + s d) (s = the octant's sign on each axis) are all empty; 0 =
// occupied. A ray only ever moves forward, so the cube BEHIND a sample (which the isotropic Chebyshev distance also
// requires to be empty) does not limit its leap: on the bench volume the mean distance over empty bricks grows from 1.6 to
// 2.7 bricks and a ray needs half as many leaps (bench/leap_model.py). The forward extent of the region is the same
// 8 d - 4 voxels from the brick centre as before, so the kernel's leap length model is unchanged.
// One relaxation step: d = min(d, 1 + min over the 7 forward neighbours); tables are initialised with the occupancy
// (0 / 255). `in` / `out` hold the 8 tables back to back.
__global__ void __launch_bounds__(256) octant_step_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int nbx, int nby,
                                                          int nbz, int border, int cap, int* __restrict__ changed) {
    const size_t cells = (size_t)nbx * nby * nbz;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= 8 * cells) return;
    const int k = (int)(g / cells);
    const size_t c = g - (size_t)k * cells;
    const uint8_t* tab = in + (size_t)k * cells;
    int d = tab[c];
    if (d != 0) {
        const int bx = (int)(c % (size_t)nbx), by = (int)((c / (size_t)nbx) % (size_t)nby), bz = (int)(c / ((size_t)nbx * nby));
        const int sx = (k & 1) ? 1 : -1, sy = (k & 2) ? 1 : -1, sz = (k & 4) ? 1 : -1;
        int m = 255;
        for (int j = 1; j < 8; ++j) {
            const int x = bx + ((j & 1) ? sx : 0), y = by + ((j & 2) ? sy : 0), z = bz + ((j & 4) ? sz : 0);
            const int v = (x < 0 || y < 0 || z < 0 || x >= nbx || y >= nby || z >= nbz) ? border : (int)tab[((size_t)z * nby + y) * nbx + x];
            m = min(m, v);
        }
        const int nd = min(min(d, m + 1), cap);
        if (nd != d && changed) *changed = 1;  // (benign race: every writer stores 1)
        d = nd;
    }
    out[g] = (uint8_t)d;
}

__global__ void __launch_bounds__(256) replicate8_kernel(const uint8_t* __restrict__ occ, uint8_t* __restrict__ out, size_t cells) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < 8 * cells) out[g] = occ[g % cells];
}

// Bounding box, in bricks, of the occupied cells (dist == 0) of the finished distance field:
// out = {min x, min y, min z, max x, max y, max z, number of EMPTY cells (saturating at INT_MAX)}, initialised to
// {INT_MAX x3, -1 x3, 0} by the launcher.
__global__ void __launch_bounds__(256) occupied_bounds_kernel(const uint8_t* __restrict__ dist, int nbx, int nby, int nbz, int* __restrict__ out) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = c < (size_t)nbx * nby * nbz;
    const bool occ = in && dist[c] == 0;
    const int n_empty = __popc(__ballot_sync(0xffffffffu, in && !occ));
    if ((threadIdx.x & 31u) == 0u && n_empty) {
        const int old = atomicAdd(out + 6, n_empty);
        if (old < 0 || old + n_empty < 0) atomicExch(out + 6, 0x7fffffff);
    }
    const int bx = (int)(c % (size_t)nbx), by = (int)((c / (size_t)nbx) % (size_t)nby), bz = (int)(c / ((size_t)nbx * nby));
    const int big = 0x7fffffff;
    const int lx = __reduce_min_sync(0xffffffffu, occ ? bx : big), ly = __reduce_min_sync(0xffffffffu, occ ? by : big);
    const int lz = __reduce_min_sync(0xffffffffu, occ ? bz : big);
    const int hx = __reduce_max_sync(0xffffffffu, occ ? bx : -1), hy = __reduce_max_sync(0xffffffffu, occ ? by : -1);
    const int hz = __reduce_max_sync(0xffffffffu, occ ? bz : -1);
    if ((threadIdx.x & 31u) == 0u && hx >= 0) {
        atomicMin(out + 0, lx); atomicMin(out + 1, ly); atomicMin(out + 2, lz);
        atomicMax(out + 3, hx); atomicMax(out + 4, hy); atomicMax(out + 5, hz);
    }
}

// ---- shaders/xor.wgsl -------------------------------------------------------------------------
__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }
// :18-20. The product is rounded before the floor is subtracted (no FMA contraction): the hash amplifies
// any rounding difference by 43758, so only sinf's own last-ulp differences from libm remain.
__device__ __forceinline__ float hash1(float h) {
    const float x = __fmul_rn(sinf(h), 43758.5453123f);
    return __fsub_rn(x, floorf(x));
}
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ float noise3(float x, float y, float z) {  // :22-33
    const float px = floorf(x), py = floorf(y), pz = floorf(z);
    float fx = fractf(x), fy = fractf(y), fz = fractf(z);
    fx = fx * fx * (3.0f - 2.0f * fx);
    fy = fy * fy * (3.0f - 2.0f * fy);
    fz = fz * fz * (3.0f - 2.0f * fz);
    const float n = px + py * 157.0f + 113.0f * pz;
    return mixf(mixf(mixf(hash1(n + 0.0f), hash1(n + 1.0f), fx), mixf(hash1(n + 157.0f), hash1(n + 158.0f), fx), fy),
                mixf(mixf(hash1(n + 113.0f), hash1(n + 114.0f), fx), mixf(hash1(n + 270.0f), hash1(n + 271.0f), fx), fy), fz);
}
__device__ float fbm(float x, float y, float z) {  // :35-44
    float f = 0.5f * noise3(x, y, z);
    x *= 2.01f; y *= 2.01f; z *= 2.01f;
    f += 0.25f * noise3(x, y, z);
    x *= 2.02f; y *= 2.02f; z *= 2.02f;
    f += 0.125f * noise3(x, y, z);
    return f;
}
__device__ __forceinline__ float smoothstep_gen(float e0, float e1, float x) {
    float t = fminf(fmaxf((x - e0) / (e1 - e0), 0.0f), 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
// returns (val, alpha): :55-61 noise_volume (which = 0), :46-53 volume (which = 1)
__device__ float2 gen_volume(float cx, float cy, float cz, float time, int which) {
    const float posx = (cx + 1.0f) * 32.0f, posy = (cy + sinf(time * 1.0f) * 0.1f) * 32.0f, posz = (cz + 21.0f) * 32.0f;
    const float len = sqrtf(cx * cx + cy * cy + cz * cz);
    if (which == 0) {
        const float val = fbm(posx, posy, posz);
        return make_float2(val, val * smoothstep_gen(0.5f, 0.25f, len));
    }
    const float res = 25.0f;
    const float val = (float)(__float2int_rz(posx * res) & __float2int_rz(posy * res) & __float2int_rz(posz * res)) / res;
    return make_float2(val, val * smoothstep_gen(0.7f, 0.0f, len));
}

__global__ void __launch_bounds__(256) generate_xor_kernel(uint2* __restrict__ color, uint2* __restrict__ normal, int n, float time, int which) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * n * n) return;
    const int x = (int)(i % (size_t)n), y = (int)((i / (size_t)n) % (size_t)n), z = (int)(i / ((size_t)n * n));
    const float dims = (float)n;
    const float cx = ((float)x - dims / 2.0f) / dims, cy = ((float)y - dims / 2.0f) / dims, cz = ((float)z - dims / 2.0f) / dims;
    const float2 vol = gen_volume(cx, cy, cz, time, which);
    // gradient (:63-67), eps = 0.0001
    const float eps = 0.0001f;
    const float a = vol.y;
    const float dx = a - gen_volume(cx - eps, cy, cz, time, which).y;
    const float dy = a - gen_volume(cx, cy - eps, cz, time, which).y;
    const float dz = a - gen_volume(cx, cy, cz - eps, time, which).y;
    const float len = sqrtf(dx * dx + dy * dy + dz * dz);
    const float nx_ = dx / len, ny_ = dy / len, nz_ = dz / len;  // 0/0 = NaN in empty space, like the reference
    const float nl = sqrtf(nx_ * nx_ + ny_ * ny_ + nz_ * nz_);
    color[i] = pack_rgba16f(vol.x / 2.0f, vol.x / 2.0f, vol.x / 2.0f, vol.y);
    normal[i] = pack_rgba16f(nx_, ny_, nz_, nl);
}

// Writes a buffer larger than L2 so the next frame starts with a cold L2 (bench hygiene only).
__global__ void __launch_bounds__(256) flush_l2_kernel(uint4* __restrict__ buf, size_t n16, uint32_t tag) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) buf[i] = make_uint4(tag, tag, tag, tag);
}

// ---- "next" row N3: scalar volume -> the rgba16f pair raycast_compute.wgsl consumes, so that M0's lighting
// works on u8/f16/f32 data (e.g. the bonsai scan through the compute raycaster = BASELINE configs[0]).
// Same construction as shaders/xor.wgsl cs_main (:69-78) on a sampled field: colour = (a/2, a/2, a/2, a)
// with a = the scalar; normal = normalize(a(p) - (a(p-ex), a(p-ey), a(p-ez))) — gradient() (:63-67) with
// the one-voxel backward differences a grid offers, clamped at the border; normal.w = length(normal).
// A zero gradient gives NaN normals, exactly like the reference's empty space.
template <int DTYPE>
__global__ void __launch_bounds__(256) scalar_to_rgba16f_kernel(const void* __restrict__ vol, uint2* __restrict__ color, uint2* __restrict__ normal,
                                                                int nx, int ny, int nz) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nx * ny * nz) return;
    const int x = (int)(i % (size_t)nx), y = (int)((i / (size_t)nx) % (size_t)ny), z = (int)(i / ((size_t)nx * ny));
    const float a = load_scalar<DTYPE>(vol, i);
    const float ax = load_scalar<DTYPE>(vol, i - (x > 0 ? 1 : 0));
    const float ay = load_scalar<DTYPE>(vol, i - (y > 0 ? (size_t)nx : 0));
    const float az = load_scalar<DTYPE>(vol, i - (z > 0 ? (size_t)nx * ny : 0));
    const float dx = __fsub_rn(a, ax), dy = __fsub_rn(a, ay), dz = __fsub_rn(a, az);
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const float n0 = __fdiv_rn(dx, len), n1 = __fdiv_rn(dy, len), n2 = __fdiv_rn(dz, len);
    const float nl = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(n0, n0), __fmul_rn(n1, n1)), __fmul_rn(n2, n2)));
    const float h = __fdiv_rn(a, 2.0f);
    color[i] = pack_rgba16f(h, h, h, a);
    normal[i] = pack_rgba16f(n0, n1, n2, nl);
}

// ---- synthetic scalar volumes of the shapes BASELINE.json names (generated on the device: configs 3-5
// are 2 GiB .. 256 GiB and cannot be staged from host files). Integer hashes + exactly rounded fp32
// operations only, so the same (kind, seed, coordinates) give the same voxel on any GPU and any window.
__device__ __forceinline__ uint32_t hash3(uint32_t x, uint32_t y, uint32_t z, uint32_t seed) {
    uint32_t v = (x * 0x9E3779B1u) ^ (y * 0x85EBCA77u) ^ (z * 0xC2B2AE3Du) ^ (seed * 0x27D4EB2Fu);
    v ^= v >> 15; v *= 0x2C1B3C6Du;
    v ^= v >> 12; v *= 0x297A2D39u;
    v ^= v >> 15;
    return v;
}
__device__ __forceinline__ float hash01(int x, int y, int z, uint32_t seed) {
    return (float)(hash3((uint32_t)x, (uint32_t)y, (uint32_t)z, seed) >> 8) * (1.0f / 16777216.0f);
}
// kind 0: uniform hash noise, box-filtered 3^3 once, mapped to a thin fog 0.08 + 0.14 * noise (config 3):
// most values sit just above the transfer function's 0.1 threshold, so rays traverse the whole volume
// (~N samples per ray) instead of terminating after a few samples
__device__ float synth_noise(int x, int y, int z, int gnx, int gny, int gnz, uint32_t seed) {
    float acc = 0.0f;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx)
                acc = __fadd_rn(acc, hash01(min(max(x + dx, 0), gnx - 1), min(max(y + dy, 0), gny - 1), min(max(z + dz, 0), gnz - 1), seed));
    return __fadd_rn(0.08f, __fmul_rn(0.14f, __fmul_rn(acc, 1.0f / 27.0f)));
}
// kind 1: 90 % of the 64^3 super-bricks exactly 0 (Bernoulli per super-brick, so empties are spatially
// coherent); an occupied super-brick holds a dense ball that fades to 0 before its faces (config 4)
__device__ float synth_sparse(int x, int y, int z, uint32_t seed) {
    const int sx = x >> 6, sy = y >> 6, sz = z >> 6;
    if (hash3((uint32_t)sx, (uint32_t)sy, (uint32_t)sz, seed ^ 0xA511E9B3u) % 10u != 0u) return 0.0f;
    const float fx = __fmul_rn((float)((x & 63) - 32) + 0.5f, 1.0f / 32.0f), fy = __fmul_rn((float)((y & 63) - 32) + 0.5f, 1.0f / 32.0f),
                fz = __fmul_rn((float)((z & 63) - 32) + 0.5f, 1.0f / 32.0f);
    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)), __fmul_rn(fz, fz));
    const float fall = fmaxf(__fsub_rn(1.0f, __fmul_rn(r2, 1.2f)), 0.0f);  // 0 for r >= 0.913: zero on the faces
    const float grain = __fadd_rn(0.75f, __fmul_rn(0.25f, hash01(x >> 2, y >> 2, z >> 2, seed)));
    return fminf(__fmul_rn(__fmul_rn(fall, 1.6f), grain), 1.0f);
}
// kind 2: smooth lattice noise, period 16 voxels, trilinear between hashed lattice points (config 5)
__device__ float synth_smooth(int x, int y, int z, uint32_t seed) {
    const int lx = x >> 4, ly = y >> 4, lz = z >> 4;
    const float fx = (float)(x & 15) * (1.0f / 16.0f), fy = (float)(y & 15) * (1.0f / 16.0f), fz = (float)(z & 15) * (1.0f / 16.0f);
    float c[2][2][2];
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < 2; ++j)
            for (int i = 0; i < 2; ++i) c[k][j][i] = hash01(lx + i, ly + j, lz + k, seed);
    float r[2];
    for (int k = 0; k < 2; ++k) {
        const float a = __fadd_rn(c[k][0][0], __fmul_rn(fx, __fsub_rn(c[k][0][1], c[k][0][0])));
        const float b = __fadd_rn(c[k][1][0], __fmul_rn(fx, __fsub_rn(c[k][1][1], c[k][1][0])));
        r[k] = __fadd_rn(a, __fmul_rn(fy, __fsub_rn(b, a)));
    }
    const float v = __fadd_rn(r[0], __fmul_rn(fz, __fsub_rn(r[1], r[0])));
    return __fmul_rn(__fmul_rn(v, v), 0.9f);  // skew towards low values: part of the field is transparent
}

// bricked = 1: the output is laid out as 8^3-voxel bricks (x fastest inside a brick, bricks x fastest), dims
// padded up to multiples of 8; thread i writes element i of that layout (coalesced), padding voxels get 0.
template <int DTYPE>
__global__ void __launch_bounds__(256) synth_kernel(void* __restrict__ out, int kind, int nx, int ny, int nz, int ox, int oy, int oz,
                                                    int gnx, int gny, int gnz, uint32_t seed, int bricked) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int lx, ly, lz;
    if (bricked) {
        const int bnx = (nx + 7) >> 3, bny = (ny + 7) >> 3, bnz = (nz + 7) >> 3;
        if (i >= (size_t)bnx * bny * bnz * 512) return;
        const size_t b = i >> 9;
        const int in = (int)(i & 511);
        lx = (int)(b % (size_t)bnx) * 8 + (in & 7);
        ly = (int)((b / (size_t)bnx) % (size_t)bny) * 8 + ((in >> 3) & 7);
        lz = (int)(b / ((size_t)bnx * bny)) * 8 + (in >> 6);
        if (lx >= nx || ly >= ny || lz >= nz) {
            if (DTYPE == VKRT_U8) ((uint8_t*)out)[i] = 0;
            else if (DTYPE == VKRT_F16) ((__half*)out)[i] = __float2half_rn(0.0f);
            else ((float*)out)[i] = 0.0f;
            return;
        }
    } else {
        if (i >= (size_t)nx * ny * nz) return;
        lx = (int)(i % (size_t)nx); ly = (int)((i / (size_t)nx) % (size_t)ny); lz = (int)(i / ((size_t)nx * ny));
    }
    const int x = lx + ox, y = ly + oy, z = lz + oz;
    float v = kind == 0 ? synth_noise(x, y, z, gnx, gny, gnz, seed) : (kind == 1 ? synth_sparse(x, y, z, seed) : synth_smooth(x, y, z, seed));
    // kind 3: the smooth lattice as a very thin fog just around the transfer function's 0.1 threshold, so
    // that rays cross a 4096-voxel grid without saturating (every brick of a sort-last run does work)
    if (kind == 3) v = __fadd_rn(0.095f, __fmul_rn(0.033f, v));
    if (DTYPE == VKRT_U8) ((uint8_t*)out)[i] = (uint8_t)__float2int_rn(__fmul_rn(fminf(fmaxf(v, 0.0f), 1.0f), 255.0f));
    else if (DTYPE == VKRT_F16) ((__half*)out)[i] = __float2half_rn(v);
    else ((float*)out)[i] = v;
}

// ---- device-side flags for the sort-first exchange (system scope: they live in rank 0's memory and
// are touched by every GPU of the node over NVLink) ------------------------------------------------
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spins until *flag >= target. Gives up after 10 s (a dead peer must not hang the GPU) and counts
// the timeout in `timeouts` so the host can report it.
__global__ void flag_wait_kernel(const unsigned long long* flag, unsigned long long target, unsigned long long* timeouts) {
    const volatile unsigned long long* f = flag;
    const unsigned long long t0 = global_timer_ns();
    while (*f < target) {
        __nanosleep(200);
        if (global_timer_ns() - t0 > 10000000000ull) {
            atomicAdd_system(timeouts, 1ull);
            break;
        }
    }
    __threadfence_system();
}
__global__ void flag_add_kernel(unsigned long long* flag, unsigned long long v) {
    __threadfence_system();  // (the preceding kernel in this stream has completed: its peer stores are performed)
    atomicAdd_system(flag, v);
}
__global__ void flag_set_kernel(unsigned long long* flag, unsigned long long v) {
    __threadfence_system();
    *(volatile unsigned long long*)flag = v;
}

// Sort-first: ship the tiles a rank has rendered into its LOCAL frame to the same pixels of a frame in rank 0's memory
// (a CUDA-IPC mapping: the stores go out over NVLink). Whole tile rows with 16-byte stores by consecutive lanes, so the
// link carries full 128-byte lines — the raycast kernel's own 8-byte stores of 8x4-pixel warps (64-byte row segments)
// cost its launch 20-30 % (profiles/scaling_r01.md). grid = (tile, group of 16 rows).
// clip: per frame of the launch, the pixel rectangle (inclusive; x0 even, x1 odd) outside of which every pixel holds the
// clear colour — the cull rectangle around the projected box, api.cu push_clip — and is NOT shipped: rank 0 fills those
// pixels of the slot itself (fill_outside_kernel). At zoom 3 the box covers about a quarter of a 16:9 frame, so the
// gather moves a quarter of the bytes.
template <class V, int PX>
__global__ void __launch_bounds__(256) push_tiles_kernel(const uint2* __restrict__ src, uint2* __restrict__ dst, const VkrtOffset* __restrict__ offsets,
                                                         int tile, int W, int H, const PushClip clip) {
    const VkrtOffset o = offsets[blockIdx.x];
    const uint32_t x0 = __float2uint_rz(o.x), y0 = __float2uint_rz(o.y);
    if (x0 >= (uint32_t)W || y0 >= (uint32_t)H) return;
    src += (size_t)blockIdx.z * W * H;  // grid.z = frame of a batched tile share: consecutive local frames -> consecutive ring slots
    dst += (size_t)blockIdx.z * W * H;
    // the tile's pixels inside the clip rectangle: columns [xs, xe), rows [r0, r1) of this block's 16-row group
    const int xs = max((int)x0, clip.x0[blockIdx.z]), xe = min(min((int)x0 + tile, W), clip.x1[blockIdx.z] + 1);
    const int r0 = max((int)blockIdx.y * 16, clip.y0[blockIdx.z] - (int)y0);
    const int r1 = min(min(min((int)blockIdx.y * 16 + 16, tile), H - (int)y0), clip.y1[blockIdx.z] + 1 - (int)y0);
    if (xs >= xe || r0 >= r1) return;
    const int cols = (xe - xs) / PX;  // vectors per row (PX pixels each)
    for (int i = (int)threadIdx.x; i < (r1 - r0) * cols; i += (int)blockDim.x) {
        const int r = r0 + i / cols, cidx = i % cols;
        const size_t px = ((size_t)(y0 + r) * W + xs) + (size_t)cidx * PX;
        *reinterpret_cast<V*>(dst + px) = *reinterpret_cast<const V*>(src + px);
    }
    if (PX == 2 && ((xe - xs) & 1)) {  // odd width (the frame's last column): the last pixel of each row
        for (int r = r0 + (int)threadIdx.x; r < r1; r += (int)blockDim.x) {
            const size_t px = ((size_t)(y0 + r) * W + xe - 1);
            dst[px] = src[px];
        }
    }
}

// rank 0: every pixel of `frame` outside the rectangle (inclusive bounds; an empty rectangle = the whole frame) := texel
__global__ void __launch_bounds__(256) fill_outside_kernel(uint2* __restrict__ frame, int W, int H, int x0, int y0, int x1, int y1, uint2 texel) {
    const int y = (int)blockIdx.y;
    const bool row_inside = y >= y0 && y <= y1;
    for (int x = (int)(blockIdx.x * blockDim.x + threadIdx.x); x < W; x += (int)(gridDim.x * blockDim.x))
        if (!row_inside || x < x0 || x > x1) frame[(size_t)y * W + x] = texel;
}

}  // namespace

cudaError_t launch_push_tiles(const uint2* src, uint2* dst, const VkrtOffset* d_offsets, int n_tiles, int tile, int W, int H, bool vec16,
                              cudaStream_t s, int n_frames, const PushClip* clip) {
    if (n_tiles <= 0) return cudaSuccess;
    PushClip all;
    if (!clip) {
        for (int f = 0; f < kMaxPushDst + 1; ++f) { all.x0[f] = 0; all.y0[f] = 0; all.x1[f] = W - 1; all.y1[f] = H - 1; }
        clip = &all;
    }
    const dim3 grid((unsigned)n_tiles, (unsigned)((tile + 15) / 16), (unsigned)(n_frames > 0 ? n_frames : 1));
    if (vec16) push_tiles_kernel<uint4, 2><<<grid, 256, 0, s>>>(src, dst, d_offsets, tile, W, H, *clip);
    else push_tiles_kernel<uint2, 1><<<grid, 256, 0, s>>>(src, dst, d_offsets, tile, W, H, *clip);
    return cudaGetLastError();
}

cudaError_t launch_fill_outside(uint2* frame, int W, int H, int x0, int y0, int x1, int y1, uint2 texel, cudaStream_t s) {
    const dim3 grid((unsigned)std::min((W + 255) / 256, 64), (unsigned)H);
    fill_outside_kernel<<<grid, 256, 0, s>>>(frame, W, H, x0, y0, x1, y1, texel);
    return cudaGetLastError();
}

cudaError_t launch_scalar_to_rgba16f(const void* vol, int dtype, uint2* color, uint2* normal, int nx, int ny, int nz, cudaStream_t s) {
    const size_t total = (size_t)nx * ny * nz;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    switch (dtype) {
        case VKRT_U8: scalar_to_rgba16f_kernel<VKRT_U8><<<blocks, 256, 0, s>>>(vol, color, normal, nx, ny, nz); break;
        case VKRT_F16: scalar_to_rgba16f_kernel<VKRT_F16><<<blocks, 256, 0, s>>>(vol, color, normal, nx, ny, nz); break;
        case VKRT_F32: scalar_to_rgba16f_kernel<VKRT_F32><<<blocks, 256, 0, s>>>(vol, color, normal, nx, ny, nz); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// linear window -> bricked window (upload path; the generator writes bricked directly)
template <class T>
__global__ void __launch_bounds__(256) brick_window_kernel(const T* __restrict__ lin, T* __restrict__ out, int nx, int ny, int nz) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int bnx = (nx + 7) >> 3, bny = (ny + 7) >> 3, bnz = (nz + 7) >> 3;
    if (i >= (size_t)bnx * bny * bnz * 512) return;
    const size_t b = i >> 9;
    const int in = (int)(i & 511);
    const int lx = (int)(b % (size_t)bnx) * 8 + (in & 7), ly = (int)((b / (size_t)bnx) % (size_t)bny) * 8 + ((in >> 3) & 7),
              lz = (int)(b / ((size_t)bnx * bny)) * 8 + (in >> 6);
    out[i] = (lx < nx && ly < ny && lz < nz) ? lin[((size_t)lz * ny + ly) * nx + lx] : T(0);
}

cudaError_t launch_brick_window(const void* lin, void* out, int elem_bytes, int nx, int ny, int nz, cudaStream_t s) {
    const size_t total = (size_t)((nx + 7) >> 3) * ((ny + 7) >> 3) * ((nz + 7) >> 3) * 512;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (elem_bytes == 1) brick_window_kernel<uint8_t><<<blocks, 256, 0, s>>>((const uint8_t*)lin, (uint8_t*)out, nx, ny, nz);
    else if (elem_bytes == 2) brick_window_kernel<uint16_t><<<blocks, 256, 0, s>>>((const uint16_t*)lin, (uint16_t*)out, nx, ny, nz);
    else brick_window_kernel<uint32_t><<<blocks, 256, 0, s>>>((const uint32_t*)lin, (uint32_t*)out, nx, ny, nz);
    return cudaGetLastError();
}

cudaError_t launch_synth(void* out, int kind, int dtype, int nx, int ny, int nz, int ox, int oy, int oz, int gnx, int gny, int gnz,
                         uint32_t seed, cudaStream_t s, int bricked) {
    const size_t total = bricked ? (size_t)((nx + 7) >> 3) * ((ny + 7) >> 3) * ((nz + 7) >> 3) * 512 : (size_t)nx * ny * nz;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    switch (dtype) {
        case VKRT_U8: synth_kernel<VKRT_U8><<<blocks, 256, 0, s>>>(out, kind, nx, ny, nz, ox, oy, oz, gnx, gny, gnz, seed, bricked); break;
        case VKRT_F16: synth_kernel<VKRT_F16><<<blocks, 256, 0, s>>>(out, kind, nx, ny, nz, ox, oy, oz, gnx, gny, gnz, seed, bricked); break;
        case VKRT_F32: synth_kernel<VKRT_F32><<<blocks, 256, 0, s>>>(out, kind, nx, ny, nz, ox, oy, oz, gnx, gny, gnz, seed, bricked); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_flag_wait(const unsigned long long* flag, unsigned long long target, unsigned long long* timeouts, cudaStream_t s) {
    flag_wait_kernel<<<1, 1, 0, s>>>(flag, target, timeouts);
    return cudaGetLastError();
}
cudaError_t launch_flag_add(unsigned long long* flag, unsigned long long v, cudaStream_t s) {
    flag_add_kernel<<<1, 1, 0, s>>>(flag, v);
    return cudaGetLastError();
}
cudaError_t launch_flag_set(unsigned long long* flag, unsigned long long v, cudaStream_t s) {
    flag_set_kernel<<<1, 1, 0, s>>>(flag, v);
    return cudaGetLastError();
}

// Sort-last direct-send (vkrt_exchange_*): one pass over a W*H float image, stored into up to kMaxPushDst peer tables over
// NVLink — consecutive lanes write consecutive 16-byte vectors, so the links carry full 128-byte lines; the image is
// read once however many ranks need it.
__global__ void __launch_bounds__(256) push_many_kernel(const float4* __restrict__ src, PushDst d, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(src + i);
        for (int k = 0; k < d.n; ++k) reinterpret_cast<float4*>(d.ptr[k])[i] = v;
    }
}
// spin until every flag >= target (slot reuse: the receivers have resolved the frame that used this parity before)
__global__ void flag_wait_many_kernel(PushDst d, unsigned long long target, unsigned long long* timeouts) {
    const unsigned long long t0 = global_timer_ns();
    for (int k = 0; k < d.n; ++k) {
        const volatile unsigned long long* f = reinterpret_cast<const volatile unsigned long long*>(d.ptr[k]);
        while (*f < target) {
            __nanosleep(200);
            if (global_timer_ns() - t0 > 10000000000ull) {
                atomicAdd_system(timeouts, 1ull);
                break;
            }
        }
    }
    __threadfence_system();
}
__global__ void flag_add_many_kernel(PushDst d, unsigned long long v) {
    __threadfence_system();  // (the push kernel before this one in the stream has completed: its peer stores are performed)
    for (int k = 0; k < d.n; ++k) atomicAdd_system(reinterpret_cast<unsigned long long*>(d.ptr[k]), v);
}
cudaError_t launch_push_many(const float* src, const PushDst& d, size_t n, cudaStream_t s) {
    if (d.n > 0) push_many_kernel<<<148 * 8, 256, 0, s>>>(reinterpret_cast<const float4*>(src), d, n / 4);
    return cudaGetLastError();
}
cudaError_t launch_flag_wait_many(const PushDst& d, unsigned long long target, unsigned long long* timeouts, cudaStream_t s) {
    if (d.n > 0) flag_wait_many_kernel<<<1, 1, 0, s>>>(d, target, timeouts);
    return cudaGetLastError();
}
cudaError_t launch_flag_add_many(const PushDst& d, unsigned long long v, cudaStream_t s) {
    if (d.n > 0) flag_add_many_kernel<<<1, 1, 0, s>>>(d, v);
    return cudaGetLastError();
}

cudaError_t launch_flush_l2(uint4* buf, size_t n16, cudaStream_t s) {
    static uint32_t tag = 0;
    flush_l2_kernel<<<148 * 8, 256, 0, s>>>(buf, n16, ++tag);
    return cudaGetLastError();
}

// LAYOUT_QUAD staging: texel (tx, ty, z) of an (nx+1) x (ny+1) x nz grid = the 2x2 xy footprint whose low corner is
// voxel (tx-1, ty-1, z), both taps of each axis clamped to the grid (clamp-to-edge, like the oracle's scalar_at()).
// Channels: (x0,y0), (x1,y0), (x0,y1), (x1,y1). A sample with x0 = floor(qx - 0.5) in [-1, nx-1] reads texel x0 + 1.
template <class T4, class T>
__global__ void __launch_bounds__(256) pregather_quads_kernel(const T* __restrict__ vol, T4* __restrict__ out, int nx, int ny, int nz) {
    const size_t total = (size_t)(nx + 1) * (ny + 1) * nz;
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    const int tx = (int)(o % (size_t)(nx + 1)), ty = (int)((o / (size_t)(nx + 1)) % (size_t)(ny + 1)), z = (int)(o / ((size_t)(nx + 1) * (ny + 1)));
    const int x0 = max(tx - 1, 0), x1 = min(tx, nx - 1), y0 = max(ty - 1, 0), y1 = min(ty, ny - 1);
    const size_t r0 = ((size_t)z * ny + y0) * nx, r1 = ((size_t)z * ny + y1) * nx;
    T4 q;
    q.x = __ldg(vol + r0 + x0); q.y = __ldg(vol + r0 + x1); q.z = __ldg(vol + r1 + x0); q.w = __ldg(vol + r1 + x1);
    out[o] = q;
}

cudaError_t launch_pregather_quads(const void* vol, int dtype, void* out, int nx, int ny, int nz, cudaStream_t s) {
    const size_t total = (size_t)(nx + 1) * (ny + 1) * nz;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    switch (dtype) {
        case VKRT_U8: pregather_quads_kernel<uchar4, unsigned char><<<blocks, 256, 0, s>>>((const unsigned char*)vol, (uchar4*)out, nx, ny, nz); break;
        case VKRT_F16: pregather_quads_kernel<ushort4, unsigned short><<<blocks, 256, 0, s>>>((const unsigned short*)vol, (ushort4*)out, nx, ny, nz); break;
        case VKRT_F32: pregather_quads_kernel<float4, float><<<blocks, 256, 0, s>>>((const float*)vol, (float4*)out, nx, ny, nz); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_interleave_bricked(const uint2* color, const uint2* normal, uint4* out, int nx, int ny, int nz, int nbx, int nby,
                                      int nbz, cudaStream_t s) {
    const size_t total = (size_t)nbx * nby * nbz * 512;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    interleave_bricked_kernel<<<blocks, 256, 0, s>>>(color, normal, out, nx, ny, nz, nbx, nby, total);
    return cudaGetLastError();
}

cudaError_t launch_occupancy_m0(const uint2* color, const uint2* normal, int nx, int ny, int nz, int nbx, int nby, int nbz, int bs, uint8_t* dist,
                                cudaStream_t s) {
    const size_t cells = (size_t)nbx * nby * nbz;
    occupancy_m0_kernel<<<(unsigned)((cells + 7) / 8), 256, 0, s>>>(color, normal, nx, ny, nz, nbx, nby, nbz, bs, dist);
    return cudaGetLastError();
}

// After max_d relaxation steps every brick holds min(true distance, max_d); the result is left in `dist`.
cudaError_t launch_distance_transform(uint8_t* dist, uint8_t* scratch, int nbx, int nby, int nbz, int border, int max_d, cudaStream_t s) {
    const size_t cells = (size_t)nbx * nby * nbz;
    const unsigned blocks = (unsigned)((cells + 255) / 256);
    uint8_t *a = dist, *b = scratch;
    for (int i = 0; i < max_d; ++i) {
        distance_step_kernel<<<blocks, 256, 0, s>>>(a, b, nbx, nby, nbz, border, max_d);
        uint8_t* t = a; a = b; b = t;
    }
    if (a != dist) {
        cudaError_t e = cudaMemcpyAsync(dist, a, cells, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

// occ: occupancy (0 / 255), cells bytes. oct, scratch: 8 * cells bytes each; d_flag: one int. After k relaxation steps every
// table holds min(true directional distance, k) (255 where nothing within k is occupied); the loop stops at max_d steps or,
// checked every 8 steps, as soon as a step changes nothing (then every value is final, and min(value, max_d) is applied by
// the step's own cap). Synchronises the stream. The result is left in `oct`.
cudaError_t launch_octant_distance(const uint8_t* occ, uint8_t* oct, uint8_t* scratch, int* d_flag, int nbx, int nby, int nbz, int border, int max_d,
                                   cudaStream_t s) {
    const size_t cells = (size_t)nbx * nby * nbz;
    const unsigned blocks = (unsigned)((8 * cells + 255) / 256);
    replicate8_kernel<<<blocks, 256, 0, s>>>(occ, oct, cells);
    uint8_t *a = oct, *b = scratch;
    for (int i = 0; i < max_d; ++i) {
        const bool probe = (i % 8) == 7 && i + 1 < max_d;
        cudaError_t e = probe ? cudaMemsetAsync(d_flag, 0, sizeof(int), s) : cudaSuccess;
        if (e != cudaSuccess) return e;
        octant_step_kernel<<<blocks, 256, 0, s>>>(a, b, nbx, nby, nbz, border, max_d, probe ? d_flag : nullptr);
        uint8_t* t = a; a = b; b = t;
        if (probe) {
            int changed = 1;
            e = cudaMemcpyAsync(&changed, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) return e;
            if (!changed) break;
        }
    }
    if (a != oct) {
        cudaError_t e = cudaMemcpyAsync(oct, a, 8 * cells, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

cudaError_t launch_occupied_bounds(const uint8_t* dist, int nbx, int nby, int nbz, int* d_out7, cudaStream_t s) {
    static const int init[7] = {0x7fffffff, 0x7fffffff, 0x7fffffff, -1, -1, -1, 0};
    cudaError_t e = cudaMemcpyAsync(d_out7, init, sizeof init, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    const size_t cells = (size_t)nbx * nby * nbz;
    occupied_bounds_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(dist, nbx, nby, nbz, d_out7);
    return cudaGetLastError();
}

// The distance field with one layer of "occupied" cells appended on the high side of every axis (M1, see RenderArgs::dist):
// a sample whose voxel index is one past the grid (q == N, reached when p rounds to the box face) or slightly negative in a
// wrapped index lands on a pad cell and is simply evaluated, so the march needs no per-sample bounds test.
__global__ void __launch_bounds__(256) pad_dist_kernel(const uint8_t* __restrict__ dist, uint8_t* __restrict__ out, int nbx, int nby, int nbz, int tables) {
    const size_t padded = (size_t)(nbx + 1) * (nby + 1) * (nbz + 1), cells = (size_t)nbx * nby * nbz;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= padded * (size_t)tables) return;
    const size_t k = g / padded, o = g - k * padded;
    const int bx = (int)(o % (size_t)(nbx + 1)), by = (int)((o / (size_t)(nbx + 1)) % (size_t)(nby + 1)), bz = (int)(o / ((size_t)(nbx + 1) * (nby + 1)));
    out[g] = (bx < nbx && by < nby && bz < nbz) ? dist[k * cells + ((size_t)bz * nby + by) * nbx + bx] : (uint8_t)0;
}
// `tables` fields back to back on both sides
cudaError_t launch_pad_dist(const uint8_t* dist, uint8_t* out, int nbx, int nby, int nbz, int tables, cudaStream_t s) {
    const size_t total = (size_t)(nbx + 1) * (nby + 1) * (nbz + 1) * (size_t)tables;
    pad_dist_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(dist, out, nbx, nby, nbz, tables);
    return cudaGetLastError();
}

cudaError_t launch_occupancy_m1(const void* scalar, int dtype, int nx, int ny, int nz, int nbx, int nby, int nbz, int bs, uint8_t* dist,
                                cudaStream_t s) {
    const size_t cells = (size_t)nbx * nby * nbz;
    const unsigned blocks = (unsigned)((cells + 7) / 8);
    switch (dtype) {
        case VKRT_U8: occupancy_m1_kernel<VKRT_U8><<<blocks, 256, 0, s>>>(scalar, nx, ny, nz, nbx, nby, nbz, bs, dist); break;
        case VKRT_F16: occupancy_m1_kernel<VKRT_F16><<<blocks, 256, 0, s>>>(scalar, nx, ny, nz, nbx, nby, nbz, bs, dist); break;
        case VKRT_F32: occupancy_m1_kernel<VKRT_F32><<<blocks, 256, 0, s>>>(scalar, nx, ny, nz, nbx, nby, nbz, bs, dist); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_generate_xor(uint2* color, uint2* normal, int n, float time, int which, cudaStream_t s) {
    const size_t total = (size_t)n * n * n;
    generate_xor_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(color, normal, n, time, which);
    return cudaGetLastError();
}

}  // namespace vkrt
