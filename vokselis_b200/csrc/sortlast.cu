// sortlast.cu — brick-partitioned (sort-last) rendering for volumes larger than one GPU (BASELINE config 5).
//
// The reference is single-device; this extends its march (shaders/raycast_compute.wgsl:62-97 for M0,
// shaders/raycast_naive.wgsl:96-119 for M1) to a rank that holds only a WINDOW of the global grid: its
// own brick [own_lo, own_hi) plus a one-voxel halo. Every rank walks the GLOBAL t sequence of each ray
// (same floats as the single-GPU loop) and evaluates exactly the samples whose voxel index falls in its
// brick; all other samples are no-ops for it. Early termination needs the alpha accumulated in front of
// the brick, so a frame is two passes:
//   PASS_ALPHA : per pixel, transmittance T_r = prod(1 - alpha_k) over the rank's samples
//   (exchange) : a_in = 1 - (1 - a0) * prod of T over the bricks in front (visibility order)
//   PASS_COLOR : the reference's recurrence from (0,0,0,a_in), early termination at the threshold;
//                output = premultiplied rgb, so the composite is a plain SUM over ranks
//   finalize   : frame = clear/initial colour + sum of partials on hit pixels, clear colour elsewhere
#include "raycast.cuh"
#include "vkrt_device.cuh"

namespace vkrt {

namespace {

enum { PASS_ALPHA = 1, PASS_COLOR = 2 };

template <int DTYPE> __device__ __forceinline__ float win_load(const void* p, size_t i);
template <> __device__ __forceinline__ float win_load<VKRT_U8>(const void* p, size_t i) {
    return __uint_as_float(0x4B000000u | (uint32_t)__ldg((const uint8_t*)p + i)) - 8388608.0f;
}
template <> __device__ __forceinline__ float win_load<VKRT_F16>(const void* p, size_t i) {
    return __half2float(__ushort_as_half(__ldg((const unsigned short*)p + i)));
}
template <> __device__ __forceinline__ float win_load<VKRT_F32>(const void* p, size_t i) { return __ldg((const float*)p + i); }

// Raw taps of one sample: all eight loads are issued before anything consumes them (volatile keeps ptxas from
// sinking the first interpolation between the loads, which cut the loads in flight per thread from 8 to 2-3 and
// cost 13 % at 32 GiB per rank, where the march waits on HBM).
template <int DTYPE> __device__ __forceinline__ uint32_t win_raw(const void* p) {
    uint32_t v;
    if (DTYPE == VKRT_U8) asm("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    else if (DTYPE == VKRT_F16) asm("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p));
    else asm("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <int DTYPE> __device__ __forceinline__ float win_cvt(uint32_t v) {
    if (DTYPE == VKRT_U8) return __uint_as_float(0x4B000000u | v) - 8388608.0f;
    if (DTYPE == VKRT_F16) return __half2float(__ushort_as_half((unsigned short)v));
    return __uint_as_float(v);
}

// Element index of window-local voxel (lx, ly, lz) as a sum of three per-axis parts (occupancy pass).
__device__ __forceinline__ size_t part_x(const PartialArgs& A, int lx) { return A.bricked ? (size_t)(lx >> 3) * 512 + (lx & 7) : (size_t)lx; }
__device__ __forceinline__ size_t part_y(const PartialArgs& A, int ly) {
    return A.bricked ? (size_t)(ly >> 3) * A.bnx * 512 + (size_t)(ly & 7) * 8 : (size_t)ly * A.nx;
}
__device__ __forceinline__ size_t part_z(const PartialArgs& A, int lz) {
    return A.bricked ? (size_t)(lz >> 3) * A.bny * A.bnx * 512 + (size_t)(lz & 7) * 64 : (size_t)lz * A.nx * A.ny;
}

template <int DTYPE> struct WinElem;
template <> struct WinElem<VKRT_U8> { typedef uint8_t type; };
template <> struct WinElem<VKRT_F16> { typedef unsigned short type; };
template <> struct WinElem<VKRT_F32> { typedef float type; };

// Trilinear sample of the window array; taps clamp to the GLOBAL grid, then shift by the window origin.
//
// Bricked windows (scalar data) store 8^3-voxel bricks contiguously — 2 KB for fp32 — so the 8 taps of a
// sample fall into one or two bricks instead of 8 cache lines that are 16 KB (a row) and 64 MB (a slice)
// apart at 4096^3: far fewer sectors, DRAM pages and TLB entries per sample.
//
// Addressing: ONE 64-bit element index for the low corner (xa, ya, za), plus three 32-bit strides to the
// high tap of each axis (1 / 8 / 64 inside a brick, the jump to the neighbouring brick when the low tap sits
// on the brick's last plane, 0 where the global clamp folds both taps onto one voxel). The other seven
// addresses are 32-bit sums of those strides added to the corner pointer (one IMAD.WIDE each). The first
// version built six 64-bit per-axis parts and four 64-bit row sums per sample; ncu showed that address
// arithmetic as 29 % of the kernel's instructions on an ALU-bound kernel (profiles/r01_v3_prof_final_partial.md).
struct Taps {
    uint32_t r[8];     // raw taps c000 c100 c010 c110 c001 c101 c011 c111
    float fx, fy, fz;  // fractional weights
};
// WINDOW: also clamp the low corner into the window (a SPECULATIVE fetch of a sample that may turn out not to be this
// rank's: the address must be valid, the value is thrown away if the sample is not evaluated).
template <int DTYPE, bool WINDOW> __device__ __forceinline__ void win_fetch(const PartialArgs& A, float qx, float qy, float qz, Taps& T) {
    const float ux = __fsub_rn(qx, 0.5f), uy = __fsub_rn(qy, 0.5f), uz = __fsub_rn(qz, 0.5f);
    const float flx = floorf(ux), fly = floorf(uy), flz = floorf(uz);
    T.fx = __fsub_rn(ux, flx); T.fy = __fsub_rn(uy, fly); T.fz = __fsub_rn(uz, flz);
    const int x0 = (int)flx, y0 = (int)fly, z0 = (int)flz;
    const int xa = min(max(x0, 0), A.gnx - 1), ya = min(max(y0, 0), A.gny - 1), za = min(max(z0, 0), A.gnz - 1);
    bool mx = xa != min(max(x0 + 1, 0), A.gnx - 1), my = ya != min(max(y0 + 1, 0), A.gny - 1), mz = za != min(max(z0 + 1, 0), A.gnz - 1);
    uint32_t lx = (uint32_t)(xa - A.wx), ly = (uint32_t)(ya - A.wy), lz = (uint32_t)(za - A.wz);
    if (WINDOW) {  // the high taps of a clamped corner fold onto it (strides 0) unless they are inside the window too
        const int cx = min(max(xa - A.wx, 0), A.nx - 1), cy = min(max(ya - A.wy, 0), A.ny - 1), cz = min(max(za - A.wz, 0), A.nz - 1);
        mx = mx && cx + 1 < A.nx; my = my && cy + 1 < A.ny; mz = mz && cz + 1 < A.nz;
        lx = (uint32_t)cx; ly = (uint32_t)cy; lz = (uint32_t)cz;
    }
    size_t base;
    uint32_t dx, dy, dz;
    if (A.bricked) {
        const uint32_t rowb = (uint32_t)A.bnx * 512u, slabb = (uint32_t)A.bny * rowb;  // elements per row / slab of bricks
        const uint32_t brick = ((lz >> 3) * (uint32_t)A.bny + (ly >> 3)) * (uint32_t)A.bnx + (lx >> 3);
        base = ((size_t)brick << 9) + (((lz & 7u) << 6) | ((ly & 7u) << 3) | (lx & 7u));
        dx = mx ? ((lx & 7u) == 7u ? 512u - 7u : 1u) : 0u;
        dy = my ? ((ly & 7u) == 7u ? rowb - 56u : 8u) : 0u;
        dz = mz ? ((lz & 7u) == 7u ? slabb - 448u : 64u) : 0u;
    } else {
        base = ((size_t)lz * (uint32_t)A.ny + ly) * (uint32_t)A.nx + lx;
        dx = mx ? 1u : 0u;
        dy = my ? (uint32_t)A.nx : 0u;
        dz = mz ? (uint32_t)A.nx * (uint32_t)A.ny : 0u;
    }
    const typename WinElem<DTYPE>::type* c = (const typename WinElem<DTYPE>::type*)A.vol_a + base;
    const uint32_t dxy = dx + dy;
    const typename WinElem<DTYPE>::type *p100 = c + dx, *p010 = c + dy, *p110 = c + dxy, *p001 = c + dz, *p101 = c + (dz + dx),
                                        *p011 = c + (dz + dy), *p111 = c + (dz + dxy);
    T.r[0] = win_raw<DTYPE>(c); T.r[1] = win_raw<DTYPE>(p100); T.r[2] = win_raw<DTYPE>(p010); T.r[3] = win_raw<DTYPE>(p110);
    T.r[4] = win_raw<DTYPE>(p001); T.r[5] = win_raw<DTYPE>(p101); T.r[6] = win_raw<DTYPE>(p011); T.r[7] = win_raw<DTYPE>(p111);
}
template <int DTYPE> __device__ __forceinline__ float win_resolve(const Taps& T) {
    const float c000 = win_cvt<DTYPE>(T.r[0]), c100 = win_cvt<DTYPE>(T.r[1]), c010 = win_cvt<DTYPE>(T.r[2]), c110 = win_cvt<DTYPE>(T.r[3]);
    const float c001 = win_cvt<DTYPE>(T.r[4]), c101 = win_cvt<DTYPE>(T.r[5]), c011 = win_cvt<DTYPE>(T.r[6]), c111 = win_cvt<DTYPE>(T.r[7]);
    const float fx = T.fx, fy = T.fy, fz = T.fz;
    const float c00 = fmaf(fx, __fsub_rn(c100, c000), c000), c10 = fmaf(fx, __fsub_rn(c110, c010), c010);
    const float c01 = fmaf(fx, __fsub_rn(c101, c001), c001), c11 = fmaf(fx, __fsub_rn(c111, c011), c011);
    const float c0 = fmaf(fy, __fsub_rn(c10, c00), c00), c1 = fmaf(fy, __fsub_rn(c11, c01), c01);
    return __fmul_rn(fmaf(fz, __fsub_rn(c1, c0), c0), DTYPE == VKRT_U8 ? 1.0f / 255.0f : 1.0f);
}

// speculative fetch of the sample at parameter t (exact position arithmetic: the very q the loop computes for that t)
template <int DTYPE> __device__ __forceinline__ void fetch_at(const PartialArgs& A, f3 eye, f3 dir, float t, Taps& T) {
    const float qx = xmul(xadd(xadd(eye.x, xmul(t, dir.x)), 1.0f), A.hx), qy = xmul(xadd(xadd(eye.y, xmul(t, dir.y)), 1.0f), A.hy),
                qz = xmul(xadd(xadd(eye.z, xmul(t, dir.z)), 1.0f), A.hz);
    win_fetch<DTYPE, true>(A, qx, qy, qz, T);
}

template <int MODE, int DTYPE, int PASS>
__global__ void __launch_bounds__(256) partial_kernel(const __grid_constant__ PartialArgs A) {
    const uint32_t px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= (uint32_t)A.W || py >= (uint32_t)A.H) return;
    const size_t o = (size_t)py * A.W + px;
    f3 eye, dir;
    gen_ray(A.inv, (float)px, (float)py, 0.0f, 0.0f, (float)A.W, (float)A.H, eye, dir);
    float t0, t1;
    intersect_box(eye, dir, t0, t1);
    const bool hit = t0 < t1;
    t0 = fmaxf(t0, 0.0f);

    float T = 1.0f;  // PASS_ALPHA
    Rgba col = {0.f, 0.f, 0.f, 0.f};
    if (PASS == PASS_COLOR) {
        // a_in == nullptr: RELATIVE pass (march from alpha 0; the result is scaled by the incoming
        // transmittance afterwards, resolve_kernel). a_in[o] < 0: this pixel needs no (re-)march.
        const float ain = A.a_in ? A.a_in[o] : 0.0f;
        if (A.a_in && ain < 0.0f) return;
        col.a = hit ? ain : 1.0f;
    }
    bool live = hit && (PASS == PASS_ALPHA || col.a < A.alpha_threshold);

    if (live) {
        const float dt = step_dt(dir, A.fx, A.fy, A.fz, A.dt_scale, A.dt_floor);
        // Approximate entry/exit of the own brick along the ray, in samples, with a 2-sample margin: the
        // samples in between are tested exactly, everything before is leapt over with advance_t (which
        // lands on the very floats the repeated addition visits) and everything after is not mine.
        float te = t0, tx = t1;
        {
            const float lo[3] = {(float)A.own_lo[0] / A.hx - 1.0f, (float)A.own_lo[1] / A.hy - 1.0f, (float)A.own_lo[2] / A.hz - 1.0f};
            const float hi[3] = {(float)A.own_hi[0] / A.hx - 1.0f, (float)A.own_hi[1] / A.hy - 1.0f, (float)A.own_hi[2] / A.hz - 1.0f};
            const float o3[3] = {eye.x, eye.y, eye.z}, d3[3] = {dir.x, dir.y, dir.z};
            for (int i = 0; i < 3; ++i) {
                if (fabsf(d3[i]) > 1e-12f) {
                    const float a = (lo[i] - o3[i]) / d3[i], b = (hi[i] - o3[i]) / d3[i];
                    te = fmaxf(te, fminf(a, b));
                    tx = fminf(tx, fmaxf(a, b));
                } else if (o3[i] < lo[i] - 1e-4f || o3[i] > hi[i] + 1e-4f) {
                    tx = -1.0f;  // parallel and outside the slab
                }
            }
        }
        float t = t0;
        if (te - 2.0f * dt <= tx + 2.0f * dt) {
            // Each replayed addition rounds by <= ulp(t)/2 <= t1 * 2^-24, always the same way inside a binade, so
            // after s steps the t sequence may run ahead of t0 + s*dt by s * drift steps (several steps at
            // 4096^3, where a ray takes ~10^4 of them): keep that, doubled, plus 2 samples, in hand.
            const float s_pre = (te - t0) / dt;
            const float drift_per_step = (t1 * 5.9604645e-08f) / dt * 2.0f;
            const int n_pre = (int)fminf(s_pre - fmaf(s_pre, drift_per_step, 2.0f), 1.0e9f);
            if (n_pre > 0) t = advance_t(t, dt, n_pre);
            const float t_stop = fminf(t1, tx + 2.0f * dt);
            const float dqx = dir.x * A.hx * dt, dqy = dir.y * A.hy * dt, dqz = dir.z * A.hz * dt;
            const float rqx = fabsf(dqx) > 1e-12f ? 1.0f / dqx : 1e30f, rqy = fabsf(dqy) > 1e-12f ? 1.0f / dqy : 1e30f,
                        rqz = fabsf(dqz) > 1e-12f ? 1.0f / dqz : 1e30f;
            Taps cur, pre;
            float pre_t = -1.0f;  // the t whose taps `pre` holds (t >= 0 always)
            while (t < t1 && t < t_stop) {
                f3 p = {xadd(eye.x, xmul(t, dir.x)), xadd(eye.y, xmul(t, dir.y)), xadd(eye.z, xmul(t, dir.z))};
                const float qx = xmul(xadd(p.x, 1.0f), A.hx), qy = xmul(xadd(p.y, 1.0f), A.hy), qz = xmul(xadd(p.z, 1.0f), A.hz);
                const int ix = __float2int_rz(qx), iy = __float2int_rz(qy), iz = __float2int_rz(qz);
                const bool inb = (unsigned)ix < (unsigned)A.gnx && (unsigned)iy < (unsigned)A.gny && (unsigned)iz < (unsigned)A.gnz;
                // owner = the brick holding the (M1: clamped) voxel index; M0 samples outside the grid read 0: nobody's
                const int cx = min(max(ix, 0), A.gnx - 1), cy = min(max(iy, 0), A.gny - 1), cz = min(max(iz, 0), A.gnz - 1);
                bool mine = cx >= A.own_lo[0] && cx < A.own_hi[0] && cy >= A.own_lo[1] && cy < A.own_hi[1] && cz >= A.own_lo[2] && cz < A.own_hi[2];
                if (MODE == VKRT_MODE_M0) mine = mine && inb;
                int n = 1;
                if (mine && inb) {
                    const int lcx = (ix >> 3) - A.cox, lcy = (iy >> 3) - A.coy, lcz = (iz >> 3) - A.coz;
                    const uint32_t d = __ldg(A.dist + ((size_t)lcz * A.cny + lcy) * A.cnx + lcx);
                    if (d == 0u) {
                        n = 0;
                    } else {
                        const int r = (int)d - 1;
                        float lox = (float)(((ix >> 3) - r) * 8), loy = (float)(((iy >> 3) - r) * 8), loz = (float)(((iz >> 3) - r) * 8);
                        float hix = (float)(((ix >> 3) + r + 1) * 8), hiy = (float)(((iy >> 3) + r + 1) * 8), hiz = (float)(((iz >> 3) + r + 1) * 8);
                        if (MODE == VKRT_MODE_M1) {  // clamp-to-edge: samples at the grid faces are not empty
                            lox = fmaxf(lox, 0.0f); loy = fmaxf(loy, 0.0f); loz = fmaxf(loz, 0.0f);
                            hix = fminf(hix, A.fx); hiy = fminf(hiy, A.fy); hiz = fminf(hiz, A.fz);
                        }
                        const float sx = ((dqx > 0.f ? hix - A.leap_eps : lox + A.leap_eps) - qx) * rqx;
                        const float sy = ((dqy > 0.f ? hiy - A.leap_eps : loy + A.leap_eps) - qy) * rqy;
                        const float sz = ((dqz > 0.f ? hiz - A.leap_eps : loz + A.leap_eps) - qz) * rqz;
                        n = max(__float2int_rz(fminf(fminf(sx, sy), fminf(sz, 4096.0f))) - 1, 1);
                    }
                } else if (mine) {
                    n = 0;  // M1 sample outside the grid owned through its clamped index: evaluate
                }
                if (n > 0) {
                    if (n >= 16) t = advance_t(t, dt, n);
                    else for (int j = 0; j < n; ++j) t = xadd(t, dt);
                    continue;
                }
                if (MODE == VKRT_MODE_M0) {
                    const size_t i = ((size_t)(iz - A.wz) * A.ny + (iy - A.wy)) * A.nx + (ix - A.wx);
                    const float4 c = unpack_rgba16f(__ldg(reinterpret_cast<const uint2*>(A.vol_a) + i));
                    if (PASS == PASS_ALPHA) {
                        T *= 1.0f - m0_alpha(c.w);
                    } else {
                        const float4 nrm = unpack_rgba16f(__ldg(reinterpret_cast<const uint2*>(A.vol_b) + i));
                        m0_shade(col, c, nrm, p, A.clear);
                    }
                } else {
                    // Two samples in flight. At 32-128 GiB per rank a sample's taps mostly miss L1 and L2, the march waits
                    // on that latency (ncu at 32 GiB: issue active 21 %, DRAM 35 %, long-scoreboard stalls 34 of the 41
                    // cycles between issues; more resident warps make it SLOWER — they evict each other's sectors —
                    // fewer make it faster, profiles/r02_sortlast_march.md), and the next sample's addresses depend on
                    // nothing but t. So the taps of t + dt are requested before this sample's are consumed: every second
                    // wait finds its data already there. A speculative fetch that is not used (leap, other rank's
                    // voxel, termination) costs its loads and nothing else; the values are those the unpipelined loop reads.
                    // (Three in flight: 98 registers, faster for some view directions and slower for others; not kept.)
                    if (!(pre_t == t)) win_fetch<DTYPE, false>(A, qx, qy, qz, cur);
                    else cur = pre;
                    const float tn = xadd(t, dt);
                    pre_t = -1.0f;
                    if (tn < t1 && tn < t_stop) {
                        fetch_at<DTYPE>(A, eye, dir, tn, pre);
                        pre_t = tn;
                    }
                    const float s = win_resolve<DTYPE>(cur);
                    if (PASS == PASS_ALPHA) T *= 1.0f - m1_alpha(s);
                    else m1_shade(col, s);
                }
                if (PASS == PASS_ALPHA ? (T <= 1.0f - A.alpha_threshold) : (col.a >= A.alpha_threshold)) break;
                t = xadd(t, dt);
            }
        }
    }
    if (PASS == PASS_ALPHA) {
        A.T_out[o] = T;
    } else {
        A.rgba_out[o] = make_float4(col.r, col.g, col.b, col.a);
        if (A.T_out) A.T_out[o] = hit ? 1.0f - col.a : 1.0f;  // relative pass: transmittance of the brick
    }
}

// Deferred early termination (one march instead of two). Input: the RELATIVE partials
// (premultiplied rgb and alpha accumulated from 0 inside the brick) and every rank's transmittance.
//   a_in >= threshold                      -> the ray ended in front of this brick: partial = 0
//   the 0.95 crossing can fall inside the brick, or the relative march stopped on its own alpha
//                                          -> re-march this pixel from a_in (a_in_out >= 0)
//   otherwise                              -> partial = (1 - a_in) * relative partial (a_in_out = -1)
__global__ void __launch_bounds__(256) resolve_kernel(const float* __restrict__ T_all, size_t stride, const int* __restrict__ before,
                                                      int n_before, float a0, float thr, float4* __restrict__ rgba,
                                                      float* __restrict__ a_in_out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float T = 1.0f - a0;
    for (int k = 0; k < n_before; ++k) T *= __ldg(T_all + (size_t)before[k] * stride + i);
    const float ain = 1.0f - T;
    float4 c = rgba[i];
    float flag = -1.0f;
    if (ain >= thr) {
        c = make_float4(0.f, 0.f, 0.f, ain);
    } else {
        const float a_out = 1.0f - T * (1.0f - c.w);
        if (a_out >= thr - 1e-4f || c.w >= thr - 1e-4f) {
            flag = ain;  // exact recurrence needed: PASS_COLOR overwrites rgba[i]
        } else {
            c = make_float4(T * c.x, T * c.y, T * c.z, a_out);
        }
    }
    rgba[i] = c;
    a_in_out[i] = flag;
}

// a_in = 1 - (1 - a0) * prod_{j in before} T_j   (T_all = [world][W*H])
__global__ void __launch_bounds__(256) ain_kernel(const float* __restrict__ T_all, size_t stride, const int* __restrict__ before, int n_before,
                                                  float a0, float* __restrict__ a_in, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float T = 1.0f - a0;
    for (int k = 0; k < n_before; ++k) T *= __ldg(T_all + (size_t)before[k] * stride + i);
    a_in[i] = 1.0f - T;
}

// frame = initial colour + sum of partials (hit) | clear colour (miss); alpha = 1
__global__ void __launch_bounds__(256) finalize_kernel(const __grid_constant__ PartialArgs A, const float4* __restrict__ sum, uint2* __restrict__ frame,
                                                       int mode, int m1_srgb) {
    const uint32_t px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= (uint32_t)A.W || py >= (uint32_t)A.H) return;
    const size_t o = (size_t)py * A.W + px;
    f3 eye, dir;
    gen_ray(A.inv, (float)px, (float)py, 0.0f, 0.0f, (float)A.W, (float)A.H, eye, dir);
    float t0, t1;
    intersect_box(eye, dir, t0, t1);
    float r = A.clear[0], g = A.clear[1], b = A.clear[2];
    if (t0 < t1) {
        const float4 s = sum[o];
        if (mode == VKRT_MODE_M0) { r = A.clear[0] + s.x; g = A.clear[1] + s.y; b = A.clear[2] + s.z; }
        else {
            r = s.x; g = s.y; b = s.z;
            if (m1_srgb) { r = linear_to_srgb_naive(r); g = linear_to_srgb_naive(g); b = linear_to_srgb_naive(b); }
        }
    }
    frame[o] = pack_rgba16f(r, g, b, 1.0f);
}

// occupancy of the OWN cells of a window (cells are 8^3 voxels in global coordinates; the own brick is
// cell-aligned). M1: a cell is empty iff every voxel within its one-voxel apron (inside the window) is
// <= the transparent bound; M0: iff alpha(texel) == 0 for all its voxels.
template <int MODE, int DTYPE>
__global__ void __launch_bounds__(256) window_occupancy_kernel(const PartialArgs A, uint8_t* __restrict__ dist) {
    const uint32_t cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31u;
    if (cell >= (uint32_t)A.cnx * A.cny * A.cnz) return;
    const int lcx = (int)(cell % (uint32_t)A.cnx), lcy = (int)((cell / (uint32_t)A.cnx) % (uint32_t)A.cny), lcz = (int)(cell / ((uint32_t)A.cnx * A.cny));
    const int gx0 = (lcx + A.cox) * 8, gy0 = (lcy + A.coy) * 8, gz0 = (lcz + A.coz) * 8;
    const int ap = MODE == VKRT_MODE_M1 ? 1 : 0;
    const int x0 = max(gx0 - ap, 0), x1 = min(gx0 + 7 + ap, A.gnx - 1);
    const int y0 = max(gy0 - ap, 0), y1 = min(gy0 + 7 + ap, A.gny - 1);
    const int z0 = max(gz0 - ap, 0), z1 = min(gz0 + 7 + ap, A.gnz - 1);
    const int wyn = y1 - y0 + 1, wzn = z1 - z0 + 1;
    bool any = false;
    for (int rr = (int)lane; rr < wyn * wzn; rr += 32) {
        const int iy = y0 + rr % wyn, iz = z0 + rr / wyn;
        const size_t row = MODE == VKRT_MODE_M0 ? ((size_t)(iz - A.wz) * A.ny + (iy - A.wy)) * A.nx : part_z(A, iz - A.wz) + part_y(A, iy - A.wy);
        for (int ix = x0; ix <= x1; ++ix) {
            if (MODE == VKRT_MODE_M0) {
                const size_t i = row + (ix - A.wx);
                any = any || !m0_texel_skippable(__ldg(reinterpret_cast<const uint2*>(A.vol_a) + i), __ldg(reinterpret_cast<const uint2*>(A.vol_b) + i));
            } else {
                const float v = win_load<DTYPE>(A.vol_a, row + part_x(A, ix - A.wx)) * (DTYPE == VKRT_U8 ? 1.0f / 255.0f : 1.0f);
                any = any || !(v <= 0.0999999f);
            }
        }
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) dist[cell] = any ? 0 : 255;
}

}  // namespace

template <int MODE, int DTYPE> static cudaError_t launch_pass(const PartialArgs& A, int pass, cudaStream_t s) {
    const dim3 block(8, 8, 1), grid((unsigned)((A.W + 7) / 8), (unsigned)((A.H + 7) / 8), 1);
    if (pass == PASS_ALPHA) partial_kernel<MODE, DTYPE, PASS_ALPHA><<<grid, block, 0, s>>>(A);
    else partial_kernel<MODE, DTYPE, PASS_COLOR><<<grid, block, 0, s>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_partial(const PartialArgs& A, int mode, int dtype, int pass, cudaStream_t s) {
    if (mode == VKRT_MODE_M0) return launch_pass<VKRT_MODE_M0, 0>(A, pass, s);
    switch (dtype) {
        case VKRT_U8: return launch_pass<VKRT_MODE_M1, VKRT_U8>(A, pass, s);
        case VKRT_F16: return launch_pass<VKRT_MODE_M1, VKRT_F16>(A, pass, s);
        case VKRT_F32: return launch_pass<VKRT_MODE_M1, VKRT_F32>(A, pass, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_partial_ain(const float* T_all, size_t stride, const int* before, int n_before, float a0, float* a_in, size_t n, cudaStream_t s) {
    ain_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(T_all, stride, before, n_before, a0, a_in, n);
    return cudaGetLastError();
}

cudaError_t launch_partial_resolve(const float* T_all, size_t stride, const int* before, int n_before, float a0, float thr, float4* rgba,
                                   float* a_in_out, size_t n, cudaStream_t s) {
    resolve_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(T_all, stride, before, n_before, a0, thr, rgba, a_in_out, n);
    return cudaGetLastError();
}

cudaError_t launch_partial_finalize(const PartialArgs& A, const float4* sum, uint2* frame, int mode, int m1_srgb, cudaStream_t s) {
    const dim3 block(8, 8, 1), grid((unsigned)((A.W + 7) / 8), (unsigned)((A.H + 7) / 8), 1);
    finalize_kernel<<<grid, block, 0, s>>>(A, sum, frame, mode, m1_srgb);
    return cudaGetLastError();
}

cudaError_t launch_window_occupancy(const PartialArgs& A, int mode, int dtype, uint8_t* dist, cudaStream_t s) {
    const size_t cells = (size_t)A.cnx * A.cny * A.cnz;
    const unsigned blocks = (unsigned)((cells + 7) / 8);
    if (mode == VKRT_MODE_M0) window_occupancy_kernel<VKRT_MODE_M0, 0><<<blocks, 256, 0, s>>>(A, dist);
    else if (dtype == VKRT_U8) window_occupancy_kernel<VKRT_MODE_M1, VKRT_U8><<<blocks, 256, 0, s>>>(A, dist);
    else if (dtype == VKRT_F16) window_occupancy_kernel<VKRT_MODE_M1, VKRT_F16><<<blocks, 256, 0, s>>>(A, dist);
    else window_occupancy_kernel<VKRT_MODE_M1, VKRT_F32><<<blocks, 256, 0, s>>>(A, dist);
    return cudaGetLastError();
}

}  // namespace vkrt
