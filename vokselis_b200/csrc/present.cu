// present.cu — tonemap pass ("next" row N2): shaders/present.wgsl `fs_main` (:111-119) =
// ACESFilm (:33-35) then linear_to_srgb (:23-30), written to an Rgba8Unorm target
// (src/context/present_pipeline.rs:123-136), at 1:1 scale (present_kernel) or stretched onto a target of another size
// (present_scaled_kernel). Streaming, HBM-bound: 8 B in, 4 B out per pixel.
#include "raycast.cuh"
#include "vkrt_device.cuh"

namespace vkrt {
namespace {

__global__ void __launch_bounds__(256) present_kernel(const uint2* __restrict__ frame, uint32_t* __restrict__ rgba8, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rgba8[i] = present_pixel(__ldg(frame + i));
}

// Stretch onto a target of another size (the window differs from the backbuffer): fragment centre -> uv -> bilinear,
// clamp-to-edge `textureSample` of the rgba16f frame (present_pipeline.rs:110-118; fp32 weights, x then y), then the
// same tonemap. Every operation spelled out like present_channel.
__global__ void __launch_bounds__(256) present_scaled_kernel(const uint2* __restrict__ frame, uint32_t* __restrict__ rgba8, int W, int H, int outW,
                                                             int outH) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ox >= outW || oy >= outH) return;
    const float u = __fdiv_rn(__fadd_rn((float)ox, 0.5f), (float)outW), v = __fdiv_rn(__fadd_rn((float)oy, 0.5f), (float)outH);
    const float ux = __fsub_rn(__fmul_rn(u, (float)W), 0.5f), uy = __fsub_rn(__fmul_rn(v, (float)H), 0.5f);
    const float flx = floorf(ux), fly = floorf(uy);
    const float fx = __fsub_rn(ux, flx), fy = __fsub_rn(uy, fly);
    const int ix = (int)flx, iy = (int)fly;
    const int x0 = min(max(ix, 0), W - 1), x1 = min(max(ix + 1, 0), W - 1), y0 = min(max(iy, 0), H - 1), y1 = min(max(iy + 1, 0), H - 1);
    const float4 a = unpack_rgba16f(__ldg(frame + (size_t)y0 * W + x0)), b = unpack_rgba16f(__ldg(frame + (size_t)y0 * W + x1));
    const float4 c = unpack_rgba16f(__ldg(frame + (size_t)y1 * W + x0)), d = unpack_rgba16f(__ldg(frame + (size_t)y1 * W + x1));
    auto bil = [&](float pa, float pb, float pc, float pd) {
        const float ab = __fadd_rn(pa, __fmul_rn(fx, __fsub_rn(pb, pa))), cd = __fadd_rn(pc, __fmul_rn(fx, __fsub_rn(pd, pc)));
        return __fadd_rn(ab, __fmul_rn(fy, __fsub_rn(cd, ab)));
    };
    const float r = bil(a.x, b.x, c.x, d.x), g = bil(a.y, b.y, c.y, d.y), bl = bil(a.z, b.z, c.z, d.z), al = bil(a.w, b.w, c.w, d.w);
    rgba8[(size_t)oy * outW + ox] = unorm8(present_channel(r)) | unorm8(present_channel(g)) << 8 | unorm8(present_channel(bl)) << 16 | unorm8(al) << 24;
}

}  // namespace

cudaError_t launch_present_scaled(const uint2* frame, uint32_t* rgba8, int W, int H, int outW, int outH, cudaStream_t s) {
    const dim3 block(32, 8), grid((unsigned)((outW + 31) / 32), (unsigned)((outH + 7) / 8));
    present_scaled_kernel<<<grid, block, 0, s>>>(frame, rgba8, W, H, outW, outH);
    return cudaGetLastError();
}

cudaError_t launch_present(const uint2* frame, uint32_t* rgba8, int W, int H, cudaStream_t s) {
    const int n = W * H;
    present_kernel<<<(n + 255) / 256, 256, 0, s>>>(frame, rgba8, n);
    return cudaGetLastError();
}

}  // namespace vkrt
