// present.cu — tonemap pass ("next" row N2): shaders/present.wgsl `fs_main` (:111-119) =
// ACESFilm (:33-35) then linear_to_srgb (:23-30), written to an Rgba8Unorm target
// (src/context/present_pipeline.rs:123-136) at 1:1 scale. Streaming, HBM-bound: 8 B in, 4 B out per pixel.
#include "raycast.cuh"
#include "vkrt_device.cuh"

namespace vkrt {
namespace {

__device__ __forceinline__ float present_channel(float x) {
    // ACESFilm; clamp via fminf/fmaxf so NaN -> 0 like the oracle
    float v = fminf(fmaxf((x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f), 0.0f), 1.0f);
    // linear_to_srgb: selector = ceil(v - 0.0031308) in {0,1}; mix(under, over, selector)
    const float sel = ceilf(v - 0.0031308f);
    const float under = 12.92f * v;
    const float over = 1.055f * powf(v, 0.41666f) - 0.055f;
    return under * (1.0f - sel) + over * sel;
}
__device__ __forceinline__ uint32_t unorm8(float x) { return (uint32_t)__float2int_rn(__saturatef(x) * 255.0f); }

__global__ void __launch_bounds__(256) present_kernel(const uint2* __restrict__ frame, uint32_t* __restrict__ rgba8, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = unpack_rgba16f(__ldg(frame + i));
    rgba8[i] = unorm8(present_channel(c.x)) | unorm8(present_channel(c.y)) << 8 | unorm8(present_channel(c.z)) << 16 | unorm8(c.w) << 24;
}

}  // namespace

cudaError_t launch_present(const uint2* frame, uint32_t* rgba8, int W, int H, cudaStream_t s) {
    const int n = W * H;
    present_kernel<<<(n + 255) / 256, 256, 0, s>>>(frame, rgba8, n);
    return cudaGetLastError();
}

}  // namespace vkrt
