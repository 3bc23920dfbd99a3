// present.cu — tonemap pass ("next" row N2): shaders/present.wgsl `fs_main` (:111-119) =
// ACESFilm (:33-35) then linear_to_srgb (:23-30), written to an Rgba8Unorm target
// (src/context/present_pipeline.rs:123-136) at 1:1 scale. Streaming, HBM-bound: 8 B in, 4 B out per pixel.
#include "raycast.cuh"
#include "vkrt_device.cuh"

namespace vkrt {
namespace {

__global__ void __launch_bounds__(256) present_kernel(const uint2* __restrict__ frame, uint32_t* __restrict__ rgba8, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rgba8[i] = present_pixel(__ldg(frame + i));
}

}  // namespace

cudaError_t launch_present(const uint2* frame, uint32_t* rgba8, int W, int H, cudaStream_t s) {
    const int n = W * H;
    present_kernel<<<(n + 255) / 256, 256, 0, s>>>(frame, rgba8, n);
    return cudaGetLastError();
}

}  // namespace vkrt
