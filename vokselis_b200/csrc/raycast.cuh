// raycast.cuh — launch interface between the C ABI (api.cu) and the kernels (raycast.cu, volume.cu, present.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vokselis_rt.h"

namespace vkrt {

// Everything one raycast launch reads; passed by value as a __grid_constant__ kernel parameter
// (lives in the constant bank, no global loads for the camera or the parameters).
struct RenderArgs {
    float inv[16];  // CameraUniform.inv_proj, column-major (src/camera.rs:10)
    int W, H;
    // tiles: n_tiles == 0 -> `single`; else grid.z indexes `offsets` (device memory)
    const VkrtOffset* offsets;
    int n_tiles, tile_size;
    // volume
    const void* vol_a;  // LINEAR M0: colour texels; BRICKED M0: interleaved 16-B texels; M1: scalar grid
    const void* vol_b;  // LINEAR M0: normal texels
    cudaTextureObject_t tex_a, tex_b;
    int nx, ny, nz;
    float fx, fy, fz;  // dims as f32   (textureDimensions -> vec3<f32>)
    float hx, hy, hz;  // dims / 2
    int nbx, nby, nbz;  // 8^3-voxel bricks per axis (occupancy grid and bricked layout)
    const uint8_t* dist;  // per brick: 0 = occupied, d = bricks within Chebyshev radius d-1 are all empty
    float leap_eps;       // safety shrink of leap regions, in voxels
    // parameters (VkrtParams)
    float dt_scale, dt_floor, alpha_threshold, initial_alpha;
    float clear[4];
    int m1_srgb;
    // outputs
    uint2* frame;                  // W*H rgba16f
    uint32_t* aux;                 // optional W*H: bit31 hit, low bits iterations
    unsigned long long* counters;  // optional [3]: rays_hit, samples_reference, samples_fetched
};

cudaError_t launch_raycast(const RenderArgs& A, int mode, int layout, int dtype, bool skip, bool dbg, cudaStream_t s);

// volume.cu
cudaError_t launch_interleave_bricked(const uint2* color, const uint2* normal, uint4* out, int nx, int ny, int nz,
                                      int nbx, int nby, int nbz, cudaStream_t s);
// dist: one byte per brick, built in two steps: occupancy (0 / 255) then the Chebyshev distance transform
cudaError_t launch_occupancy_m0(const uint2* color, int nx, int ny, int nz, int nbx, int nby, int nbz, uint8_t* dist,
                                cudaStream_t s);
cudaError_t launch_occupancy_m1(const void* scalar, int dtype, int nx, int ny, int nz, int nbx, int nby, int nbz,
                                uint8_t* dist, cudaStream_t s);
cudaError_t launch_distance_transform(uint8_t* dist, uint8_t* scratch, int nbx, int nby, int nbz, int border, int max_d, cudaStream_t s);
cudaError_t launch_generate_xor(uint2* color, uint2* normal, int n, float time, int which, cudaStream_t s);

cudaError_t launch_synth(void* out, int kind, int dtype, int nx, int ny, int nz, int ox, int oy, int oz, int gnx, int gny, int gnz,
                         uint32_t seed, cudaStream_t s);
cudaError_t launch_flush_l2(uint4* buf, size_t n16, cudaStream_t s);
cudaError_t launch_flag_wait(const unsigned long long* flag, unsigned long long target, unsigned long long* timeouts, cudaStream_t s);
cudaError_t launch_flag_add(unsigned long long* flag, unsigned long long v, cudaStream_t s);
cudaError_t launch_flag_set(unsigned long long* flag, unsigned long long v, cudaStream_t s);

// present.cu
cudaError_t launch_present(const uint2* frame, uint32_t* rgba8, int W, int H, cudaStream_t s);

}  // namespace vkrt
