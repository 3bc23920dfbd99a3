// raycast.cuh — launch interface between the C ABI (api.cu) and the kernels (raycast.cu, volume.cu, present.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/vokselis_rt.h"

namespace vkrt {

// Everything one raycast launch reads; passed by value as a __grid_constant__ kernel parameter
// (lives in the constant bank, no global loads for the camera or the parameters).
constexpr int kMaxBatch = VKRT_MAX_BATCH;
constexpr int kHullEdges = 6;
struct RenderArgs {
    // ---- what every iteration of the march reads, kept together at the front (ptxas re-reads kernel parameters with
    // LDCU inside the loop rather than keeping them in registers; tried and rejected: forcing them into registers through
    // an opaque zero costs 4-9 registers, i.e. a resident block per SM) ----
    float hx, hy, hz;  // dims / 2
    // 1.0f, as a value ptxas cannot see. ptxas contracts mul.rn.f32x2 followed by add.rn.f32x2 into one FFMA2 (even with
    // -fmad=false; the scalar .rn forms are left alone), which would round p = eye + t*dir once instead of twice. The
    // packed addition is therefore written fma(m, one, e) = RN(m * 1 + e) = RN(m + e): exact, and not fusable.
    float one;
    int nx, ny, nz;
    int nbx;              // 8^3-voxel bricks per axis of the BRICKED layout (M0)
    // Directional distance fields over the bricks, one table per ray octant (bit 0 / 1 / 2 = travelling towards +x / +y / +z),
    // back to back, dist_tab cells each: 0 = occupied, d = the d^3 bricks [b, b + s d) in the direction of travel are all
    // empty (volume.cu octant_step_kernel). Row / slab strides dsx, dsy. M0: nbx, nby, indexed after the bounds test; M1:
    // every table padded by one occupied layer on the high side of every axis (nbx + 1, nby + 1, dsz_f = nbz + 1), indexed
    // without a bounds test and clamped to dist_last (memory safety; the host only enables skipping for cameras whose ray
    // origins are near enough for the voxel index to stay within the pad, see tame_camera in api.cu). dist_bias: see api.cu.
    const uint8_t* dist;
    cudaTextureObject_t tex_d;
    int nby, nbz;
    uint32_t dsx, dsy, dist_last, dist_tab, dist_bias;
    float dsz_f;
    // the occupancy brick: edge B = 2^obs voxels (chosen per volume by api.cu occupancy_brick_shift), as floats B, B / 2, 1 / B
    int obs;
    float brick, half_brick, inv_brick;
    cudaTextureObject_t tex_a, tex_b;
    float alpha_threshold;
    float leap_r0;        // -(B / 2 + leap_eps): exit-plane offset from the brick centre is B d + leap_r0
    int leap_clip;        // M1: some dim is not a multiple of B, the partial last brick sticks out: clip regions to the grid
    float leap_eps;       // safety shrink of leap regions, in voxels
    float leap_lim[3];    // M1: dims - leap_eps (the clip)
    float dt_scale, dt_floor, initial_alpha;
    float fx, fy, fz;  // dims as f32   (textureDimensions -> vec3<f32>)
    // One launch renders n_frames cameras (`single` entry: grid.z = frame; frame f is stored at frame + f*W*H).
    // A 1080p frame with a fifth of its pixels on the box cannot fill 148 SMs, and its duration is bounded by
    // the dependent march of its longest rays; several frames of a sweep in one grid do (measured 1.85x).
    float inv[kMaxBatch][16];  // per frame: CameraUniform.inv_proj, column-major (src/camera.rs:10)
    float cull[kMaxBatch][4];  // per frame: x0, y0, x1, y1 (ray coordinates = gid + offset): pixels outside cannot hit the box
    float hull[kMaxBatch][kHullEdges][3];  // per frame: inward half-planes a*cx + b*cy + c >= 0 of the box's silhouette (api.cu cull_rect)
    int n_frames;
    int W, H;
    float aspect_hw;  // (float)H / (float)W, IEEE: raycast_compute.wgsl:105, computed once on the host
    // tiles: n_tiles == 0 -> `single`; else grid.z indexes `offsets` (device memory)
    const VkrtOffset* offsets;
    int n_tiles, tile_size;
    // volume
    const void* vol_a;  // LINEAR M0: colour texels; BRICKED M0: interleaved 16-B texels; M1: scalar grid
    const void* vol_b;  // LINEAR M0: normal texels
    // Bounding box of the occupied bricks, grown by one voxel, in the box's own coordinates ([-1,1]^3): every
    // sample outside it is a no-op. bb_lo[0] > bb_hi[0] = nothing is occupied.
    float bb_lo[3], bb_hi[3];
    // parameters (VkrtParams)
    float clear[4];
    uint2 clear_texel;  // what a miss stores: pack_rgba16f(clear.rgb, 1), packed on the host (same RN conversions)
    int m1_srgb;
    // outputs
    uint2* frame;                  // n_frames * W*H rgba16f
    uint32_t* rgba8;               // optional, n_frames * W*H: the presented pixel (ACES + sRGB, present.wgsl) from the same thread
    uint32_t* aux;                 // optional W*H: bit31 hit, low bits iterations
    unsigned long long* counters;  // optional [3]: rays_hit, samples_reference, samples_fetched
};

// One rank's view of a brick-partitioned volume (sortlast.cu): a WINDOW of the global grid.
struct PartialArgs {
    float inv[16];
    int W, H;
    int gnx, gny, gnz;             // global grid (drives ray/sample arithmetic: identical on every rank)
    float fx, fy, fz, hx, hy, hz;  // global dims as f32, and halves
    const void* vol_a;             // window array: scalar grid (M1) | colour texels (M0), x fastest
    const void* vol_b;             // M0: normal texels
    int wx, wy, wz;                // window origin in global voxel coordinates
    int nx, ny, nz;                // window dims
    int bricked;                   // scalar window stored as 8^3-voxel bricks (2 KB for fp32), x fastest inside a brick
    int bnx, bny;                  // bricks per axis of the (padded) window
    int own_lo[3], own_hi[3];      // owned voxel range (multiples of 8, or the grid edge)
    const uint8_t* dist;           // distance field over the OWN 8^3 cells
    int cox, coy, coz, cnx, cny, cnz;  // first own cell (global cell coords) and own cells per axis
    float leap_eps;
    float dt_scale, dt_floor, alpha_threshold;
    float clear[4];
    const float* a_in;   // PASS_COLOR input,  W*H
    float* T_out;        // PASS_ALPHA output, W*H
    float4* rgba_out;    // PASS_COLOR output, W*H (premultiplied rgb, alpha out)
};
cudaError_t launch_partial(const PartialArgs& A, int mode, int dtype, int pass, cudaStream_t s);
cudaError_t launch_partial_ain(const float* T_all, size_t stride, const int* before, int n_before, float a0, float* a_in, size_t n, cudaStream_t s);
cudaError_t launch_partial_resolve(const float* T_all, size_t stride, const int* before, int n_before, float a0, float thr, float4* rgba,
                                   float* a_in_out, size_t n, cudaStream_t s);
cudaError_t launch_partial_finalize(const PartialArgs& A, const float4* sum, uint2* frame, int mode, int m1_srgb, cudaStream_t s);
cudaError_t launch_window_occupancy(const PartialArgs& A, int mode, int dtype, uint8_t* dist, cudaStream_t s);

cudaError_t launch_raycast(const RenderArgs& A, int mode, int layout, int dtype, bool skip, bool dbg, cudaStream_t s);

// volume.cu
cudaError_t launch_interleave_bricked(const uint2* color, const uint2* normal, uint4* out, int nx, int ny, int nz,
                                      int nbx, int nby, int nbz, cudaStream_t s);
// dist: one byte per brick, built in two steps: occupancy (0 / 255) then the Chebyshev distance transform
// (nbx, nby, nbz = the occupancy grid: bricks of 2^bs voxels per edge)
cudaError_t launch_occupancy_m0(const uint2* color, const uint2* normal, int nx, int ny, int nz, int nbx, int nby, int nbz, int bs, uint8_t* dist,
                                cudaStream_t s);
cudaError_t launch_occupancy_m1(const void* scalar, int dtype, int nx, int ny, int nz, int nbx, int nby, int nbz, int bs,
                                uint8_t* dist, cudaStream_t s);
// d_out7: min x, y, z, max x, y, z of the occupied cells and the number of empty cells
cudaError_t launch_occupied_bounds(const uint8_t* dist, int nbx, int nby, int nbz, int* d_out7, cudaStream_t s);
cudaError_t launch_pad_dist(const uint8_t* dist, uint8_t* out, int nbx, int nby, int nbz, int tables, cudaStream_t s);
// the 8 directional distance fields of the ray octants (volume.cu), back to back
cudaError_t launch_octant_distance(const uint8_t* occ, uint8_t* oct, uint8_t* scratch, int* d_flag, int nbx, int nby, int nbz, int border,
                                   int max_d, cudaStream_t s);
cudaError_t launch_distance_transform(uint8_t* dist, uint8_t* scratch, int nbx, int nby, int nbz, int border, int max_d, cudaStream_t s);
cudaError_t launch_pregather_quads(const void* vol, int dtype, void* out_texels, int nx, int ny, int nz, cudaStream_t s);
cudaError_t launch_generate_xor(uint2* color, uint2* normal, int n, float time, int which, cudaStream_t s);

cudaError_t launch_scalar_to_rgba16f(const void* vol, int dtype, uint2* color, uint2* normal, int nx, int ny, int nz, cudaStream_t s);
cudaError_t launch_synth(void* out, int kind, int dtype, int nx, int ny, int nz, int ox, int oy, int oz, int gnx, int gny, int gnz,
                         uint32_t seed, cudaStream_t s, int bricked = 0);
cudaError_t launch_brick_window(const void* lin, void* out, int elem_bytes, int nx, int ny, int nz, cudaStream_t s);
// sort-first direct-send: up to kMaxPushDst destinations (peer memory) / frames per launch
constexpr int kMaxPushDst = 15;
// per frame of a tile push: the pixel rectangle (inclusive) that is shipped; everything outside holds the clear colour
struct PushClip {
    int x0[kMaxPushDst + 1], y0[kMaxPushDst + 1], x1[kMaxPushDst + 1], y1[kMaxPushDst + 1];
};
// vec16: every tile origin x, clip x0 and W are even (16-byte aligned rows): two pixels per store. clip == nullptr: whole tiles
cudaError_t launch_push_tiles(const uint2* src, uint2* dst, const VkrtOffset* d_offsets, int n_tiles, int tile, int W, int H, bool vec16,
                              cudaStream_t s, int n_frames = 1, const PushClip* clip = nullptr);  // n_frames consecutive frames of W*H texels on both sides
cudaError_t launch_fill_outside(uint2* frame, int W, int H, int x0, int y0, int x1, int y1, uint2 texel, cudaStream_t s);
cudaError_t launch_flush_l2(uint4* buf, size_t n16, cudaStream_t s);
// sort-last direct-send: up to kMaxPushDst destinations (peer memory) per launch
struct PushDst {
    void* ptr[kMaxPushDst];
    int n;
};
cudaError_t launch_push_many(const float* src, const PushDst& d, size_t n_floats, cudaStream_t s);  // n_floats % 4 == 0
cudaError_t launch_flag_wait_many(const PushDst& d, unsigned long long target, unsigned long long* timeouts, cudaStream_t s);
cudaError_t launch_flag_add_many(const PushDst& d, unsigned long long v, cudaStream_t s);
cudaError_t launch_flag_wait(const unsigned long long* flag, unsigned long long target, unsigned long long* timeouts, cudaStream_t s);
cudaError_t launch_flag_add(unsigned long long* flag, unsigned long long v, cudaStream_t s);
cudaError_t launch_flag_set(unsigned long long* flag, unsigned long long v, cudaStream_t s);

// present.cu
cudaError_t launch_present(const uint2* frame, uint32_t* rgba8, int W, int H, cudaStream_t s);
cudaError_t launch_present_scaled(const uint2* frame, uint32_t* rgba8, int W, int H, int outW, int outH, cudaStream_t s);

}  // namespace vkrt
