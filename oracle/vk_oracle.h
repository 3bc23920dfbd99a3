/*
 * vk_oracle.h — C ABI of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a scalar fp32 restatement, in plain C++, of the reference's algorithm for the
 * raycast path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it. The product (libvokselis_rt.so) never links, loads or calls it.
 *
 * PARITY PIN STATUS: the reference (pudnax/vokselis) ships no tests, golden vectors or fixtures
 * for this path and cannot be built here (Rust nightly + wgpu + Vulkan; none present). This
 * restatement is therefore pinned against (1) oracle/_ref — the reference's own WGSL source
 * machine-translated to C++ by oracle/wgsl2cpp.py and compiled, (2) golden vectors generated from
 * that build and committed under tests/golden/, (3) analytic known answers. See DESIGN.md §3.
 *
 * Struct layouts come from include/vokselis_rt.h (same 144/48/8-byte ABI structs and VkrtParams).
 */
#ifndef VK_ORACLE_H
#define VK_ORACLE_H

#include "../include/vokselis_rt.h"

#ifdef __cplusplus
extern "C" {
#endif

#define VKO_API __attribute__((visibility("default")))

/* A host-resident volume in the upload layout (x fastest, then y, then z). */
typedef struct VkoVolume {
    int32_t nx, ny, nz;
    int32_t dtype;          /* -1: rgba16f pair (M0); else VkrtDtype scalar (M1) */
    const uint16_t* color;  /* M0: nx*ny*nz*4 halfs */
    const uint16_t* normal; /* M0: nx*ny*nz*4 halfs */
    const void* scalar;     /* M1 */
} VkoVolume;

/* Optional sub-volume restriction for sort-last rendering: only samples whose voxel-space
 * position falls in [lo, hi) on every axis contribute (global t sequence). */
typedef struct VkoBrick {
    float lo[3];
    float hi[3];
} VkoBrick;

/* Render. offsets == NULL / n_offsets == 0 -> `single` (raycast_compute.wgsl:133-137);
 * otherwise one `tile` dispatch per offset (:139-144), params->tile_size^2 threads each.
 * frame: W*H*4 halfs, row-major, top row first; tile mode writes only the pixels it covers.
 * aux (optional, W*H u32): bit31 = ray hit the box, bits 0..30 = loop iterations executed.
 * nthreads <= 0 -> all cores. */
VKO_API int vko_render(const VkoVolume* vol, const VkrtParams* params, const VkrtCameraUniform* cam,
                       int W, int H, const VkrtOffset* offsets, int n_offsets, uint16_t* frame,
                       uint32_t* aux, VkrtStats* stats, int nthreads);

/* Sort-last partial: premultiplied rgb + alpha (fp32 x4 per pixel) for the samples inside
 * `brick`, starting from (0,0,0,a_in[pixel]) — see DESIGN.md §6. */
VKO_API int vko_render_partial(const VkoVolume* vol, const VkrtParams* params, const VkrtCameraUniform* cam,
                               int W, int H, const VkoBrick* brick, const float* a_in, float* partial_rgba,
                               int nthreads);

/* shaders/xor.wgsl cs_main: which = 0 noise_volume (:55-61), 1 = bit-pattern volume (:46-53). */
VKO_API int vko_generate_xor(int n, float time, int which, uint16_t* color, uint16_t* normal, int nthreads);

/* N3: scalar grid -> (colour, normal) rgba16f pair (xor.wgsl cs_main/gradient on a sampled field). */
VKO_API int vko_scalar_to_rgba16f(const void* scalar, int dtype, int nx, int ny, int nz, uint16_t* color, uint16_t* normal);

/* shaders/present.wgsl:23-35,111-119 at 1:1 scale (no stretch): rgba16f -> rgba8 unorm. */
VKO_API int vko_present(const uint16_t* frame, int W, int H, uint8_t* rgba8);
/* the same pass stretched onto an outW x outH target (bilinear clamp-to-edge sampler, present_pipeline.rs:110-118) */
VKO_API int vko_present_scaled(const uint16_t* frame, int W, int H, int outW, int outH, uint8_t* rgba8);

/* src/camera.rs:93-113,148-171 */
VKO_API int vko_camera_uniform(float zoom, float pitch, float yaw, const float target[3], float aspect,
                               VkrtCameraUniform* out);

/* Ray generation + slab test only (raycast_compute.wgsl:102-116,42-53): writes eye(3), dir(3),
 * t0, t1 per pixel (8 floats). For known-answer tests. */
VKO_API int vko_rays(const VkrtCameraUniform* cam, int W, int H, float offx, float offy, float* out8);

/* shaders/raycast_naive.wgsl:83-125 fs_main, literal, for `count` fragments (eye, ray_dir). */
VKO_API int vko_naive_fs(const uint8_t* vol, int nx, int ny, int nz, int count, const float* eye3,
                         const float* dir3, float* out4, int nthreads);

VKO_API uint16_t vko_f32_to_f16(float f);
VKO_API float vko_f16_to_f32(uint16_t h);
VKO_API int vko_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
