// wgsl_ref_driver.cpp — dispatch harness around the machine-translated reference shaders.
// TEST INFRASTRUCTURE (oracle/_ref); never linked into the product.
//
// The *.gen.hpp files included below are produced by oracle/wgsl2cpp.py from
// /root/reference/shaders/{raycast_compute,xor,raycast_naive,present}.wgsl at build time and live
// only in oracle/_ref/ (git-ignored). This file plays the role the reference's Rust host plays:
// it binds resources and issues dispatches exactly as the host code does —
//   raycast `single`: examples/xor/main.rs:226-233  (ceil(W/8) x ceil(H/8) groups of 8x8)
//   raycast `tile`  : examples/xor/main.rs:235-253  (per offset: 16x16 groups of 16x16)
//   xor `cs_main`   : examples/xor/xor_compute.rs:188-200 (n/8 cubed groups of 8x8x8)
//   present         : src/context/present_pipeline.rs:123-136 (full-screen triangle)
#include <cmath>
#include <cstdint>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/vokselis_rt.h"
#include "present.gen.hpp"
#include "raycast_compute.gen.hpp"
#include "raycast_naive.gen.hpp"
#include "xor.gen.hpp"

#define WREF_API extern "C" __attribute__((visibility("default")))

namespace {
int threads(int n) {
#ifdef _OPENMP
    return n > 0 ? n : omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
template <class U> void bind_uniform(U& dst, const VkrtUniform* un) {
    dst.pos = wgsl::vec3f(un->pos[0], un->pos[1], un->pos[2]);
    dst.frame = un->frame;
    dst.resolution = wgsl::vec2f(un->resolution[0], un->resolution[1]);
    dst.mouse = wgsl::vec2f(un->mouse[0], un->mouse[1]);
    dst.mouse_pressed = un->mouse_pressed;
    dst.time = un->time;
    dst.time_delta = un->time_delta;
}
template <class Cam> void bind_camera(Cam& dst, const VkrtCameraUniform* cam) {
    dst.view_pos = wgsl::vec4f(cam->view_position[0], cam->view_position[1], cam->view_position[2], cam->view_position[3]);
    for (int c = 0; c < 4; ++c) {
        dst.proj_view[c] = wgsl::vec4f(cam->proj_view[c * 4], cam->proj_view[c * 4 + 1], cam->proj_view[c * 4 + 2], cam->proj_view[c * 4 + 3]);
        dst.inv_proj[c] = wgsl::vec4f(cam->inv_proj[c * 4], cam->inv_proj[c * 4 + 1], cam->inv_proj[c * 4 + 2], cam->inv_proj[c * 4 + 3]);
    }
}
inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
}  // namespace

WREF_API int wref_num_threads(void) { return threads(0); }

WREF_API int wref_raycast_compute(int entry, const VkrtCameraUniform* cam, const VkrtUniform* un,
                                  const VkrtOffset* offsets, int n_offsets, int tile_size, const uint16_t* color,
                                  const uint16_t* normal, int nx, int ny, int nz, int W, int H, uint16_t* frame,
                                  int nthreads) {
    namespace S = wgsl_raycast_compute;
    if (!cam || !un || !color || !normal || !frame) return -1;
    bind_uniform(S::un, un);
    bind_camera(S::cam, cam);
    S::volume = wgsl::StorageTex3D{const_cast<uint16_t*>(color), nx, ny, nz};
    S::volume_normal = wgsl::StorageTex3D{const_cast<uint16_t*>(normal), nx, ny, nz};
    S::out_tex = wgsl::StorageTex2D{frame, W, H};
    const int nt = threads(nthreads);
    (void)nt;
    if (entry == 0) {
        const uint32_t wg = S::single_workgroup_size[0], hg = S::single_workgroup_size[1];
        const int gy_n = (int)(ceil_div((uint32_t)H, hg) * hg), gx_n = (int)(ceil_div((uint32_t)W, wg) * wg);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt)
        for (int y = 0; y < gy_n; ++y)
            for (int x = 0; x < gx_n; ++x) S::single(wgsl::vec3u((uint32_t)x, (uint32_t)y, 0u));
        return 0;
    }
    if (!offsets || n_offsets <= 0) return -1;
    const uint32_t wg = S::tile_workgroup_size[0], hg = S::tile_workgroup_size[1];
    const int gy_n = (int)(ceil_div((uint32_t)tile_size, hg) * hg), gx_n = (int)(ceil_div((uint32_t)tile_size, wg) * wg);
    for (int k = 0; k < n_offsets; ++k) {
        S::dyn_offset.x = offsets[k].x;
        S::dyn_offset.y = offsets[k].y;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt)
        for (int y = 0; y < gy_n; ++y)
            for (int x = 0; x < gx_n; ++x) S::tile(wgsl::vec3u((uint32_t)x, (uint32_t)y, 0u));
    }
    return 0;
}

WREF_API int wref_xor_generate(int n, const VkrtUniform* un, uint16_t* color, uint16_t* normal, int nthreads) {
    namespace S = wgsl_xor;
    if (!un || !color || !normal || n <= 0) return -1;
    bind_uniform(S::un, un);
    S::xor_tex = wgsl::StorageTex3D{color, n, n, n};
    S::normal_tex = wgsl::StorageTex3D{normal, n, n, n};
    const int nt = threads(nthreads);
    (void)nt;
    const int g = (int)(ceil_div((uint32_t)n, S::cs_main_workgroup_size[0]) * S::cs_main_workgroup_size[0]);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int z = 0; z < g; ++z)
        for (int y = 0; y < g; ++y)
            for (int x = 0; x < g; ++x) S::cs_main(wgsl::vec3u((uint32_t)x, (uint32_t)y, (uint32_t)z));
    return 0;
}

// Full-screen pass at outW x outH: uv = fragment centre / size (vs_main's interpolated uv,
// shaders/present.wgsl:98-104). Writes the `secnd` Rgba8Unorm attachment (:111-119).
WREF_API int wref_present(const uint16_t* frame, int W, int H, int outW, int outH, uint8_t* rgba8) {
    namespace S = wgsl_present;
    if (!frame || !rgba8) return -1;
    S::src_texture = wgsl::Tex2D{frame, W, H};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < outH; ++y) {
        for (int x = 0; x < outW; ++x) {
            S::VertexOutput vin;
            vin.uv = wgsl::vec2f(((float)x + 0.5f) / (float)outW, ((float)y + 0.5f) / (float)outH);
            S::FragmentOutput o = S::fs_main(vin);
            float c[4] = {o.secnd.x, o.secnd.y, o.secnd.z, o.secnd.w};
            for (int k = 0; k < 4; ++k) {
                float q = std::fmin(std::fmax(c[k], 0.0f), 1.0f) * 255.0f;
                rgba8[((size_t)y * outW + x) * 4 + k] = (uint8_t)std::nearbyint(q);
            }
        }
    }
    return 0;
}

// raycast_naive.wgsl fs_main for `count` fragments given (transformed_eye, ray_dir) per fragment —
// what vs_main hands over (:40-48). out4 = the fragment colour.
WREF_API int wref_naive_fs(const uint8_t* vol, int nx, int ny, int nz, int count, const float* eye3,
                           const float* dir3, float* out4, int nthreads) {
    namespace S = wgsl_raycast_naive;
    if (!vol || !eye3 || !dir3 || !out4) return -1;
    S::volume = wgsl::Tex3D{vol, nx, ny, nz};
    const int nt = threads(nthreads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
    for (int i = 0; i < count; ++i) {
        S::VertexOutput vin;
        vin.transformed_eye = wgsl::vec3f(eye3[i * 3], eye3[i * 3 + 1], eye3[i * 3 + 2]);
        vin.ray_dir = wgsl::vec3f(dir3[i * 3], dir3[i * 3 + 1], dir3[i * 3 + 2]);
        wgsl::vec4f c = S::fs_main(vin);
        out4[i * 4] = c.x; out4[i * 4 + 1] = c.y; out4[i * 4 + 2] = c.z; out4[i * 4 + 3] = c.w;
    }
    return 0;
}
