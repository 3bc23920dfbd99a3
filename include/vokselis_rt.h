/*
 * vokselis_rt.h — C ABI of the B200-native volume raycaster (libvokselis_rt.so).
 *
 * This is the drop-in boundary for ONE path of pudnax/vokselis: the compute raycast that
 * `Xor::render` records every frame (reference: examples/xor/main.rs:210-262) and that executes
 * shaders/raycast_compute.wgsl (entry points `single` :133-137 and `tile` :139-144).
 *
 * The reference binds five bind groups and dispatches (examples/xor/raycast.rs:64-81,
 * shaders/raycast_compute.wgsl:22-33):
 *   group 0  Uniform   48-B UBO   (src/context/global_ubo.rs:52-65)       -> VkrtUniform
 *   group 1  Camera   144-B UBO   (src/camera.rs:5-11)                    -> VkrtCameraUniform
 *   group 2  volume + volume_normal, rgba16float 3-D storage textures
 *            (examples/xor/xor_compute.rs:93-118)                         -> vkrt_upload_rgba16f
 *   group 3  out_tex, rgba16float 2-D storage texture
 *            (src/context/hdr_backbuffer.rs:10-11,48-59)                  -> the context's frame
 *   group 4  Offset, 8-B read-only storage with dynamic offset
 *            (examples/xor/main.rs:20-25,77-118)                          -> VkrtOffset
 * Every entry point below names the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * VkrtStatus; nothing throws or aborts across this boundary; vkrt_last_error() returns a
 * human-readable message for the last failing call on the calling thread. Host pointers are
 * borrowed for the duration of the call. A context is bound to one CUDA device and one stream and
 * must be driven from one host thread at a time. There is NO CPU fallback: if no CUDA device is
 * usable, vkrt_create fails with VKRT_ERR_CUDA.
 */
#ifndef VOKSELIS_RT_H
#define VOKSELIS_RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VKRT_API __declspec(dllexport)
#else
#define VKRT_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------------------------ */
/* Structs with frozen layouts (byte-compatible with the reference's #[repr(C)] Pod structs).   */

/* src/camera.rs:5-11 `CameraUniform` — matrices are column-major (glam `to_cols_array_2d`).
 * NOTE `inv_proj` is inverse(proj * view), not inverse(proj) (src/camera.rs:164-170). */
typedef struct VkrtCameraUniform {
    float view_position[4];
    float proj_view[16];
    float inv_proj[16];
} VkrtCameraUniform; /* 144 bytes */

/* src/context/global_ubo.rs:52-65 `Uniform`. The raycast reads `time` and discards it
 * (shaders/raycast_compute.wgsl:100); the generator reads `time` (shaders/xor.wgsl:56). */
typedef struct VkrtUniform {
    float pos[3];
    uint32_t frame;
    float resolution[2];
    float mouse[2];
    uint32_t mouse_pressed;
    float time;
    float time_delta;
    float _padding;
} VkrtUniform; /* 48 bytes */

/* examples/xor/main.rs:20-25 `Offset` — tile origin in pixels, as two f32. */
typedef struct VkrtOffset {
    float x;
    float y;
} VkrtOffset; /* 8 bytes */

/* ------------------------------------------------------------------------------------------ */

typedef enum VkrtStatus {
    VKRT_OK = 0,
    VKRT_ERR_INVALID = -1,   /* bad argument / bad state (the reference would hit a wgpu validation error) */
    VKRT_ERR_CUDA = -2,      /* CUDA runtime failure (device lost, OOM, no device) */
    VKRT_ERR_NO_VOLUME = -3, /* render before any upload/generate */
    VKRT_ERR_UNSUPPORTED = -4
} VkrtStatus;

/* Which march body runs.
 *  M0: shaders/raycast_compute.wgsl:62-97 literally — two rgba16f volumes, nearest texel fetch,
 *      inline shading, initial alpha 0.1.
 *  M1: scalar volume + transfer function, trilinear; march body and `vertigo` transfer function
 *      of shaders/raycast_naive.wgsl:96-119 driven by M0's ray generation and box. */
typedef enum VkrtMode { VKRT_MODE_M0 = 0, VKRT_MODE_M1 = 1 } VkrtMode;

typedef enum VkrtDtype { VKRT_U8 = 0, VKRT_F16 = 1, VKRT_F32 = 2 } VkrtDtype;

/* How the volume is staged in HBM (a B200 design choice; results are identical for LINEAR and
 * BRICKED; TEXTURE uses hardware filtering: exact for M0 point fetches, 8-bit weights for M1). */
typedef enum VkrtLayout {
    VKRT_LAYOUT_LINEAR = 0,  /* as uploaded: x fastest, then y, then z */
    VKRT_LAYOUT_BRICKED = 1, /* cache-line bricks, see DESIGN.md */
    VKRT_LAYOUT_TEXTURE = 2, /* cudaArray + tex3D */
    VKRT_LAYOUT_GATHER = 3,  /* M1 only: layered cudaArray + two tld4 gathers per sample (the 8 taps in 2 texture
                                instructions), fp32 interpolation weights in the SM: exact like LINEAR */
    VKRT_LAYOUT_QUAD = 4     /* M1 only: 3-D cudaArray whose texel (x+1, y+1, z) holds the pre-gathered 2x2 xy footprint
                                v(x..x+1, y..y+1, z) (clamp-to-edge baked in); two tex3D POINT fetches per sample return
                                the 8 taps, fp32 weights in the SM: exact like LINEAR, 4x the volume's memory */
} VkrtLayout;

/* Everything the reference hard-codes as a WGSL literal or Rust const on this path. Defaults
 * (vkrt_default_params) reproduce the reference exactly. */
typedef struct VkrtParams {
    uint32_t struct_size;     /* = sizeof(VkrtParams); for ABI growth */
    int32_t mode;             /* VkrtMode */
    float dt_scale;           /* raycast_compute.wgsl:67 `dt_scale` = 1.0 */
    float dt_floor;           /* raycast_compute.wgsl:68 literal 0.01; M1 default 0 (raycast_naive.wgsl:99) */
    float alpha_threshold;    /* raycast_compute.wgsl:92 literal 0.95 */
    float initial_alpha;      /* raycast_compute.wgsl:63 literal 0.1; M1 default 0 (raycast_naive.wgsl:96) */
    float clear_color[4];     /* raycast_compute.wgsl:118 (0.023, 0.02, 0.02, 0.0) */
    int32_t tile_size;        /* examples/xor/main.rs:12 TILE_SIZE = 256 */
    int32_t layout;           /* VkrtLayout */
    int32_t skip_empty;       /* 1: occupancy-grid empty-space skipping (exact) */
    int32_t count_samples;    /* 1: accumulate VkrtStats counters (adds one atomic per block) */
    int32_t m1_srgb;          /* M1 only: apply raycast_naive.wgsl:63-68,121-123 in-shader sRGB */
    int32_t reserved[7];
} VkrtParams;

typedef struct VkrtStats {
    uint64_t rays_hit;          /* pixels whose ray entered the box (t0 < t1) */
    uint64_t samples_reference; /* loop iterations the reference shader executes (incl. early termination) */
    uint64_t samples_fetched;   /* iterations that actually fetched texels (after empty-space skipping) */
    float last_render_ms;       /* CUDA-event time of the last render call's kernels */
    float _pad;
} VkrtStats;

typedef struct VkrtContext VkrtContext; /* opaque */

/* ------------------------------------------------------------------------------------------ */
/* Context — replaces `Context::new` (src/context.rs:71-181) + `HdrBackBuffer::new`
 * (src/context/hdr_backbuffer.rs:41-87) for this path: picks the device, creates the stream and
 * the W x H rgba16f frame. The reference hard-wires 1280x720 (hdr_backbuffer.rs:11). */
VKRT_API int vkrt_create(int device, int width, int height, VkrtContext** out_ctx);
VKRT_API int vkrt_destroy(VkrtContext* ctx);
/* Replaces `Context::resize` (src/context.rs:238-249); unlike the reference it recreates the frame. */
VKRT_API int vkrt_resize(VkrtContext* ctx, int width, int height);

VKRT_API const char* vkrt_last_error(void);
VKRT_API void vkrt_default_params(int mode, VkrtParams* out);
VKRT_API int vkrt_set_params(VkrtContext* ctx, const VkrtParams* params);
VKRT_API int vkrt_get_params(VkrtContext* ctx, VkrtParams* out);

/* ------------------------------------------------------------------------------------------ */
/* Volume resources.
 * vkrt_upload_rgba16f replaces the two `Rgba16Float` 3-D textures of
 * `XorCompute::new_with_module` (examples/xor/xor_compute.rs:93-118): `color` = (r,g,b,alpha),
 * `normal` = (n.xyz, |n|), each nx*ny*nz texels of 4 halfs, x fastest then y then z. */
VKRT_API int vkrt_upload_rgba16f(VkrtContext* ctx, const uint16_t* color, const uint16_t* normal,
                                 int nx, int ny, int nz);
/* vkrt_upload_scalar replaces `VolumeTexture::new` (src/context/volume_texture.rs:32-59): a
 * tightly packed scalar grid, bytes_per_row = nx*sizeof(T), rows_per_image = ny. u8 is read as
 * unorm (R8Unorm); f16/f32 are read as-is. Sampling is linear, clamp-to-edge (:61-66). */
VKRT_API int vkrt_upload_scalar(VkrtContext* ctx, const void* data, int dtype, int nx, int ny, int nz);
/* vkrt_generate_xor replaces `XorCompute::record` (examples/xor/xor_compute.rs:188-200), i.e.
 * shaders/xor.wgsl `cs_main`, writing both rgba16f volumes on the device. `which`: 0 = live
 * `noise_volume` (:55-61), 1 = the dead bit-pattern `volume` (:46-53). Only `un->time` is read. */
VKRT_API int vkrt_generate_xor(VkrtContext* ctx, const VkrtUniform* un, int n, int which);
/* Synthetic scalar volumes of the shapes BASELINE.json names, generated on the device (no reference
 * counterpart: the reference's only dataset, bonsai_256x256x256_uint8.raw, is missing from it).
 * kind 0: hash noise box-filtered 3^3 (config 3); 1: 90 % empty 64^3 super-bricks, dense balls in the
 * rest (config 4); 2: smooth lattice noise; 3: the same as a thin fog (config 5). Deterministic in (kind, seed, voxel
 * coordinates). */
VKRT_API int vkrt_generate_synthetic(VkrtContext* ctx, int kind, int dtype, int nx, int ny, int nz, uint32_t seed);
VKRT_API int vkrt_download_scalar(VkrtContext* ctx, void* out);
/* "Next" row N3: turn the resident scalar volume into the rgba16f pair raycast_compute.wgsl consumes —
 * colour = (a/2, a/2, a/2, a), normal = normalised one-voxel backward-difference gradient of a and its
 * length: shaders/xor.wgsl cs_main (:69-78) / gradient (:63-67) on a sampled field. After it, mode M0
 * renders e.g. the bonsai scan through the compute raycaster (BASELINE configs[0]). */
VKRT_API int vkrt_scalar_to_rgba16f(VkrtContext* ctx);
/* Read the device-resident rgba16f volumes back in the upload layout (tests, screenshots). */
VKRT_API int vkrt_download_rgba16f(VkrtContext* ctx, uint16_t* color, uint16_t* normal);

/* ------------------------------------------------------------------------------------------ */
/* The hot path. vkrt_render replaces the compute pass of `Xor::render`
 * (examples/xor/main.rs:223-254): offset == NULL records `single` over the whole frame
 * (:226-233); offset != NULL records `tile` for the tile_size^2 tile at that origin (:235-253).
 * Asynchronous on the context's stream, like a wgpu queue submit. */
VKRT_API int vkrt_render(VkrtContext* ctx, const VkrtCameraUniform* cam, const VkrtUniform* un,
                         const VkrtOffset* offset);
/* All `n` tiles of an offset table in ONE launch (the reference loops 18 dispatches,
 * examples/xor/main.rs:242-253). This is also the sort-first seam: a rank passes its share. */
VKRT_API int vkrt_render_tiles(VkrtContext* ctx, const VkrtCameraUniform* cam, const VkrtUniform* un,
                               const VkrtOffset* offsets, int n);
/* The reference's tile table (examples/xor/main.rs:80-95): (w/ts+1)*(h/ts+1) origins, row-major.
 * Returns the count; writes at most `cap` entries. */
VKRT_API int vkrt_tile_table(int width, int height, int tile_size, VkrtOffset* out, int cap);
/* Host-only helper (no GPU): the screen rectangle, in the shader's pixel coordinates gid + offset
 * (raycast_compute.wgsl:102-105), outside which no ray of `render` (:99-131) can pass the slab test of
 * `intersect_box` (:42-53), plus the pixel row the box centre projects to (-1 if unknown). The raycast
 * launch uses it to build no rays for pixels that cannot hit and to issue the expensive block rows first.
 * rect = {x0, y0, x1, y1}, inclusive; +-3e38 when the projection is not trustworthy (camera plane cuts the
 * box, singular matrix): then nothing is culled. */
VKRT_API int vkrt_box_screen_bounds(const VkrtCameraUniform* cam, int width, int height, float rect[4], int* centre_row);
/* The same for the box's silhouette: up to six inward half-planes a*cx + b*cy + c >= 0 (unit normals, pixel units, two
 * pixels of margin) of the convex hull of the eight projected corners; a pixel violating one cannot hit. Unused planes
 * (and all of them when the projection is not trustworthy) are (0, 0, 1). The launch tests pixels inside the rectangle
 * against these before building a ray: about half of the rectangle lies outside the silhouette. */
VKRT_API int vkrt_box_screen_hull(const VkrtCameraUniform* cam, int width, int height, float planes[6][3]);

/* Batched `single`: the cameras of n consecutive frames of a sweep (n consecutive `Demo::render` calls of the
 * event loop, src/lib.rs:178-200; the recorder's frame sequence, src/lib.rs:132-140) rendered by ONE launch,
 * grid.z = frame. One 1080p frame with a fifth of its pixels on the box cannot fill a B200 and its duration is
 * bounded by the dependent march of its longest rays; a few frames per launch do fill it. Every frame is
 * bit-identical to what vkrt_render produces for its camera. Frames land in context-owned batch buffers
 * (vkrt_batch_frame_device_ptr / vkrt_readback_batch, valid until the next batch call). */
#define VKRT_MAX_BATCH 32
VKRT_API int vkrt_render_batch(VkrtContext* ctx, const VkrtCameraUniform* cams, int n, const VkrtUniform* un);
VKRT_API void* vkrt_batch_frame_device_ptr(VkrtContext* ctx, int i); /* W*H rgba16f, frame i of the last vkrt_render_batch */
VKRT_API int vkrt_readback_batch(VkrtContext* ctx, int i, uint16_t* rgba16f /* W*H*4 halfs */);
/* n frames (any n) for a host consumer — the screenshot/recorder path (src/context/screenshot.rs:37-77,
 * src/utils/recorder.rs:79-127) for a whole sweep: cameras in, presented RGBA8 frames out into host memory
 * (n * W*H*4 bytes, frame i at offset i*W*H*4; page-locked memory keeps the copies asynchronous). Internally
 * groups of `group` frames (<= VKRT_MAX_BATCH; 0 = default 4) go out as one launch each, and the raycast of a
 * group overlaps present + D2H of the previous one. Blocks until every frame is in `rgba8`. */
VKRT_API int vkrt_frames_host(VkrtContext* ctx, const VkrtCameraUniform* cams, int n, const VkrtUniform* un, uint8_t* rgba8, int group);

/* Present — replaces the present pass (shaders/present.wgsl:23-35,111-119;
 * src/context/present_pipeline.rs:123-136): ACES + sRGB of the frame into a W x H RGBA8 buffer. */
VKRT_API int vkrt_present(VkrtContext* ctx);
/* The same pass onto a target of another size, as when the window differs from the 1280x720 backbuffer
 * (src/context.rs:285-289 samples the backbuffer with the bilinear clamp-to-edge sampler of
 * src/context/present_pipeline.rs:110-118; README's volume.png is such a stretched capture): the presented
 * out_w x out_h RGBA8 image is copied to `rgba8` (host, out_w*out_h*4 bytes); blocks. */
VKRT_API int vkrt_present_scaled(VkrtContext* ctx, int out_w, int out_h, uint8_t* rgba8);

/* Result consumers — replace `ScreenshotCtx::capture_frame` (src/context/screenshot.rs:37-77).
 * Both block until the stream is idle. Rows are top-first, tightly packed. */
VKRT_API int vkrt_readback(VkrtContext* ctx, uint16_t* rgba16f /* W*H*4 halfs */);
VKRT_API int vkrt_readback_rgba8(VkrtContext* ctx, uint8_t* rgba8 /* W*H*4 bytes */);
/* Enqueue only (page-locked destination); the pixels are there after vkrt_sync. Lets a consumer loop keep the
 * stream busy instead of blocking once per frame. */
VKRT_API int vkrt_readback_rgba8_async(VkrtContext* ctx, uint8_t* rgba8 /* W*H*4 bytes, page-locked */);
VKRT_API int vkrt_sync(VkrtContext* ctx);
/* Validation output (no reference counterpart; the reference cannot report it): after a render with
 * params.count_samples = 1, one u32 per pixel: bit 31 = the ray entered the box (t0 < t1,
 * raycast_compute.wgsl:123), bits 0..30 = loop iterations the reference shader executes for it. */
VKRT_API int vkrt_readback_aux(VkrtContext* ctx, uint32_t* aux /* W*H */);

/* One frame end to end from HOST buffers: camera+uniform H2D, raycast, present, RGBA8 D2H into
 * the caller's `rgba8` (W*H*4 bytes); blocks until this frame's pixels are there. The _async/_wait
 * pair does the same through one of two pinned staging slots so that the D2H of frame i overlaps
 * the raycast of frame i+1 (the reference's Fifo presentation is similarly one frame deep). */
VKRT_API int vkrt_frame_host(VkrtContext* ctx, const VkrtCameraUniform* cam, const VkrtUniform* un,
                             uint8_t* rgba8);
/* Page-locked host memory for result buffers: vkrt_frame_host / vkrt_readback* DMA straight into a
 * buffer obtained here (a pageable buffer is staged through the context's own pinned slots). */
VKRT_API int vkrt_alloc_host(size_t bytes, void** out);
VKRT_API int vkrt_free_host(void* ptr);
/* Page-lock host memory the CALLER owns (e.g. a POSIX shared-memory segment mapped by several processes, one per
 * GPU) so that every GPU can DMA its frames straight into ONE consumer's memory over its own PCIe link
 * (cudaHostRegister, portable). The screenshot readback of the reference (src/context/screenshot.rs:37-77) copies
 * into a mapped buffer of its own process; this is its multi-GPU counterpart. */
VKRT_API int vkrt_host_register(void* ptr, size_t bytes);
VKRT_API int vkrt_host_unregister(void* ptr);
VKRT_API int vkrt_frame_host_async(VkrtContext* ctx, const VkrtCameraUniform* cam, const VkrtUniform* un,
                                   int slot /* 0 or 1 */);
VKRT_API int vkrt_frame_host_wait(VkrtContext* ctx, int slot, uint8_t* rgba8 /* may be NULL */);
VKRT_API const uint8_t* vkrt_frame_host_slot_ptr(VkrtContext* ctx, int slot); /* pinned staging, valid after _wait */

/* Device pointers (for zero-copy consumers, P2P and torch.distributed wrapping). */
VKRT_API void* vkrt_frame_device_ptr(VkrtContext* ctx);       /* W*H rgba16f */
VKRT_API void* vkrt_frame_rgba8_device_ptr(VkrtContext* ctx); /* W*H rgba8 */
VKRT_API void* vkrt_stream(VkrtContext* ctx);                 /* cudaStream_t */
VKRT_API int vkrt_stats(VkrtContext* ctx, VkrtStats* out);    /* synchronises */
VKRT_API int vkrt_reset_stats(VkrtContext* ctx);
/* Measurement — replaces the timestamp queries bracketing the raycast pass
 * (examples/xor/main.rs:120-131,217,258-259): vkrt_timing_enable allocates `capacity` CUDA-event
 * pairs; every vkrt_render* call then brackets its kernel with the next pair, and vkrt_timing_read
 * returns the device times (ms) of the last n renders, oldest first. vkrt_flush_l2 writes a buffer
 * larger than L2 on the context's stream (cold-cache timing hygiene; not part of the path). */
VKRT_API int vkrt_timing_enable(VkrtContext* ctx, int capacity);
VKRT_API int vkrt_timing_read(VkrtContext* ctx, float* ms, int n);
VKRT_API int vkrt_flush_l2(VkrtContext* ctx);
/* ------------------------------------------------------------------------------------------ */
/* Sort-first across the GPUs of one node, one process (one context) per GPU. The reference is
 * single-device; its `tile` entry point + Offset table (shaders/raycast_compute.wgsl:139-144,
 * examples/xor/main.rs:80-95) is the seam this uses: the volume is replicated, every participating
 * rank renders a disjoint set of tiles (or, for small frames, whole frames in turn) and its kernel
 * stores the pixels straight into a ring of frames in rank 0's memory over NVLink (CUDA IPC peer
 * mapping); arrival and slot reuse are signalled with device-side flags. No host round trip per frame.
 *   rank 0 : vkrt_sortfirst_create_root(world, slots) -> ship the handle to the other processes
 *   rank r : vkrt_sortfirst_join
 *   ranks taking part in frame f : vkrt_sortfirst_render(..., f) with their tiles (NULL/0 = whole frame)
 *   rank 0, for f = 0, 1, 2, ... in order : vkrt_sortfirst_wait(f, arrivals_target) then use the frame
 *            (vkrt_readback*, vkrt_present) then vkrt_sortfirst_consume(f) to free the ring slot.
 * arrivals_target = cumulative number of vkrt_sortfirst_render calls (over all ranks and all frames
 * so far, this one included) that target f's ring slot (slot = f % slots). */
typedef struct VkrtSortFirstHandle {
    unsigned char ipc[64];
    int32_t width, height, world, slots;
} VkrtSortFirstHandle;
VKRT_API int vkrt_sortfirst_create_root(VkrtContext* ctx, int world, int slots, VkrtSortFirstHandle* out);
VKRT_API int vkrt_sortfirst_join(VkrtContext* ctx, int rank, const VkrtSortFirstHandle* handle);
VKRT_API int vkrt_sortfirst_leave(VkrtContext* ctx);
/* This rank's share of the tile grid (tiles intersecting the frame, dealt round-robin). Returns the count. */
VKRT_API int vkrt_sortfirst_partition(int width, int height, int tile_size, int rank, int world, VkrtOffset* out, int cap);
VKRT_API int vkrt_sortfirst_render(VkrtContext* ctx, const VkrtCameraUniform* cam, const VkrtUniform* un,
                                   const VkrtOffset* offsets, int n, uint64_t frame_index);
/* `frames` granularity with batched launches: this rank renders n consecutive frames of the sweep
 * (first_frame .. first_frame+n-1, n <= VKRT_MAX_BATCH) in ONE launch into n consecutive ring slots; needs
 * slots % n == 0 and first_frame % n == 0. Each frame counts one arrival on its own slot. Peers render into a local
 * buffer and ship the group with one copy-engine transfer over NVLink (overlapping their next launch); the root
 * renders on its second stream so that its in-order waits never hold back its own launches. */
#define VKRT_SF_FLUSH_L2 1 /* flags: write an L2-sized buffer before the launch, on the stream the launch uses (benchmarks) */
VKRT_API int vkrt_sortfirst_render_batch(VkrtContext* ctx, const VkrtCameraUniform* cams, int n, const VkrtUniform* un, uint64_t first_frame,
                                         int flags);
/* "tiles" granularity, batched: this rank's tiles (offsets, n) of n_frames <= 15 consecutive frames in ONE launch
 * (grid.z = frame x tile) into n_frames consecutive ring slots; every rank calls it with its own tile table. A frame's
 * share on 1/N of the GPUs is a short launch bounded by its longest rays, so shares are batched like whole frames. */
VKRT_API int vkrt_sortfirst_render_tiles_batch(VkrtContext* ctx, const VkrtCameraUniform* cams, int n_frames, const VkrtUniform* un,
                                               const VkrtOffset* offsets, int n, uint64_t first_frame);
VKRT_API int vkrt_sortfirst_wait(VkrtContext* ctx, uint64_t frame_index, uint64_t arrivals_target);
VKRT_API int vkrt_sortfirst_consume(VkrtContext* ctx, uint64_t frame_index, int do_present);
/* Number of device-side waits that gave up after 10 s (a peer died); 0 on a healthy group. Synchronises. */
VKRT_API int vkrt_sortfirst_timeouts(VkrtContext* ctx, uint64_t* out);
/* ------------------------------------------------------------------------------------------ */
/* Sort-last for volumes larger than one GPU (BASELINE config 5). The global grid gn[3] is cut into
 * axis-aligned bricks (bounds multiples of 8); a context holds ONE brick [own_lo, own_hi) plus a
 * one-voxel halo (the "window"). Every rank walks the same global t sequence per ray and evaluates
 * the samples whose voxel index lies in its brick. A frame, with ranks in visibility order:
 *   all   : vkrt_partial_alpha(cam, d_T)            T = per-pixel transmittance of the rank's brick
 *   (T of the ranks in front must reach every rank — any transport; vokselis_b200/sortlast.py uses vkrt_exchange_*, or NCCL)
 *   all   : vkrt_partial_ain(d_T_all, ranks in front of me, n, d_ain)
 *   all   : vkrt_partial_color(cam, d_ain, d_rgba)  premultiplied rgb + alpha, early termination exact
 *   (sum d_rgba over ranks onto rank 0 — NCCL reduce)
 *   rank 0: vkrt_partial_finalize(cam, d_sum)       writes the context's rgba16f frame
 * All d_* are DEVICE pointers owned by the caller: W*H floats (T, ain), world*W*H (T_all), W*H*4 (rgba). */
VKRT_API int vkrt_upload_window(VkrtContext* ctx, const void* a, const void* b /* M0 normals or NULL */, int dtype /* -1 = rgba16f pair */,
                                const int gn[3], const int own_lo[3], const int own_hi[3]);
VKRT_API int vkrt_generate_synthetic_window(VkrtContext* ctx, int kind, int dtype, const int gn[3], const int own_lo[3],
                                            const int own_hi[3], uint32_t seed);
VKRT_API int vkrt_window_info(VkrtContext* ctx, int win_lo[3], int win_n[3]);
VKRT_API int vkrt_partial_alpha(VkrtContext* ctx, const VkrtCameraUniform* cam, const VkrtUniform* un, float* d_T);
VKRT_API int vkrt_partial_ain(VkrtContext* ctx, const float* d_T_all, const int* ranks_before, int n_before, float* d_ain);
VKRT_API int vkrt_partial_color(VkrtContext* ctx, const VkrtCameraUniform* cam, const VkrtUniform* un, const float* d_ain, float* d_rgba);
/* One-march variant (deferred early termination): vkrt_partial_relative marches the brick ONCE from alpha 0
 * and writes the relative partial + the brick's transmittance; after the all-gather, vkrt_partial_resolve
 * scales each pixel by the transmittance in front of the brick, zeroes the pixels that ended before it,
 * and flags (d_ain >= 0) the few pixels whose 0.95 crossing can fall inside the brick; a following
 * vkrt_partial_color(cam, d_ain, d_rgba) re-marches only those (it skips d_ain < 0). */
VKRT_API int vkrt_partial_relative(VkrtContext* ctx, const VkrtCameraUniform* cam, const VkrtUniform* un, float* d_rgba, float* d_T);
VKRT_API int vkrt_partial_resolve(VkrtContext* ctx, const float* d_T_all, const int* ranks_before, int n_before, float* d_rgba, float* d_ain);
VKRT_API int vkrt_partial_finalize(VkrtContext* ctx, const VkrtCameraUniform* cam, const VkrtUniform* un, const float* d_sum_rgba);

/* Direct-send exchange of the transmittance images between the ranks of a sort-last group (one process per GPU), over
 * NVLink peer memory instead of a collective: every rank owns a double-buffered table T[2][world][W*H] plus a mailbox,
 * exports it (CUDA IPC) and maps everybody else's. Per frame f, with the ranks in visibility order:
 *   vkrt_exchange_push(d_T, ranks behind me, n, f)   one kernel reads d_T once and stores it into slot [f&1][my rank] of every
 *                                                    rank that composites behind me (16-byte stores, full NVLink lines), then
 *                                                    raises their arrival counters; waits first until they have resolved f-2
 *   vkrt_exchange_wait(f, arrivals)                  device-side wait until `arrivals` images of frame f have landed here
 *   vkrt_exchange_table(f)                           the [world][W*H] table of frame f (for vkrt_partial_ain / _resolve)
 *   vkrt_exchange_done(f)                            after the resolve: the parity of f may be overwritten by frame f+2
 * Only ranks IN FRONT of a rank send to it, so a rank receives what its resolve reads and nothing else (an all-gather moves
 * every image to every rank). All calls are asynchronous on the context's stream; waits give up after 10 s
 * (vkrt_exchange_timeouts). Replaces nothing in the reference (single-device). */
typedef struct VkrtExchangeHandle {
    uint8_t ipc[64];
    int32_t width, height, world, rank;
} VkrtExchangeHandle;
VKRT_API int vkrt_exchange_create(VkrtContext* ctx, int rank, int world, VkrtExchangeHandle* out);
VKRT_API int vkrt_exchange_open(VkrtContext* ctx, const VkrtExchangeHandle* all_ranks /* [world], index = rank */);
VKRT_API int vkrt_exchange_push(VkrtContext* ctx, const float* d_T, const int* ranks_behind, int n_behind, uint64_t frame);
VKRT_API int vkrt_exchange_wait(VkrtContext* ctx, uint64_t frame, int arrivals);
VKRT_API const float* vkrt_exchange_table(VkrtContext* ctx, uint64_t frame);
VKRT_API int vkrt_exchange_done(VkrtContext* ctx, uint64_t frame);
VKRT_API int vkrt_exchange_timeouts(VkrtContext* ctx, uint64_t* out);
VKRT_API int vkrt_exchange_close(VkrtContext* ctx);

/* Eight user events on the context's stream: vkrt_mark records one, vkrt_mark_elapsed returns the
 * device time between two (ms). */
VKRT_API int vkrt_mark(VkrtContext* ctx, int idx);
VKRT_API int vkrt_mark_elapsed(VkrtContext* ctx, int from, int to, float* ms);

/* Edge, in voxels, of the bricks of the raycast's occupancy grid (exact empty-space skipping) for the volumes uploaded or
 * generated AFTER this call: 0 = automatic (the finest of 2..32 whose eight directional distance tables fit a fixed cache
 * budget), else 2, 4, 8, 16 or 32. A tuning knob: frames are bit-identical for every value. No reference counterpart (the
 * reference marches every sample, shaders/raycast_compute.wgsl:70-96). */
VKRT_API int vkrt_set_occupancy_brick(VkrtContext* ctx, int edge);
/* kind: 0 none, 1 rgba16f pair, 2 scalar. bricks = cells of the occupancy grid (see above; sort-last windows: 8^3 voxels). */
VKRT_API int vkrt_volume_info(VkrtContext* ctx, int* kind, int* dtype, int dims[3], uint64_t* bricks_total,
                              uint64_t* bricks_occupied);

/* ------------------------------------------------------------------------------------------ */
/* Host-side camera — replaces `Camera::new` + `Camera::get_proj_view_matrix`
 * (src/camera.rs:93-113,148-171; glam 0.20.5 look_at_rh / perspective_rh / inverse). */
VKRT_API int vkrt_camera_uniform(float zoom, float pitch, float yaw, const float target[3], float aspect,
                                 VkrtCameraUniform* out);
/* src/utils/mod.rs:15-18 */
VKRT_API uint32_t vkrt_dispatch_optimal(uint32_t len, uint32_t subgroup_size);

#ifdef __cplusplus
} /* extern "C" */
static_assert(sizeof(VkrtCameraUniform) == 144, "CameraUniform ABI (src/camera.rs:5-11)");
static_assert(sizeof(VkrtUniform) == 48, "Uniform ABI (src/context/global_ubo.rs:52-65)");
static_assert(sizeof(VkrtOffset) == 8, "Offset ABI (examples/xor/main.rs:20-25)");
#endif

#endif /* VOKSELIS_RT_H */
