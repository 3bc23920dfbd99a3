// api.cu — implementation of the C ABI declared in include/vokselis_rt.h.
//
// A VkrtContext plays the role of the reference's `Context` (src/context.rs:39-69) for this path:
// it owns the device, one compute stream (the wgpu Queue), the rgba16f frame (HdrBackBuffer), the
// Rgba8 capture target (present_pipeline's second attachment) and the volume resources that the
// xor example keeps in `XorCompute` (examples/xor/xor_compute.rs:10-16).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "raycast.cuh"

using namespace vkrt;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return fail(VKRT_ERR_CUDA, buf);
}
#define CK(call)                                                  \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);       \
    } while (0)

enum VolKind { VOL_NONE = 0, VOL_RGBA16F = 1, VOL_SCALAR = 2 };
// Leaps are capped at this many bricks (8 voxels each): bounds the drift of the repeated addition that a
// leap replays and the number of relaxation passes at upload.
constexpr int kMaxLeapBricks = 16;
// The raycast's occupancy grid: a distance field may see at most kMaxLeapVoxels ahead, and the brick edge is the smallest
// power of two whose eight padded tables fit kOctTableBudget bytes (and 32-bit cell indices). Finer bricks hug the occupied
// voxels more tightly — fewer samples evaluated, longer leaps in voxels (bench/leap_model.py) — and measured faster down to
// 2-voxel bricks even when the tables are 1 GB (2048^3 at edge 4: 2,059 frames/s, edge 8: 1,824, edge 16: 1,456;
// 256^3 at edge 2 / 4 / 8 / 16: 14,838 / 13,812 / 11,960 / 9,569; profiles/r02_octant_ab.md).
constexpr int kMaxLeapVoxels = 128;
constexpr size_t kOctTableBudget = (size_t)5 << 28;  // 1.25 GiB: 2048^3 at edge 4
constexpr int kMinAutoBrickShift = 1;

}  // namespace

struct VkrtContext {
    int device = 0;
    int W = 0, H = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t stream2 = nullptr;  // second render stream: sort-first tiles of consecutive frames alternate streams, so one frame's tail overlaps the next frame's head
    // Sort-first tile shares of consecutive frames rotate over kSfLanes render streams, each with its own local frame on a
    // peer: a frame's share on 1/N of the GPUs is a short launch whose duration is its longest rays (config 4 at 4K: 72 of
    // 576 tiles take 0.30 ms against 0.78 ms for all of them, bench/tiles_subset.py), so several frames must be in flight.
    static constexpr int kSfLanes = 8;
    cudaStream_t sf_stream[kSfLanes] = {};
    uint2* sf_local[kSfLanes] = {};
    cudaEvent_t sf_ready[kSfLanes] = {}, sf_copied[kSfLanes] = {};
    int sf_lane = 0;
    int sf_local_cap = 0;  // frames each sf_local buffer holds
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    // recorded on `stream` after a layout (BRICKED / TEXTURE / GATHER) was built there; a render launched on another
    // stream (the sort-first root renders on copy_stream) waits for it before reading the layout
    cudaEvent_t ev_layout = nullptr;
    bool layout_pending = false;
    cudaEvent_t ev_slot_ready[2] = {nullptr, nullptr}, ev_slot_copied[2] = {nullptr, nullptr};
    VkrtParams params{};
    // frame resources
    uint2* frame = nullptr;
    uint32_t* rgba8 = nullptr;
    uint32_t* rgba8_slot[2] = {nullptr, nullptr};  // device copies for the async host path
    // batches (vkrt_render_batch / vkrt_frames_host): two groups of up to VKRT_MAX_BATCH frames each, so the copy
    // stream can present + copy group j while the raycast of group j+1 runs
    uint2* batch_frames[2] = {nullptr, nullptr};
    uint32_t* batch_rgba8[2] = {nullptr, nullptr};
    int batch_cap = 0;  // frames per group currently allocated
    int batch_last_n = 0;
    cudaEvent_t ev_group_ready[2] = {nullptr, nullptr}, ev_group_copied[2] = {nullptr, nullptr};
    uint8_t* host_slot[2] = {nullptr, nullptr};    // pinned
    uint32_t* aux = nullptr;
    unsigned long long* counters = nullptr;
    bool timed = false;
    std::vector<cudaEvent_t> ring_begin, ring_end;  // per-render event pairs (vkrt_timing_enable)
    size_t ring_next = 0, ring_count = 0;
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    // sort-first group state (vkrt_sortfirst_*): a ring of two frames + a mailbox in rank 0's memory,
    // mapped into every peer through CUDA IPC
    int sf_rank = -1, sf_world = 0, sf_slots = 2;
    int sf_parity = 0;  // peers: which local group buffer the next vkrt_sortfirst_render_batch renders into
    cudaEvent_t marks[8] = {};
    unsigned char* sf_base = nullptr;  // [mailbox 4 KiB][frame slot 0][frame slot 1]
    bool sf_owner = false;
    uint2* own_frame = nullptr;        // the context's private frame while c->frame points into the ring
    // volume resources
    int kind = VOL_NONE, dtype = 0;
    int nx = 0, ny = 0, nz = 0, nbx = 0, nby = 0, nbz = 0;  // nb*: 8^3-voxel bricks (BRICKED layout, sort-last windows)
    // occupancy grid of the raycast: bricks of 2^obs voxels per edge (occupancy_brick_shift); obs_request: -1 = automatic
    int obs = 3, obx = 0, oby = 0, obz = 0, obs_request = -1;
    bool occ_full = false;  // every brick is occupied: there is nothing to skip
    void* lin_a = nullptr;  // rgba16f colour | scalar grid (upload layout)
    void* lin_b = nullptr;  // rgba16f normal
    uint4* bricked = nullptr;
    cudaArray_t arr_a = nullptr, arr_b = nullptr, arr_g = nullptr, arr_q = nullptr;
    cudaTextureObject_t tex_a = 0, tex_b = 0, tex_g = 0, tex_q = 0;  // tex_g: layered + gather (LAYOUT_GATHER); tex_q: pre-gathered quads (LAYOUT_QUAD)
    uint8_t* dist = nullptr;  // brick occupancy (0 = occupied, 255 = empty); sort-last windows: the Chebyshev distance field
    // the 8 directional distance fields of the ray octants, back to back (RenderArgs::dist); scalar volumes: each padded by
    // one occupied layer on the high sides
    uint8_t* dist_oct = nullptr;
    int occ_lo[3] = {0, 0, 0}, occ_hi[3] = {-1, -1, -1};  // bounding box of the occupied bricks (inclusive); hi < lo = none
    // brick-partitioned (sort-last) state: the resident volume is a window of a larger global grid
    bool windowed = false;
    bool win_bricked = false;  // scalar windows are stored as 8^3 bricks (sortlast.cu)
    int gn[3] = {0, 0, 0}, win_lo[3] = {0, 0, 0}, own_lo[3] = {0, 0, 0}, own_hi[3] = {0, 0, 0};
    int cell_lo[3] = {0, 0, 0}, cell_n[3] = {0, 0, 0};
    int* d_before = nullptr;
    // sort-last direct-send exchange (vkrt_exchange_*): own table + mailbox, and the peers' mapped through CUDA IPC
    static constexpr int kXMaxWorld = 16;
    static constexpr size_t kXMailbox = 4096;  // u64 [0], [1]: arrivals of the even / odd frames; [2]: frames resolved; [3]: timeouts
    unsigned char* x_base = nullptr;
    unsigned char* x_peer[kXMaxWorld] = {};
    int x_rank = -1, x_world = 0;
    unsigned long long x_expected[2] = {0, 0};  // cumulative arrivals expected per parity
    // tile offsets
    VkrtOffset* d_offsets = nullptr;
    int offsets_cap = 0;
    std::vector<VkrtOffset> offsets_cache;
};

namespace {

void free_layouts(VkrtContext* c) {
    if (c->tex_a) cudaDestroyTextureObject(c->tex_a);
    if (c->tex_b) cudaDestroyTextureObject(c->tex_b);
    if (c->tex_g) cudaDestroyTextureObject(c->tex_g);
    if (c->tex_q) cudaDestroyTextureObject(c->tex_q);
    if (c->arr_g) cudaFreeArray(c->arr_g);
    if (c->arr_q) cudaFreeArray(c->arr_q);
    c->tex_g = c->tex_q = 0;
    c->arr_g = c->arr_q = nullptr;
    if (c->arr_a) cudaFreeArray(c->arr_a);
    if (c->arr_b) cudaFreeArray(c->arr_b);
    if (c->bricked) cudaFree(c->bricked);
    c->tex_a = c->tex_b = 0;
    c->arr_a = c->arr_b = nullptr;
    c->bricked = nullptr;
    c->layout_pending = false;
}
void free_volume(VkrtContext* c) {
    free_layouts(c);
    if (c->lin_a) cudaFree(c->lin_a);
    if (c->lin_b) cudaFree(c->lin_b);
    if (c->dist) cudaFree(c->dist);
    if (c->dist_oct) cudaFree(c->dist_oct);
    c->lin_a = c->lin_b = nullptr;
    c->dist = nullptr;
    c->dist_oct = nullptr;
    c->kind = VOL_NONE;
    c->windowed = false;
}
// ---- sort-last direct-send exchange: helpers ----
size_t x_image_floats(const VkrtContext* c) { return (size_t)c->W * c->H; }
float* x_table(const VkrtContext* c, unsigned char* base, uint64_t frame) {
    return reinterpret_cast<float*>(base + VkrtContext::kXMailbox) + (size_t)(frame & 1u) * (size_t)c->x_world * x_image_floats(c);
}
unsigned long long* x_flag(unsigned char* base, int i) { return reinterpret_cast<unsigned long long*>(base) + i; }
void x_release(VkrtContext* c) {
    for (int r = 0; r < VkrtContext::kXMaxWorld; ++r) {
        if (c->x_peer[r] && r != c->x_rank) cudaIpcCloseMemHandle(c->x_peer[r]);
        c->x_peer[r] = nullptr;
    }
    if (c->x_base) cudaFree(c->x_base);
    c->x_base = nullptr;
    c->x_rank = -1;
    c->x_world = 0;
    c->x_expected[0] = c->x_expected[1] = 0;
}
int sf_lanes_ensure(VkrtContext* c, bool local_frames, int frames = 1) {
    if (local_frames && frames > c->sf_local_cap) {  // a lane's local buffer holds the shares of `frames` consecutive frames
        for (int i = 0; i < VkrtContext::kSfLanes; ++i) {
            if (c->sf_stream[i]) CK(cudaStreamSynchronize(c->sf_stream[i]));
            if (c->sf_local[i]) cudaFree(c->sf_local[i]);
            c->sf_local[i] = nullptr;
        }
        CK(cudaStreamSynchronize(c->copy_stream));
        c->sf_local_cap = 0;
    }
    for (int i = 0; i < VkrtContext::kSfLanes; ++i) {
        if (!c->sf_stream[i]) CK(cudaStreamCreateWithFlags(&c->sf_stream[i], cudaStreamNonBlocking));
        if (!c->sf_ready[i]) CK(cudaEventCreateWithFlags(&c->sf_ready[i], cudaEventDisableTiming));
        if (!c->sf_copied[i]) CK(cudaEventCreateWithFlags(&c->sf_copied[i], cudaEventDisableTiming));
        if (local_frames && !c->sf_local[i]) CK(cudaMalloc(&c->sf_local[i], (size_t)c->W * c->H * sizeof(uint2) * (size_t)frames));
    }
    if (local_frames && frames > c->sf_local_cap) c->sf_local_cap = frames;
    return VKRT_OK;
}
void sf_lanes_free(VkrtContext* c, bool streams_too) {
    for (int i = 0; i < VkrtContext::kSfLanes; ++i) {
        if (c->sf_stream[i]) cudaStreamSynchronize(c->sf_stream[i]);
        if (c->sf_local[i]) cudaFree(c->sf_local[i]);
        c->sf_local[i] = nullptr;
        c->sf_local_cap = 0;
        if (streams_too) {
            if (c->sf_ready[i]) cudaEventDestroy(c->sf_ready[i]);
            if (c->sf_copied[i]) cudaEventDestroy(c->sf_copied[i]);
            if (c->sf_stream[i]) cudaStreamDestroy(c->sf_stream[i]);
            c->sf_ready[i] = c->sf_copied[i] = nullptr;
            c->sf_stream[i] = nullptr;
        }
    }
    c->sf_lane = 0;
}
void sf_release(VkrtContext* c) {
    sf_lanes_free(c, false);  // the local frames have the frame's size
    if (!c->sf_base) return;
    if (c->own_frame) c->frame = c->own_frame;
    c->own_frame = nullptr;
    if (c->sf_owner) cudaFree(c->sf_base);
    else cudaIpcCloseMemHandle(c->sf_base);
    c->sf_base = nullptr;
    c->sf_rank = -1;
    c->sf_world = 0;
}
void free_frame(VkrtContext* c) {
    sf_release(c);
    if (c->frame) cudaFree(c->frame);
    if (c->rgba8) cudaFree(c->rgba8);
    if (c->aux) cudaFree(c->aux);
    for (int i = 0; i < 2; ++i) {
        if (c->rgba8_slot[i]) cudaFree(c->rgba8_slot[i]);
        if (c->host_slot[i]) cudaFreeHost(c->host_slot[i]);
        c->rgba8_slot[i] = nullptr;
        c->host_slot[i] = nullptr;
    }
    for (int i = 0; i < 2; ++i) {
        if (c->batch_frames[i]) cudaFree(c->batch_frames[i]);
        if (c->batch_rgba8[i]) cudaFree(c->batch_rgba8[i]);
        c->batch_frames[i] = nullptr;
        c->batch_rgba8[i] = nullptr;
    }
    c->batch_cap = 0;
    c->batch_last_n = 0;
    c->frame = nullptr;
    c->rgba8 = nullptr;
    c->aux = nullptr;
}
// batch frame groups, allocated on first use (2 x frames x W x H x (8 + 4) bytes)
int ensure_batch(VkrtContext* c, int frames) {
    if (frames <= c->batch_cap) return VKRT_OK;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->copy_stream));
    const size_t n = (size_t)c->W * c->H * frames;
    for (int i = 0; i < 2; ++i) {
        if (c->batch_frames[i]) cudaFree(c->batch_frames[i]);
        if (c->batch_rgba8[i]) cudaFree(c->batch_rgba8[i]);
        c->batch_frames[i] = nullptr;
        c->batch_rgba8[i] = nullptr;
        c->batch_cap = 0;
        CK(cudaMalloc(&c->batch_frames[i], n * sizeof(uint2)));
        CK(cudaMalloc(&c->batch_rgba8[i], n * 4));
    }
    c->batch_cap = frames;
    return VKRT_OK;
}
int alloc_frame(VkrtContext* c, int W, int H) {
    free_frame(c);
    const size_t n = (size_t)W * H;
    CK(cudaMalloc(&c->frame, n * sizeof(uint2)));
    CK(cudaMalloc(&c->rgba8, n * 4));
    CK(cudaMemsetAsync(c->frame, 0, n * sizeof(uint2), c->stream));
    CK(cudaMemsetAsync(c->rgba8, 0, n * 4, c->stream));
    c->W = W;
    c->H = H;
    return VKRT_OK;
}

int make_texture(cudaArray_t arr, bool linear_filter, bool border, bool normalized_u8, cudaTextureObject_t* out) {
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td{};
    const cudaTextureAddressMode am = border ? cudaAddressModeBorder : cudaAddressModeClamp;
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = am;
    td.filterMode = linear_filter ? cudaFilterModeLinear : cudaFilterModePoint;
    td.readMode = normalized_u8 ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
    td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(out, &rd, &td, nullptr));
    return VKRT_OK;
}

int copy_to_array(cudaArray_t arr, const void* src, size_t elem_bytes, int nx, int ny, int nz, cudaStream_t s) {
    cudaMemcpy3DParms p{};
    p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src), (size_t)nx * elem_bytes, (size_t)nx, (size_t)ny);
    p.dstArray = arr;
    p.extent = make_cudaExtent((size_t)nx, (size_t)ny, (size_t)nz);
    p.kind = cudaMemcpyDeviceToDevice;
    CK(cudaMemcpy3DAsync(&p, s));
    return VKRT_OK;
}

// Build whatever the selected layout needs (idempotent). Work is queued on c->stream.
int build_layout(VkrtContext* c) {
    if (c->kind == VOL_NONE) return fail(VKRT_ERR_NO_VOLUME, "no volume uploaded or generated");
    const int layout = c->params.layout;
    if (layout == VKRT_LAYOUT_LINEAR) return VKRT_OK;
    if (c->kind == VOL_RGBA16F) {
        if (layout == VKRT_LAYOUT_GATHER || layout == VKRT_LAYOUT_QUAD) return fail(VKRT_ERR_UNSUPPORTED, "layouts GATHER and QUAD are for scalar volumes (mode M1)");
        if (layout == VKRT_LAYOUT_BRICKED && !c->bricked) {
            const size_t total = (size_t)c->nbx * c->nby * c->nbz * 512;
            CK(cudaMalloc(&c->bricked, total * sizeof(uint4)));
            CK(launch_interleave_bricked((const uint2*)c->lin_a, (const uint2*)c->lin_b, c->bricked, c->nx, c->ny, c->nz, c->nbx,
                                         c->nby, c->nbz, c->stream));
        } else if (layout == VKRT_LAYOUT_TEXTURE && !c->tex_a) {
            const cudaChannelFormatDesc d = cudaCreateChannelDescHalf4();
            const cudaExtent ext = make_cudaExtent((size_t)c->nx, (size_t)c->ny, (size_t)c->nz);
            CK(cudaMalloc3DArray(&c->arr_a, &d, ext));
            CK(cudaMalloc3DArray(&c->arr_b, &d, ext));
            int rc = copy_to_array(c->arr_a, c->lin_a, 8, c->nx, c->ny, c->nz, c->stream);
            if (rc) return rc;
            rc = copy_to_array(c->arr_b, c->lin_b, 8, c->nx, c->ny, c->nz, c->stream);
            if (rc) return rc;
            rc = make_texture(c->arr_a, false, true, false, &c->tex_a);
            if (rc) return rc;
            rc = make_texture(c->arr_b, false, true, false, &c->tex_b);
            if (rc) return rc;
        }
        return VKRT_OK;
    }
    // scalar
    if (layout == VKRT_LAYOUT_TEXTURE) {
        if (!c->tex_a) {
            cudaChannelFormatDesc d;
            size_t eb;
            if (c->dtype == VKRT_U8) { d = cudaCreateChannelDesc<unsigned char>(); eb = 1; }
            else if (c->dtype == VKRT_F16) { d = cudaCreateChannelDescHalf(); eb = 2; }
            else { d = cudaCreateChannelDesc<float>(); eb = 4; }
            const cudaExtent ext = make_cudaExtent((size_t)c->nx, (size_t)c->ny, (size_t)c->nz);
            CK(cudaMalloc3DArray(&c->arr_a, &d, ext));
            int rc = copy_to_array(c->arr_a, c->lin_a, eb, c->nx, c->ny, c->nz, c->stream);
            if (rc) return rc;
            rc = make_texture(c->arr_a, true, false, c->dtype == VKRT_U8, &c->tex_a);
            if (rc) return rc;
        }
        return VKRT_OK;
    }
    if (layout == VKRT_LAYOUT_GATHER) {
        if (!c->tex_g) {
            if (c->nz > 2048) return fail(VKRT_ERR_UNSUPPORTED, "layout GATHER needs nz <= 2048 (2-D layered texture limit)");
            cudaChannelFormatDesc d;
            size_t eb;
            if (c->dtype == VKRT_U8) { d = cudaCreateChannelDesc<unsigned char>(); eb = 1; }
            else if (c->dtype == VKRT_F16) { d = cudaCreateChannelDescHalf(); eb = 2; }
            else { d = cudaCreateChannelDesc<float>(); eb = 4; }
            const cudaExtent ext = make_cudaExtent((size_t)c->nx, (size_t)c->ny, (size_t)c->nz);
            CK(cudaMalloc3DArray(&c->arr_g, &d, ext, cudaArrayLayered));  // (Layered | TextureGather is rejected by the runtime; tld4.a2d works on a plain layered array)
            int rc = copy_to_array(c->arr_g, c->lin_a, eb, c->nx, c->ny, c->nz, c->stream);
            if (rc) return rc;
            rc = make_texture(c->arr_g, true, false, c->dtype == VKRT_U8, &c->tex_g);
            if (rc) return rc;
        }
        return VKRT_OK;
    }
    if (layout == VKRT_LAYOUT_QUAD) {
        if (!c->tex_q) {
            cudaChannelFormatDesc d;
            size_t eb;
            if (c->dtype == VKRT_U8) { d = cudaCreateChannelDesc<uchar4>(); eb = 1; }
            else if (c->dtype == VKRT_F16) { d = cudaCreateChannelDescHalf4(); eb = 2; }
            else { d = cudaCreateChannelDesc<float4>(); eb = 4; }
            const int qx = c->nx + 1, qy = c->ny + 1;
            void* stage = nullptr;
            CK(cudaMalloc(&stage, (size_t)qx * qy * c->nz * 4 * eb));
            cudaError_t e = launch_pregather_quads(c->lin_a, c->dtype, stage, c->nx, c->ny, c->nz, c->stream);
            if (e == cudaSuccess) e = cudaMalloc3DArray(&c->arr_q, &d, make_cudaExtent((size_t)qx, (size_t)qy, (size_t)c->nz));
            int rc = e == cudaSuccess ? copy_to_array(c->arr_q, stage, 4 * eb, qx, qy, c->nz, c->stream) : cuda_fail(e, "LAYOUT_QUAD staging");
            if (rc == VKRT_OK && (e = cudaStreamSynchronize(c->stream)) != cudaSuccess) rc = cuda_fail(e, "LAYOUT_QUAD staging");
            cudaFree(stage);
            if (rc) return rc;
            rc = make_texture(c->arr_q, false, false, c->dtype == VKRT_U8, &c->tex_q);  // point filter, clamp-to-edge (z), unorm8 -> float in hardware
            if (rc) return rc;
        }
        return VKRT_OK;
    }
    return fail(VKRT_ERR_UNSUPPORTED, "layout BRICKED is not available for scalar volumes");
}

int ensure_layout(VkrtContext* c) {
    const void* before[4] = {c->bricked, (const void*)c->tex_a, (const void*)c->tex_g, (const void*)c->tex_q};
    const int rc = build_layout(c);
    if (rc) return rc;
    if (before[0] != c->bricked || before[1] != (const void*)c->tex_a || before[2] != (const void*)c->tex_g || before[3] != (const void*)c->tex_q) {
        CK(cudaEventRecord(c->ev_layout, c->stream));
        c->layout_pending = true;
    }
    return VKRT_OK;
}

int set_dims(VkrtContext* c, int nx, int ny, int nz) {
    if (nx <= 0 || ny <= 0 || nz <= 0 || nx > 8192 || ny > 8192 || nz > 8192) return fail(VKRT_ERR_INVALID, "volume dimensions out of range");
    c->nx = nx; c->ny = ny; c->nz = nz;
    c->nbx = (nx + 7) / 8; c->nby = (ny + 7) / 8; c->nbz = (nz + 7) / 8;
    return VKRT_OK;
}

int occupancy_brick_shift(const VkrtContext* c) {
    if (c->obs_request >= 0) return c->obs_request;
    for (int bs = kMinAutoBrickShift; bs < 5; ++bs) {
        const size_t B = (size_t)1 << bs;
        const size_t cells = ((c->nx + B - 1) / B + 1) * ((c->ny + B - 1) / B + 1) * ((c->nz + B - 1) / B + 1);
        if (8 * cells <= kOctTableBudget) return bs;
    }
    return 5;
}

int build_occupancy(VkrtContext* c) {
    c->obs = occupancy_brick_shift(c);
    const int B = 1 << c->obs;
    c->obx = (c->nx + B - 1) / B; c->oby = (c->ny + B - 1) / B; c->obz = (c->nz + B - 1) / B;
    const int max_d = std::min(255, std::max(4, kMaxLeapVoxels / B));
    const size_t cells = (size_t)c->obx * c->oby * c->obz;
    const bool scalar = c->kind == VOL_SCALAR;
    const size_t padded = (size_t)(c->obx + 1) * (c->oby + 1) * (c->obz + 1);
    uint8_t *scratch = nullptr, *oct = nullptr;
    int* d_bounds = nullptr;
    CK(cudaMalloc(&c->dist, cells));
    CK(cudaMalloc(&d_bounds, 64));
    cudaError_t e;
    if (!scalar) e = launch_occupancy_m0((const uint2*)c->lin_a, (const uint2*)c->lin_b, c->nx, c->ny, c->nz, c->obx, c->oby, c->obz, c->obs, c->dist, c->stream);
    else e = launch_occupancy_m1(c->lin_a, c->dtype, c->nx, c->ny, c->nz, c->obx, c->oby, c->obz, c->obs, c->dist, c->stream);
    // bounding box of the occupied bricks (the raycast clips every ray to it) and the number of empty ones
    int bounds[7] = {0, 0, 0, -1, -1, -1, 0};
    if (e == cudaSuccess) e = launch_occupied_bounds(c->dist, c->obx, c->oby, c->obz, d_bounds, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(bounds, d_bounds, sizeof bounds, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { cudaFree(d_bounds); return cuda_fail(e, "build_occupancy"); }
    for (int k = 0; k < 3; ++k) { c->occ_lo[k] = bounds[k]; c->occ_hi[k] = bounds[3 + k]; }
    c->occ_full = bounds[6] == 0;
    // (the kernel indexes the tables with 32 bits: a grid with more than 2^29 bricks renders without skipping; so does
    // a volume without a single empty brick — every iteration would consult the tables for nothing)
    if (c->occ_full || 8 * (scalar ? padded : cells) > 0xffffffffull) { cudaFree(d_bounds); return VKRT_OK; }
    e = cudaMalloc(&scratch, 8 * cells);
    if (e == cudaSuccess) e = cudaMalloc(&oct, 8 * cells);
    // bricks outside the grid: empty for M0 (out-of-range texels read 0), occupied for M1 (clamp-to-edge)
    if (e == cudaSuccess) e = launch_octant_distance(c->dist, oct, scratch, d_bounds, c->obx, c->oby, c->obz, scalar ? 0 : 255, max_d, c->stream);
    if (e == cudaSuccess && scalar) {
        e = cudaMalloc(&c->dist_oct, 8 * padded);
        if (e == cudaSuccess) e = launch_pad_dist(oct, c->dist_oct, c->obx, c->oby, c->obz, 8, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(scratch);
    cudaFree(d_bounds);
    if (e == cudaSuccess && !scalar) {
        c->dist_oct = oct;
        oct = nullptr;
    }
    cudaFree(oct);
    if (e != cudaSuccess) return cuda_fail(e, "build_occupancy");
    return VKRT_OK;
}

// An upload / generate that failed half-way must not leave a context that claims a resident volume with null
// arrays: drop whatever was installed, keep the error text.
int volume_guard(VkrtContext* c, int rc) {
    if (rc != VKRT_OK && c) {
        const std::string keep = g_last_error;
        cudaSetDevice(c->device);
        free_volume(c);
        g_last_error = keep;
    }
    return rc;
}

bool params_ok(const VkrtParams* p, std::string& why) {
    if (!p) { why = "params is NULL"; return false; }
    if (p->struct_size != sizeof(VkrtParams)) { why = "VkrtParams.struct_size mismatch"; return false; }
    if (p->mode != VKRT_MODE_M0 && p->mode != VKRT_MODE_M1) { why = "unknown mode"; return false; }
    if (p->layout < VKRT_LAYOUT_LINEAR || p->layout > VKRT_LAYOUT_QUAD) { why = "unknown layout"; return false; }
    if (!(p->dt_scale > 0.0f)) { why = "dt_scale must be > 0"; return false; }
    if (!(p->dt_floor >= 0.0f)) { why = "dt_floor must be >= 0"; return false; }
    if (p->tile_size <= 0 || p->tile_size > 16384) { why = "tile_size out of range"; return false; }
    return true;
}

// Screen rectangle (in the shader's pixel coordinates, gid + offset) that contains every pixel whose ray
// can hit the box [-1,1]^3, plus `margin` pixels, and the pixel row of the box centre.
//
// A pixel's ray is the world line through inv*(sx,sy,0,1) and inv*(sx,sy,1,1) (raycast_compute.wgsl:107-116);
// the slab test decides `hit` for the whole line (t0 < t1 is tested before t0 is clamped to 0). With
// M = inv^-1, a world point X lies on that line iff (MX).xy / (MX).w == (sx, sy). If all eight corners have
// (MX).w > 0 the box does not cross the plane w = 0, its image is the convex hull of the corner images, and
// the line hits the box iff (sx, sy) lies inside it; the bounding rectangle of the corner images is
// therefore conservative. Anything else (singular or non-finite matrix, a corner with w <= 0, i.e. the
// camera plane cuts the box) disables culling: the rectangle becomes the whole plane. The kernel's own
// rounding moves the silhouette by ~1e-6 of the frame; the margin is two pixels.
// Are the ray origins of this camera close enough to the box for a sample's voxel index to stay within one brick of the
// grid? The kernel computes p = eye + t*dir in fp32: with |eye| <= 1024 a position inside the box [-1,1]^3 is off by at most
// a few ulp(1024) ~ 1e-4, i.e. <= 1 voxel at 8192^3, so the index lies in [-1, N] and the padded distance field covers it.
// Origins are eye(sx,sy) = (inv * (sx,sy,0,1)).xyz / .w, a ratio of affine functions of the screen point: if w keeps its sign
// at the four frame corners it keeps it inside, and every origin lies in the convex hull of the corner origins. Anything
// else (far, orthographic-at-infinity, non-finite) renders with skipping off: exact by construction, only slower.
bool tame_camera(const float inv[16], int W, int H) {
    double wmin = 1e300, wmax = 0.0;
    int sign = 0;
    for (int k = 0; k < 4; ++k) {
        const double cx = (k & 1) ? (double)W + 1.0 : -1.0, cy = (k & 2) ? (double)H + 1.0 : -1.0;
        const double sx = 2.0 * cx / W - 1.0, sy = (2.0 * cy / H - 1.0) * (-(double)H / (double)W);
        double v[4];
        for (int r = 0; r < 4; ++r) v[r] = (double)inv[r] * sx + (double)inv[4 + r] * sy + (double)inv[12 + r];
        if (!std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2]) || !std::isfinite(v[3]) || v[3] == 0.0) return false;
        const int sg = v[3] > 0.0 ? 1 : -1;
        if (sign != 0 && sg != sign) return false;
        sign = sg;
        wmin = std::min(wmin, fabs(v[3])); wmax = std::max(wmax, fabs(v[3]));
        for (int r = 0; r < 3; ++r) if (!(fabs(v[r] / v[3]) <= 1024.0)) return false;
    }
    return wmin > 1e-6 * wmax;
}

// hull (optional): up to kHullEdges inward half-planes a*cx + b*cy + c >= 0 (unit normals, pixel units, 2 px of margin)
// of the convex hull of the eight projected corners — the box's silhouette; unused edges are (0, 0, 1). A pixel's ray is a
// line through the eye, and it meets the (convex) box iff the pixel lies in the convex hull of the corner images, so
// everything outside the hull misses as surely as everything outside the rectangle, and the hull is ~half its area.
void cull_rect(const float inv[16], int W, int H, float cull[4], int* centre_row, float (*hull)[3] = nullptr) {
    const float big = 3.0e38f;
    cull[0] = cull[1] = -big; cull[2] = cull[3] = big;
    *centre_row = -1;
    if (hull) for (int k = 0; k < kHullEdges; ++k) { hull[k][0] = 0.0f; hull[k][1] = 0.0f; hull[k][2] = 1.0f; }
    double px[8], py[8];
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int k = 0; k < 4; ++k) { a[r][k] = (double)inv[4 * k + r]; a[r][4 + k] = r == k ? 1.0 : 0.0; }  // column-major in
    for (int r = 0; r < 4; ++r)
        for (int k = 0; k < 4; ++k) if (!std::isfinite(a[r][k])) return;
    for (int i = 0; i < 4; ++i) {  // Gauss-Jordan, partial pivoting
        int piv = i;
        for (int r = i + 1; r < 4; ++r) if (fabs(a[r][i]) > fabs(a[piv][i])) piv = r;
        if (!(fabs(a[piv][i]) > 1e-30)) return;
        if (piv != i) for (int k = 0; k < 8; ++k) std::swap(a[i][k], a[piv][k]);
        const double d = 1.0 / a[i][i];
        for (int k = 0; k < 8; ++k) a[i][k] *= d;
        for (int r = 0; r < 4; ++r) if (r != i) { const double f = a[r][i]; if (f != 0.0) for (int k = 0; k < 8; ++k) a[r][k] -= f * a[i][k]; }
    }
    // M = a[:, 4:8]. Its overall sign is arbitrary (inv is used projectively): orient it so the box centre has w > 0.
    double sgn = a[3][7] >= 0.0 ? 1.0 : -1.0;
    double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300, wmax = 0.0, wmin = 1e300;
    for (int cidx = 0; cidx < 8; ++cidx) {
        const double X[4] = {cidx & 1 ? 1.0 : -1.0, cidx & 2 ? 1.0 : -1.0, cidx & 4 ? 1.0 : -1.0, 1.0};
        double h[4];
        for (int r = 0; r < 4; ++r) h[r] = sgn * (a[r][4] * X[0] + a[r][5] * X[1] + a[r][6] * X[2] + a[r][7] * X[3]);
        if (!std::isfinite(h[0]) || !std::isfinite(h[1]) || !std::isfinite(h[3])) return;
        wmax = std::max(wmax, h[3]); wmin = std::min(wmin, h[3]);
        if (!(h[3] > 0.0)) return;
        const double sx = h[0] / h[3], sy = h[1] / h[3];
        // sx = 2*cx/W - 1 ; sy = (2*cy/H - 1) * (-H/W)     (raycast_compute.wgsl:103-105)
        const double cx = (sx + 1.0) * 0.5 * W, cy = (1.0 - sy * (double)W / (double)H) * 0.5 * H;
        x0 = std::min(x0, cx); x1 = std::max(x1, cx); y0 = std::min(y0, cy); y1 = std::max(y1, cy);
        px[cidx] = cx; py[cidx] = cy;
    }
    if (!(wmin > 1e-6 * wmax)) return;  // a corner (almost) on the camera plane: its image is unreliable
    if (!std::isfinite(x0) || !std::isfinite(x1) || !std::isfinite(y0) || !std::isfinite(y1)) return;
    const double margin = 2.0;
    const double lim = 1.0e9;
    cull[0] = (float)std::max(-lim, floor(x0 - margin)); cull[1] = (float)std::max(-lim, floor(y0 - margin));
    cull[2] = (float)std::min(lim, ceil(x1 + margin));   cull[3] = (float)std::min(lim, ceil(y1 + margin));
    const double cyc = 0.5 * (std::max(y0, 0.0) + std::min(y1, (double)H - 1.0));
    if (cyc >= 0.0 && cyc <= (double)H - 1.0) *centre_row = (int)cyc;
    if (hull) {
        // Andrew's monotone chain over the eight corner images; collinear points dropped. A box's silhouette has at most
        // six edges; anything else (degenerate, numerically odd) leaves the hull unused: the rectangle still culls.
        int idx[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        std::sort(idx, idx + 8, [&](int a, int b) { return px[a] < px[b] || (px[a] == px[b] && py[a] < py[b]); });
        auto cross = [&](int o, int a, int b) { return (px[a] - px[o]) * (py[b] - py[o]) - (py[a] - py[o]) * (px[b] - px[o]); };
        int hv[16], m = 0;
        for (int i = 0; i < 8; ++i) { while (m >= 2 && cross(hv[m - 2], hv[m - 1], idx[i]) <= 0.0) --m; hv[m++] = idx[i]; }
        for (int i = 6, lo = m + 1; i >= 0; --i) { while (m >= lo && cross(hv[m - 2], hv[m - 1], idx[i]) <= 0.0) --m; hv[m++] = idx[i]; }
        --m;  // the last point repeats the first
        if (m >= 3 && m <= kHullEdges) {
            bool ok = true;
            float planes[kHullEdges][3];
            for (int e = 0; e < m && ok; ++e) {
                const int a = hv[e], b = hv[(e + 1) % m];  // counter-clockwise: the interior is on the left of a -> b
                const double ex = px[b] - px[a], ey = py[b] - py[a], len = sqrt(ex * ex + ey * ey);
                if (!(len > 1e-9) || !std::isfinite(len)) { ok = false; break; }
                const double nx = -ey / len, ny = ex / len;  // inward normal
                const double cc = -(nx * px[a] + ny * py[a]) + margin;
                if (!std::isfinite(nx) || !std::isfinite(ny) || !std::isfinite(cc) || fabs(cc) > 1e9) { ok = false; break; }
                planes[e][0] = (float)nx; planes[e][1] = (float)ny; planes[e][2] = (float)cc;
            }
            if (ok) for (int e = 0; e < m; ++e) { hull[e][0] = planes[e][0]; hull[e][1] = planes[e][1]; hull[e][2] = planes[e][2]; }
        }
    }
}

// n_frames > 1 (with frames_out): the `single` entry for n_frames cameras in ONE launch, frame f stored at
// frames_out + f*W*H (vkrt_render_batch / vkrt_frames_host); otherwise one camera into the context's frame.
// rgba8_out: also store the presented RGBA8 pixels (fused present pass), n_frames * W*H.
int do_render(VkrtContext* c, const VkrtCameraUniform* cam, const VkrtUniform* un, const VkrtOffset* offsets, int n, bool bracket = true,
              int n_frames = 1, uint2* frames_out = nullptr, uint32_t* rgba8_out = nullptr, cudaStream_t on = nullptr) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    if (!cam || !un) return fail(VKRT_ERR_INVALID, "camera/uniform is NULL");
    if (n_frames < 1 || n_frames > kMaxBatch) return fail(VKRT_ERR_INVALID, "batch size out of range");
    CK(cudaSetDevice(c->device));
    if (c->kind == VOL_NONE) return fail(VKRT_ERR_NO_VOLUME, "render before any volume upload/generate");
    if (c->windowed) return fail(VKRT_ERR_INVALID, "the resident volume is one brick of a partitioned grid: use vkrt_partial_*");
    const VkrtParams& P = c->params;
    if (P.mode == VKRT_MODE_M0 && c->kind != VOL_RGBA16F) return fail(VKRT_ERR_INVALID, "mode M0 needs an rgba16f volume pair (vkrt_upload_rgba16f / vkrt_generate_xor)");
    if (P.mode == VKRT_MODE_M1 && c->kind != VOL_SCALAR) return fail(VKRT_ERR_INVALID, "mode M1 needs a scalar volume (vkrt_upload_scalar)");
    int rc = ensure_layout(c);
    if (rc) return rc;
    // a layout built just now was queued on c->stream: a launch on another stream must not overtake it
    if (on && on != c->stream && c->layout_pending) CK(cudaStreamWaitEvent(on, c->ev_layout, 0));
    // Skipping is exact only when a transparent sample is a bit-exact no-op: for M0 that requires
    // clear_color.a == 0 (raycast_compute.wgsl:89,91). And the reference tests `a >= threshold` only after
    // compositing a sample (:92), so with initial_alpha >= alpha_threshold it stops after its FIRST sample, which a
    // leap would pass over. Otherwise fall back to the full march.
    bool skip = P.skip_empty && c->dist_oct && !(P.mode == VKRT_MODE_M0 && P.clear_color[3] != 0.0f) && P.initial_alpha < P.alpha_threshold;

    RenderArgs A{};
    A.W = c->W; A.H = c->H;
    A.aspect_hw = (float)c->H / (float)c->W;
    A.n_frames = n_frames;
    for (int f = 0; f < n_frames; ++f) {
        memcpy(A.inv[f], cam[f].inv_proj, sizeof A.inv[f]);
        int centre_row;
        cull_rect(A.inv[f], A.W, A.H, A.cull[f], &centre_row, A.hull[f]);
        // M1's skipping indexes the padded distance field without a bounds test (RenderArgs::dist)
        if (P.mode == VKRT_MODE_M1 && skip && !tame_camera(A.inv[f], A.W, A.H)) skip = false;
    }
    A.n_tiles = 0; A.tile_size = P.tile_size; A.offsets = nullptr;
    if (offsets && n > 0) {
        if (n > c->offsets_cap) {
            if (c->d_offsets) cudaFree(c->d_offsets);
            c->d_offsets = nullptr;
            CK(cudaMalloc(&c->d_offsets, (size_t)n * sizeof(VkrtOffset)));
            c->offsets_cap = n;
            c->offsets_cache.clear();
        }
        if ((int)c->offsets_cache.size() != n || memcmp(c->offsets_cache.data(), offsets, (size_t)n * sizeof(VkrtOffset)) != 0) {
            c->offsets_cache.assign(offsets, offsets + n);
            CK(cudaStreamSynchronize(c->copy_stream));  // a sort-first push on the copy stream may still be reading the old table
            CK(cudaStreamSynchronize(c->stream2));
            for (cudaStream_t ls : c->sf_stream) if (ls) CK(cudaStreamSynchronize(ls));  // ... or a render on one of the lanes
            CK(cudaMemcpyAsync(c->d_offsets, c->offsets_cache.data(), (size_t)n * sizeof(VkrtOffset), cudaMemcpyHostToDevice, c->stream));
            // a launch on another stream (the sort-first root renders on the copy stream) must see the new table
            CK(cudaStreamSynchronize(c->stream));
        }
        A.offsets = c->d_offsets;
        A.n_tiles = n;
    }
    const int layout = P.layout;
    if (c->kind == VOL_RGBA16F) {
        A.vol_a = layout == VKRT_LAYOUT_BRICKED ? (const void*)c->bricked : c->lin_a;
        A.vol_b = c->lin_b;
    } else {
        A.vol_a = c->lin_a;
    }
    A.tex_a = layout == VKRT_LAYOUT_GATHER ? c->tex_g : (layout == VKRT_LAYOUT_QUAD ? c->tex_q : c->tex_a); A.tex_b = c->tex_b;
    A.nx = c->nx; A.ny = c->ny; A.nz = c->nz;
    A.fx = (float)c->nx; A.fy = (float)c->ny; A.fz = (float)c->nz;
    A.hx = A.fx / 2.0f; A.hy = A.fy / 2.0f; A.hz = A.fz / 2.0f;
    A.one = 1.0f;
    A.nbx = c->nbx; A.nby = c->nby; A.nbz = c->nbz;
    A.dist = c->dist_oct;
    if (c->kind == VOL_SCALAR) {  // padded tables (M1)
        A.dsx = (uint32_t)c->obx + 1u; A.dsy = (uint32_t)c->oby + 1u;
        A.dist_tab = (uint32_t)((size_t)(c->obx + 1) * (c->oby + 1) * (c->obz + 1));
        A.dsz_f = (float)(c->obz + 1);
    } else {
        A.dsx = (uint32_t)c->obx; A.dsy = (uint32_t)c->oby;
        A.dist_tab = (uint32_t)((size_t)c->obx * c->oby * c->obz);
        A.dsz_f = (float)c->obz;
    }
    A.obs = c->obs;
    A.brick = (float)(1 << c->obs); A.half_brick = 0.5f * A.brick; A.inv_brick = 1.0f / A.brick;
    A.dist_last = 8u * A.dist_tab - 1u;
    // M1 takes the brick coordinates as the bit patterns of 1.5 * 2^23 + b (raycast.cu): their common offset, mod 2^32
    A.dist_bias = 0x4B400000u * (A.dsy * A.dsx + A.dsx + 1u);
    // shrink leap regions by ~16 ulp of the largest voxel coordinate (rounding of p and q)
    A.leap_eps = 16.0f * 1.1920929e-07f * (float)(c->nx > c->ny ? (c->nx > c->nz ? c->nx : c->nz) : (c->ny > c->nz ? c->ny : c->nz));
    {
        const int n3[3] = {c->nx, c->ny, c->nz};
        for (int k = 0; k < 3; ++k) {
            if (c->occ_hi[0] < c->occ_lo[0]) { A.bb_lo[k] = 1.0f; A.bb_hi[k] = -1.0f; continue; }
            // voxel q = (p + 1) * n/2  ->  p = 2q/n - 1; one voxel of margin (float rounding is ~1e-4 of that)
            const double lo = (double)c->occ_lo[k] * A.brick - 1.0, hi = std::min((double)(c->occ_hi[k] + 1) * A.brick, (double)n3[k]) + 1.0;
            A.bb_lo[k] = (float)(2.0 * lo / n3[k] - 1.0);
            A.bb_hi[k] = (float)(2.0 * hi / n3[k] - 1.0);
        }
    }
    A.leap_r0 = -(A.half_brick + A.leap_eps);
    A.leap_clip = ((c->nx | c->ny | c->nz) & ((1 << c->obs) - 1)) != 0;
    A.leap_lim[0] = A.fx - A.leap_eps; A.leap_lim[1] = A.fy - A.leap_eps; A.leap_lim[2] = A.fz - A.leap_eps;
    A.dt_scale = P.dt_scale; A.dt_floor = P.dt_floor; A.alpha_threshold = P.alpha_threshold; A.initial_alpha = P.initial_alpha;
    memcpy(A.clear, P.clear_color, sizeof A.clear);
    {
        const __half2 lo = __floats2half2_rn(A.clear[0], A.clear[1]), hi = __floats2half2_rn(A.clear[2], 1.0f);
        memcpy(&A.clear_texel.x, &lo, 4);
        memcpy(&A.clear_texel.y, &hi, 4);
    }
    A.m1_srgb = P.m1_srgb;
    A.frame = frames_out ? frames_out : c->frame;
    A.rgba8 = rgba8_out;
    const bool dbg = P.count_samples != 0;
    if (dbg && n_frames > 1) return fail(VKRT_ERR_INVALID, "params.count_samples needs single-frame renders (the counters are per frame)");
    if (n_frames > 1 && (offsets && n > 0) && (size_t)n * n_frames > 65535) return fail(VKRT_ERR_INVALID, "tiles x frames exceed one launch (65535)");
    if (dbg && !c->aux) {
        CK(cudaMalloc(&c->aux, (size_t)c->W * c->H * 4));
        CK(cudaMemsetAsync(c->aux, 0, (size_t)c->W * c->H * 4, c->stream));
    }
    A.aux = dbg ? c->aux : nullptr;
    A.counters = dbg ? c->counters : nullptr;
    if (!bracket) {
        CK(launch_raycast(A, P.mode, layout, c->dtype, skip, dbg, on ? on : c->stream));
        return VKRT_OK;
    }
    cudaEvent_t eb = c->ev_begin, ee = c->ev_end;
    if (!c->ring_begin.empty()) {
        eb = c->ring_begin[c->ring_next];
        ee = c->ring_end[c->ring_next];
        c->ring_next = (c->ring_next + 1) % c->ring_begin.size();
        if (c->ring_count < c->ring_begin.size()) ++c->ring_count;
    }
    CK(cudaEventRecord(eb, c->stream));
    CK(launch_raycast(A, P.mode, layout, c->dtype, skip, dbg, c->stream));
    CK(cudaEventRecord(ee, c->stream));
    if (c->ring_begin.empty()) c->timed = true;
    return VKRT_OK;
}

}  // namespace

extern "C" {

const char* vkrt_last_error(void) { return g_last_error.c_str(); }

void vkrt_default_params(int mode, VkrtParams* out) {
    if (!out) return;
    memset(out, 0, sizeof *out);
    out->struct_size = (uint32_t)sizeof(VkrtParams);
    out->mode = mode;
    out->dt_scale = 1.0f;          // raycast_compute.wgsl:67 / raycast_naive.wgsl:98
    out->alpha_threshold = 0.95f;  // raycast_compute.wgsl:92 / raycast_naive.wgsl:115
    out->tile_size = 256;          // examples/xor/main.rs:12
    out->layout = VKRT_LAYOUT_LINEAR;
    if (mode == VKRT_MODE_M0) {
        out->dt_floor = 0.01f;       // raycast_compute.wgsl:68
        out->initial_alpha = 0.1f;   // raycast_compute.wgsl:63
        out->clear_color[0] = 0.023f; out->clear_color[1] = 0.02f; out->clear_color[2] = 0.02f; out->clear_color[3] = 0.0f;  // :118
    }  // M1: dt_floor 0, initial colour 0, misses black (raycast_naive.wgsl:91,96,99)
}

int vkrt_create(int device, int width, int height, VkrtContext** out_ctx) {
    if (!out_ctx) return fail(VKRT_ERR_INVALID, "out_ctx is NULL");
    *out_ctx = nullptr;
    if (width <= 0 || height <= 0 || width > 32768 || height > 32768) return fail(VKRT_ERR_INVALID, "frame size out of range");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(VKRT_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(VKRT_ERR_INVALID, "device index out of range");
    CK(cudaSetDevice(device));
    VkrtContext* c = new (std::nothrow) VkrtContext();
    if (!c) return fail(VKRT_ERR_INVALID, "out of host memory");
    c->device = device;
    vkrt_default_params(VKRT_MODE_M0, &c->params);
    int rc = VKRT_OK;
    do {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaEventCreate(&c->ev_begin)) != cudaSuccess) break;
        if ((e = cudaEventCreate(&c->ev_end)) != cudaSuccess) break;
        if ((e = cudaEventCreateWithFlags(&c->ev_layout, cudaEventDisableTiming)) != cudaSuccess) break;
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
            e = cudaEventCreateWithFlags(&c->ev_slot_ready[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_slot_copied[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_group_ready[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_group_copied[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) break;
        if ((e = cudaMalloc(&c->counters, 3 * sizeof(unsigned long long))) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(c->counters, 0, 3 * sizeof(unsigned long long), c->stream)) != cudaSuccess) break;
        rc = alloc_frame(c, width, height);
    } while (0);
    if (e != cudaSuccess) rc = cuda_fail(e, "vkrt_create");
    if (rc != VKRT_OK) {
        std::string keep = g_last_error;
        vkrt_destroy(c);
        g_last_error = keep;
        return rc;
    }
    *out_ctx = c;
    return VKRT_OK;
}

int vkrt_destroy(VkrtContext* c) {
    if (!c) return VKRT_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    free_volume(c);
    free_frame(c);
    sf_lanes_free(c, true);
    x_release(c);
    if (c->counters) cudaFree(c->counters);
    if (c->d_offsets) cudaFree(c->d_offsets);
    if (c->d_before) cudaFree(c->d_before);
    if (c->flush_buf) cudaFree(c->flush_buf);
    for (cudaEvent_t e : c->ring_begin) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ring_end) cudaEventDestroy(e);
    for (cudaEvent_t e : c->marks) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_slot_ready[i]) cudaEventDestroy(c->ev_slot_ready[i]);
        if (c->ev_slot_copied[i]) cudaEventDestroy(c->ev_slot_copied[i]);
        if (c->ev_group_ready[i]) cudaEventDestroy(c->ev_group_ready[i]);
        if (c->ev_group_copied[i]) cudaEventDestroy(c->ev_group_copied[i]);
    }
    if (c->ev_layout) cudaEventDestroy(c->ev_layout);
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_end) cudaEventDestroy(c->ev_end);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    delete c;
    return VKRT_OK;
}

int vkrt_resize(VkrtContext* c, int width, int height) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    if (width <= 0 || height <= 0 || width > 32768 || height > 32768) return fail(VKRT_ERR_INVALID, "frame size out of range");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->copy_stream));
    return alloc_frame(c, width, height);
}

int vkrt_set_params(VkrtContext* c, const VkrtParams* p) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    std::string why;
    if (!params_ok(p, why)) return fail(VKRT_ERR_INVALID, why);
    c->params = *p;
    return VKRT_OK;
}

int vkrt_get_params(VkrtContext* c, VkrtParams* out) {
    if (!c || !out) return fail(VKRT_ERR_INVALID, "NULL argument");
    *out = c->params;
    return VKRT_OK;
}

static int upload_rgba16f_impl(VkrtContext* c, const uint16_t* color, const uint16_t* normal, int nx, int ny, int nz) {
    if (!c || !color || !normal) return fail(VKRT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    free_volume(c);
    int rc = set_dims(c, nx, ny, nz);
    if (rc) return rc;
    const size_t bytes = (size_t)nx * ny * nz * 8;
    CK(cudaMalloc(&c->lin_a, bytes));
    CK(cudaMalloc(&c->lin_b, bytes));
    CK(cudaMemcpyAsync(c->lin_a, color, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->lin_b, normal, bytes, cudaMemcpyHostToDevice, c->stream));
    c->kind = VOL_RGBA16F;
    rc = build_occupancy(c);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));  // host pointers are borrowed for the call only
    return VKRT_OK;
}

int vkrt_upload_rgba16f(VkrtContext* c, const uint16_t* color, const uint16_t* normal, int nx, int ny, int nz) { return volume_guard(c, upload_rgba16f_impl(c, color, normal, nx, ny, nz)); }

static int upload_scalar_impl(VkrtContext* c, const void* data, int dtype, int nx, int ny, int nz) {
    if (!c || !data) return fail(VKRT_ERR_INVALID, "NULL argument");
    if (dtype < VKRT_U8 || dtype > VKRT_F32) return fail(VKRT_ERR_INVALID, "unknown dtype");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    free_volume(c);
    int rc = set_dims(c, nx, ny, nz);
    if (rc) return rc;
    const size_t eb = dtype == VKRT_U8 ? 1 : (dtype == VKRT_F16 ? 2 : 4);
    const size_t bytes = (size_t)nx * ny * nz * eb;
    CK(cudaMalloc(&c->lin_a, bytes));
    CK(cudaMemcpyAsync(c->lin_a, data, bytes, cudaMemcpyHostToDevice, c->stream));
    c->kind = VOL_SCALAR;
    c->dtype = dtype;
    rc = build_occupancy(c);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return VKRT_OK;
}

int vkrt_upload_scalar(VkrtContext* c, const void* data, int dtype, int nx, int ny, int nz) { return volume_guard(c, upload_scalar_impl(c, data, dtype, nx, ny, nz)); }

static int generate_xor_impl(VkrtContext* c, const VkrtUniform* un, int n, int which) {
    if (!c || !un) return fail(VKRT_ERR_INVALID, "NULL argument");
    if (which < 0 || which > 1) return fail(VKRT_ERR_INVALID, "which must be 0 (noise_volume) or 1 (volume)");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    free_volume(c);
    int rc = set_dims(c, n, n, n);
    if (rc) return rc;
    const size_t bytes = (size_t)n * n * n * 8;
    CK(cudaMalloc(&c->lin_a, bytes));
    CK(cudaMalloc(&c->lin_b, bytes));
    CK(launch_generate_xor((uint2*)c->lin_a, (uint2*)c->lin_b, n, un->time, which, c->stream));
    c->kind = VOL_RGBA16F;
    return build_occupancy(c);
}

int vkrt_generate_xor(VkrtContext* c, const VkrtUniform* un, int n, int which) { return volume_guard(c, generate_xor_impl(c, un, n, which)); }

static int generate_synthetic_impl(VkrtContext* c, int kind, int dtype, int nx, int ny, int nz, uint32_t seed) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    if (kind < 0 || kind > 3) return fail(VKRT_ERR_INVALID, "kind must be 0 (noise fog), 1 (sparse blobs), 2 (smooth lattice) or 3 (thin smooth fog)");
    if (dtype < VKRT_U8 || dtype > VKRT_F32) return fail(VKRT_ERR_INVALID, "unknown dtype");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    free_volume(c);
    int rc = set_dims(c, nx, ny, nz);
    if (rc) return rc;
    const size_t eb = dtype == VKRT_U8 ? 1 : (dtype == VKRT_F16 ? 2 : 4);
    CK(cudaMalloc(&c->lin_a, (size_t)nx * ny * nz * eb));
    CK(launch_synth(c->lin_a, kind, dtype, nx, ny, nz, 0, 0, 0, nx, ny, nz, seed, c->stream));
    c->kind = VOL_SCALAR;
    c->dtype = dtype;
    return build_occupancy(c);
}

int vkrt_generate_synthetic(VkrtContext* c, int kind, int dtype, int nx, int ny, int nz, uint32_t seed) { return volume_guard(c, generate_synthetic_impl(c, kind, dtype, nx, ny, nz, seed)); }

static int scalar_to_rgba16f_impl(VkrtContext* c) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    if (c->kind != VOL_SCALAR || c->windowed) return fail(VKRT_ERR_NO_VOLUME, "no (whole) scalar volume resident");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const size_t bytes = (size_t)c->nx * c->ny * c->nz * 8;
    void *col = nullptr, *nrm = nullptr;
    cudaError_t e = cudaMalloc(&col, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&nrm, bytes);
    if (e == cudaSuccess) e = launch_scalar_to_rgba16f(c->lin_a, c->dtype, (uint2*)col, (uint2*)nrm, c->nx, c->ny, c->nz, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { cudaFree(col); cudaFree(nrm); return cuda_fail(e, "vkrt_scalar_to_rgba16f"); }
    const int nx = c->nx, ny = c->ny, nz = c->nz;
    free_volume(c);
    int rc = set_dims(c, nx, ny, nz);
    if (rc) { cudaFree(col); cudaFree(nrm); return rc; }
    c->lin_a = col;
    c->lin_b = nrm;
    c->kind = VOL_RGBA16F;
    return build_occupancy(c);
}

int vkrt_scalar_to_rgba16f(VkrtContext* c) { return volume_guard(c, scalar_to_rgba16f_impl(c)); }

int vkrt_download_scalar(VkrtContext* c, void* out) {
    if (!c || !out) return fail(VKRT_ERR_INVALID, "NULL argument");
    if (c->kind != VOL_SCALAR) return fail(VKRT_ERR_NO_VOLUME, "no scalar volume resident");
    CK(cudaSetDevice(c->device));
    const size_t eb = c->dtype == VKRT_U8 ? 1 : (c->dtype == VKRT_F16 ? 2 : 4);
    CK(cudaMemcpyAsync(out, c->lin_a, (size_t)c->nx * c->ny * c->nz * eb, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return VKRT_OK;
}

int vkrt_download_rgba16f(VkrtContext* c, uint16_t* color, uint16_t* normal) {
    if (!c || !color || !normal) return fail(VKRT_ERR_INVALID, "NULL argument");
    if (c->kind != VOL_RGBA16F) return fail(VKRT_ERR_NO_VOLUME, "no rgba16f volume resident");
    CK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->nx * c->ny * c->nz * 8;
    CK(cudaMemcpyAsync(color, c->lin_a, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(normal, c->lin_b, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return VKRT_OK;
}

int vkrt_render(VkrtContext* c, const VkrtCameraUniform* cam, const VkrtUniform* un, const VkrtOffset* offset) {
    return do_render(c, cam, un, offset, offset ? 1 : 0);
}

int vkrt_render_tiles(VkrtContext* c, const VkrtCameraUniform* cam, const VkrtUniform* un, const VkrtOffset* offsets, int n) {
    if (!offsets || n <= 0) return fail(VKRT_ERR_INVALID, "vkrt_render_tiles needs at least one offset");
    return do_render(c, cam, un, offsets, n);
}

int vkrt_box_screen_bounds(const VkrtCameraUniform* cam, int width, int height, float rect[4], int* centre_row) {
    if (!cam || !rect || width <= 0 || height <= 0) return fail(VKRT_ERR_INVALID, "bad box_screen_bounds arguments");
    int row = -1;
    cull_rect(cam->inv_proj, width, height, rect, &row);
    if (centre_row) *centre_row = row;
    return VKRT_OK;
}

int vkrt_box_screen_hull(const VkrtCameraUniform* cam, int width, int height, float planes[6][3]) {
    if (!cam || !planes || width <= 0 || height <= 0) return fail(VKRT_ERR_INVALID, "bad box_screen_hull arguments");
    static_assert(kHullEdges == 6, "vkrt_box_screen_hull returns six half-planes");
    float rect[4];
    int row = -1;
    cull_rect(cam->inv_proj, width, height, rect, &row, planes);
    return VKRT_OK;
}

int vkrt_tile_table(int width, int height, int tile_size, VkrtOffset* out, int cap) {
    if (width <= 0 || height <= 0 || tile_size <= 0) return fail(VKRT_ERR_INVALID, "bad tile table arguments");
    // examples/xor/main.rs:80-95: for y in 0..(h/TILE_SIZE)+1 { for x in 0..(w/TILE_SIZE)+1 { ... } }
    int k = 0;
    for (int y = 0; y < height / tile_size + 1; ++y)
        for (int x = 0; x < width / tile_size + 1; ++x, ++k)
            if (out && k < cap) {
                out[k].x = (float)(x * tile_size);
                out[k].y = (float)(y * tile_size);
            }
    return k;
}

int vkrt_present(VkrtContext* c) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(launch_present(c->frame, c->rgba8, c->W, c->H, c->stream));
    return VKRT_OK;
}

int vkrt_present_scaled(VkrtContext* c, int out_w, int out_h, uint8_t* rgba8) {
    if (!c || !rgba8) return fail(VKRT_ERR_INVALID, "NULL argument");
    if (out_w <= 0 || out_h <= 0 || out_w > 32768 || out_h > 32768) return fail(VKRT_ERR_INVALID, "target size out of range");
    CK(cudaSetDevice(c->device));
    uint32_t* tmp = nullptr;
    const size_t bytes = (size_t)out_w * out_h * 4;
    CK(cudaMalloc(&tmp, bytes));
    cudaError_t e = launch_present_scaled(c->frame, tmp, c->W, c->H, out_w, out_h, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(rgba8, tmp, bytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return cuda_fail(e, "vkrt_present_scaled");
    return VKRT_OK;
}

int vkrt_readback(VkrtContext* c, uint16_t* out) {
    if (!c || !out) return fail(VKRT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->frame, (size_t)c->W * c->H * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return VKRT_OK;
}

int vkrt_readback_rgba8(VkrtContext* c, uint8_t* out) {
    if (!c || !out) return fail(VKRT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->rgba8, (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return VKRT_OK;
}

int vkrt_readback_rgba8_async(VkrtContext* c, uint8_t* out) {
    if (!c || !out) return fail(VKRT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->rgba8, (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->stream));
    return VKRT_OK;
}

int vkrt_readback_aux(VkrtContext* c, uint32_t* out) {
    if (!c || !out) return fail(VKRT_ERR_INVALID, "NULL argument");
    if (!c->aux) return fail(VKRT_ERR_INVALID, "no aux buffer: render with params.count_samples = 1 first");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->aux, (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return VKRT_OK;
}

int vkrt_sync(VkrtContext* c) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->copy_stream));
    CK(cudaStreamSynchronize(c->stream2));
    for (cudaStream_t ls : c->sf_stream) if (ls) CK(cudaStreamSynchronize(ls));
    return VKRT_OK;
}

int vkrt_frame_host_async(VkrtContext* c, const VkrtCameraUniform* cam, const VkrtUniform* un, int slot) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    if (slot < 0 || slot > 1) return fail(VKRT_ERR_INVALID, "slot must be 0 or 1");
    CK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->W * c->H * 4;
    if (!c->host_slot[slot]) {
        CK(cudaMallocHost(&c->host_slot[slot], bytes));
        CK(cudaMalloc(&c->rgba8_slot[slot], bytes));
        CK(cudaEventRecord(c->ev_slot_copied[slot], c->copy_stream));
    }
    // the previous D2H out of this slot's device buffer must be done before we overwrite it
    CK(cudaStreamWaitEvent(c->stream, c->ev_slot_copied[slot], 0));
    int rc = do_render(c, cam, un, nullptr, 0, true, 1, nullptr, c->rgba8_slot[slot]);  // present fused into the raycast epilogue
    if (rc) return rc;
    CK(cudaEventRecord(c->ev_slot_ready[slot], c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev_slot_ready[slot], 0));
    CK(cudaMemcpyAsync(c->host_slot[slot], c->rgba8_slot[slot], bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    CK(cudaEventRecord(c->ev_slot_copied[slot], c->copy_stream));
    return VKRT_OK;
}

int vkrt_alloc_host(size_t bytes, void** out) {
    if (!out || bytes == 0) return fail(VKRT_ERR_INVALID, "bad argument");
    CK(cudaMallocHost(out, bytes));
    return VKRT_OK;
}

int vkrt_free_host(void* p) {
    if (p) CK(cudaFreeHost(p));
    return VKRT_OK;
}

int vkrt_host_register(void* p, size_t bytes) {
    if (!p || bytes == 0) return fail(VKRT_ERR_INVALID, "bad argument");
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return VKRT_OK;
}

int vkrt_host_unregister(void* p) {
    if (p) CK(cudaHostUnregister(p));
    return VKRT_OK;
}

namespace {
bool is_pinned_host(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}
}  // namespace

int vkrt_frame_host_wait(VkrtContext* c, int slot, uint8_t* rgba8) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    if (slot < 0 || slot > 1 || !c->host_slot[slot]) return fail(VKRT_ERR_INVALID, "slot has no frame in flight");
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->ev_slot_copied[slot]));
    if (rgba8) memcpy(rgba8, c->host_slot[slot], (size_t)c->W * c->H * 4);
    return VKRT_OK;
}

const uint8_t* vkrt_frame_host_slot_ptr(VkrtContext* c, int slot) {
    if (!c || slot < 0 || slot > 1) return nullptr;
    return c->host_slot[slot];
}

int vkrt_frame_host(VkrtContext* c, const VkrtCameraUniform* cam, const VkrtUniform* un, uint8_t* rgba8) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    if (!rgba8) return fail(VKRT_ERR_INVALID, "rgba8 is NULL");
    if (is_pinned_host(rgba8)) {
        // caller's buffer is page-locked (vkrt_alloc_host / cudaHostAlloc): DMA straight into it
        CK(cudaSetDevice(c->device));
        int rc = do_render(c, cam, un, nullptr, 0, true, 1, nullptr, c->rgba8);  // present fused into the raycast epilogue
        if (rc) return rc;
        CK(cudaMemcpyAsync(rgba8, c->rgba8, (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return VKRT_OK;
    }
    int rc = vkrt_frame_host_async(c, cam, un, 0);
    if (rc) return rc;
    return vkrt_frame_host_wait(c, 0, rgba8);
}

int vkrt_render_batch(VkrtContext* c, const VkrtCameraUniform* cams, int n, const VkrtUniform* un) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    if (!cams || n < 1 || n > VKRT_MAX_BATCH) return fail(VKRT_ERR_INVALID, "batch needs 1..VKRT_MAX_BATCH cameras");
    CK(cudaSetDevice(c->device));
    int rc = ensure_batch(c, n);
    if (rc) return rc;
    // group 0 may still be read by the copy stream of an earlier vkrt_frames_host
    CK(cudaStreamWaitEvent(c->stream, c->ev_group_copied[0], 0));
    rc = do_render(c, cams, un, nullptr, 0, true, n, c->batch_frames[0]);
    if (rc) return rc;
    c->batch_last_n = n;
    return VKRT_OK;
}

void* vkrt_batch_frame_device_ptr(VkrtContext* c, int i) {
    if (!c || i < 0 || i >= c->batch_last_n || !c->batch_frames[0]) return nullptr;
    return c->batch_frames[0] + (size_t)i * c->W * c->H;
}

int vkrt_readback_batch(VkrtContext* c, int i, uint16_t* out) {
    if (!c || !out) return fail(VKRT_ERR_INVALID, "NULL argument");
    if (i < 0 || i >= c->batch_last_n || !c->batch_frames[0]) return fail(VKRT_ERR_INVALID, "no such frame in the last batch");
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->W * c->H;
    CK(cudaMemcpyAsync(out, c->batch_frames[0] + (size_t)i * n, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return VKRT_OK;
}

int vkrt_frames_host(VkrtContext* c, const VkrtCameraUniform* cams, int n, const VkrtUniform* un, uint8_t* rgba8, int group) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    if (!cams || !rgba8 || n < 1) return fail(VKRT_ERR_INVALID, "bad frames_host arguments");
    if (group <= 0) group = 4;  // frames per launch: 3-4 fill a B200 at 1080p (profiles/r01_batch.md)
    if (group > VKRT_MAX_BATCH) group = VKRT_MAX_BATCH;
    CK(cudaSetDevice(c->device));
    if (c->params.count_samples) {  // the counting kernel is per frame: no batching
        for (int i = 0; i < n; ++i) {
            int rc = vkrt_frame_host(c, cams + i, un, rgba8 + (size_t)i * c->W * c->H * 4);
            if (rc) return rc;
        }
        return VKRT_OK;
    }
    int rc = ensure_batch(c, group);
    if (rc) return rc;
    const size_t px = (size_t)c->W * c->H;
    // Software pipeline over groups of `group` frames, two device buffers: the raycast of group j+1 (context
    // stream, ONE launch, present fused into its epilogue) overlaps the D2H of group j (copy stream).
    // The groups ramp up (1, 2, 4, ... frames): nothing can be copied before the first group has been rendered, so the first
    // launch is a single frame and the copy engine starts ~0.2 ms earlier (the D2H, ~0.15 ms per 1080p frame, is the bottleneck).
    for (int first = 0, j = 0, g = 0; first < n; first += g, ++j) {
        g = std::min(std::min(group, 1 << std::min(j, 8)), n - first);
        const int b = j & 1;
        CK(cudaStreamWaitEvent(c->stream, c->ev_group_copied[b], 0));  // buffer b has left the device (group j-2)
        rc = do_render(c, cams + first, un, nullptr, 0, false, g, c->batch_frames[b], c->batch_rgba8[b]);  // present fused
        if (rc) return rc;
        CK(cudaEventRecord(c->ev_group_ready[b], c->stream));
        CK(cudaStreamWaitEvent(c->copy_stream, c->ev_group_ready[b], 0));
        CK(cudaMemcpyAsync(rgba8 + (size_t)first * px * 4, c->batch_rgba8[b], (size_t)g * px * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        CK(cudaEventRecord(c->ev_group_copied[b], c->copy_stream));
    }
    c->batch_last_n = 0;  // the group buffers were recycled: nothing to read back through vkrt_readback_batch
    CK(cudaStreamSynchronize(c->copy_stream));
    return VKRT_OK;
}

void* vkrt_frame_device_ptr(VkrtContext* c) { return c ? c->frame : nullptr; }
void* vkrt_frame_rgba8_device_ptr(VkrtContext* c) { return c ? c->rgba8 : nullptr; }
void* vkrt_stream(VkrtContext* c) { return c ? (void*)c->stream : nullptr; }

int vkrt_stats(VkrtContext* c, VkrtStats* out) {
    if (!c || !out) return fail(VKRT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    unsigned long long h[3];
    CK(cudaMemcpy(h, c->counters, sizeof h, cudaMemcpyDeviceToHost));
    out->rays_hit = h[0];
    out->samples_reference = h[1];
    out->samples_fetched = h[2];
    out->last_render_ms = 0.0f;
    out->_pad = 0.0f;
    if (c->timed) CK(cudaEventElapsedTime(&out->last_render_ms, c->ev_begin, c->ev_end));
    return VKRT_OK;
}

int vkrt_reset_stats(VkrtContext* c) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(c->counters, 0, 3 * sizeof(unsigned long long), c->stream));
    return VKRT_OK;
}

int vkrt_timing_enable(VkrtContext* c, int capacity) {
    if (!c || capacity < 0 || capacity > (1 << 20)) return fail(VKRT_ERR_INVALID, "bad timing capacity");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (cudaEvent_t e : c->ring_begin) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ring_end) cudaEventDestroy(e);
    c->ring_begin.clear();
    c->ring_end.clear();
    c->ring_next = c->ring_count = 0;
    for (int i = 0; i < capacity; ++i) {
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        c->ring_begin.push_back(a);
        c->ring_end.push_back(b);
    }
    return VKRT_OK;
}

int vkrt_timing_read(VkrtContext* c, float* ms, int n) {
    if (!c || !ms || n < 0) return fail(VKRT_ERR_INVALID, "bad argument");
    if ((size_t)n > c->ring_count) return fail(VKRT_ERR_INVALID, "fewer renders recorded than requested");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->copy_stream));  // the sort-first root renders (and records its events) on its other streams
    CK(cudaStreamSynchronize(c->stream2));
    for (cudaStream_t ls : c->sf_stream) if (ls) CK(cudaStreamSynchronize(ls));
    const size_t cap = c->ring_begin.size();
    for (int i = 0; i < n; ++i) {  // oldest of the last n first
        const size_t k = (c->ring_next + cap - (size_t)n + (size_t)i) % cap;
        CK(cudaEventElapsedTime(ms + i, c->ring_begin[k], c->ring_end[k]));
    }
    return VKRT_OK;
}

namespace {
int flush_l2_on(VkrtContext* c, cudaStream_t s) {
    if (!c->flush_buf) {
        int l2 = 0;
        CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, c->device));
        c->flush_bytes = (size_t)l2 * 2 > ((size_t)256 << 20) ? (size_t)l2 * 2 : ((size_t)256 << 20);
        CK(cudaMalloc(&c->flush_buf, c->flush_bytes));
    }
    CK(launch_flush_l2((uint4*)c->flush_buf, c->flush_bytes / 16, s));
    return VKRT_OK;
}
}  // namespace

int vkrt_flush_l2(VkrtContext* c) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    return flush_l2_on(c, c->stream);
}

// ---- sort-first over several GPUs, one process per GPU --------------------------------------------
// Shared block in rank 0's memory: [mailbox 4 KiB: u64 consumed, u64 timeouts, u64 arrive[slots]][slot 0]...[slot S-1]
namespace {
constexpr size_t kSfMailbox = 4096;
constexpr int kSfMaxSlots = 256;  // arrival flags live in the 4 KiB mailbox: 2 + slots u64 words
inline size_t sf_frame_bytes(const VkrtContext* c) { return (size_t)c->W * c->H * sizeof(uint2); }
inline unsigned long long* sf_consumed(VkrtContext* c) { return reinterpret_cast<unsigned long long*>(c->sf_base); }
inline unsigned long long* sf_timeouts(VkrtContext* c) { return reinterpret_cast<unsigned long long*>(c->sf_base) + 1; }
inline unsigned long long* sf_arrive(VkrtContext* c, int slot) { return reinterpret_cast<unsigned long long*>(c->sf_base) + 2 + slot; }
inline uint2* sf_slot(VkrtContext* c, int slot) { return reinterpret_cast<uint2*>(c->sf_base + kSfMailbox + (size_t)slot * sf_frame_bytes(c)); }
}  // namespace

int vkrt_sortfirst_create_root(VkrtContext* c, int world, int slots, VkrtSortFirstHandle* out) {
    if (!c || !out || world < 1 || slots < 2 || slots > kSfMaxSlots) return fail(VKRT_ERR_INVALID, "bad argument (2 <= slots <= 256)");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    sf_release(c);
    const size_t bytes = kSfMailbox + (size_t)slots * sf_frame_bytes(c);
    CK(cudaMalloc(&c->sf_base, bytes));
    // on the context's stream and completed before the handle leaves: the legacy default stream does
    // not order against our non-blocking stream, nor against the peers
    CK(cudaMemsetAsync(c->sf_base, 0, bytes, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->sf_owner = true;
    c->sf_rank = 0;
    c->sf_world = world;
    c->sf_slots = slots;
    c->own_frame = c->frame;
    memset(out, 0, sizeof *out);
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->sf_base));
    static_assert(sizeof(h) <= sizeof(out->ipc), "IPC handle size");
    memcpy(out->ipc, &h, sizeof h);
    out->width = c->W;
    out->height = c->H;
    out->world = world;
    out->slots = slots;
    return VKRT_OK;
}

int vkrt_sortfirst_join(VkrtContext* c, int rank, const VkrtSortFirstHandle* h) {
    if (!c || !h || rank < 1 || rank >= h->world) return fail(VKRT_ERR_INVALID, "bad argument");
    if (h->width != c->W || h->height != c->H) return fail(VKRT_ERR_INVALID, "frame size differs from the root's");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    sf_release(c);
    cudaIpcMemHandle_t ih;
    memcpy(&ih, h->ipc, sizeof ih);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
    c->sf_base = (unsigned char*)p;
    c->sf_owner = false;
    c->sf_rank = rank;
    c->sf_world = h->world;
    c->sf_slots = h->slots;
    c->own_frame = c->frame;
    return VKRT_OK;
}

int vkrt_sortfirst_leave(VkrtContext* c) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->copy_stream));  // group transfers into rank 0's ring may still be in flight
    CK(cudaStreamSynchronize(c->stream2));
    for (cudaStream_t ls : c->sf_stream) if (ls) CK(cudaStreamSynchronize(ls));
    CK(cudaStreamSynchronize(c->copy_stream));  // pushes queued behind a lane's render
    sf_release(c);
    return VKRT_OK;
}

int vkrt_sortfirst_partition(int width, int height, int tile_size, int rank, int world, VkrtOffset* out, int cap) {
    if (width <= 0 || height <= 0 || tile_size <= 0 || world < 1 || rank < 0 || rank >= world) return fail(VKRT_ERR_INVALID, "bad partition arguments");
    // Tiles that intersect the frame, dealt in a skewed round-robin: tile (tx, ty) goes to rank
    // (tx + 3*ty) mod world. A plain t mod world hands every rank whole COLUMNS of tiles whenever the number
    // of tile columns is a multiple of world (32 columns at 4K / 120 px), which balances badly on content
    // that varies across the image; the skew spreads each rank's tiles over both axes.
    const int cols = (width + tile_size - 1) / tile_size, rows = (height + tile_size - 1) / tile_size;
    int k = 0;
    for (int ty = 0; ty < rows; ++ty)
        for (int tx = 0; tx < cols; ++tx) {
            if ((tx + 3 * ty) % world != rank) continue;
            if (out && k < cap) {
                out[k].x = (float)(tx * tile_size);
                out[k].y = (float)(ty * tile_size);
            }
            ++k;
        }
    return k;
}

namespace {
// Sort-first tiles: the pixel rectangle of frame f of a push that can hold anything but the clear colour — the kernel's
// own cull rectangle (cull_rect: ray coordinates gid + offset; a fractional offset stores up to one pixel to the left /
// above, hence 2 pixels of margin), clamped to the frame, x0 even and x1 odd so that 16-byte stores stay aligned. The
// peers ship only these pixels of their tiles; rank 0 writes the clear colour everywhere else (sf_fill_outside).
void push_clip(const VkrtContext* c, const VkrtCameraUniform* cam, int f, PushClip& clip) {
    float r[4];
    int row;
    cull_rect(cam->inv_proj, c->W, c->H, r, &row);
    const double W1 = c->W - 1, H1 = c->H - 1;
    int x0 = (int)std::max(0.0, std::min(W1 + 1.0, std::floor((double)r[0]) - 2.0)), x1 = (int)std::min(W1, std::max(-1.0, std::ceil((double)r[2]) + 2.0));
    int y0 = (int)std::max(0.0, std::min(H1 + 1.0, std::floor((double)r[1]) - 2.0)), y1 = (int)std::min(H1, std::max(-1.0, std::ceil((double)r[3]) + 2.0));
    x0 &= ~1;
    if (!(x1 & 1) && x1 < c->W - 1) ++x1;
    if (x0 > x1 || y0 > y1) { x0 = 0; y0 = 0; x1 = -1; y1 = -1; }  // the box is off screen: nothing to ship
    clip.x0[f] = x0; clip.y0[f] = y0; clip.x1[f] = x1; clip.y1[f] = y1;
}
int sf_fill_outside(VkrtContext* c, uint2* frame, const PushClip& clip, int f, cudaStream_t s) {
    const float* cc = c->params.clear_color;
    const __half2 lo = __floats2half2_rn(cc[0], cc[1]), hi = __floats2half2_rn(cc[2], 1.0f);  // what the kernel stores for a miss
    uint2 texel;
    memcpy(&texel.x, &lo, 4);
    memcpy(&texel.y, &hi, 4);
    CK(launch_fill_outside(frame, c->W, c->H, clip.x0[f], clip.y0[f], clip.x1[f], clip.y1[f], texel, s));
    return VKRT_OK;
}
}  // namespace

int vkrt_sortfirst_render(VkrtContext* c, const VkrtCameraUniform* cam, const VkrtUniform* un, const VkrtOffset* offsets, int n,
                          uint64_t frame_index) {
    if (!c || !c->sf_base) return fail(VKRT_ERR_INVALID, "context is not in a sort-first group");
    CK(cudaSetDevice(c->device));
    const int slot = (int)(frame_index % (uint64_t)c->sf_slots);
    const bool tiles = offsets && n > 0;
    cudaEvent_t eb = nullptr, ee = nullptr;
    if (!c->ring_begin.empty()) {
        eb = c->ring_begin[c->ring_next];
        ee = c->ring_end[c->ring_next];
        c->ring_next = (c->ring_next + 1) % c->ring_begin.size();
        if (c->ring_count < c->ring_begin.size()) ++c->ring_count;
    }
    // slot reuse: the frame that used this slot last (frame_index - slots) must have been consumed
    const bool must_wait = frame_index >= (uint64_t)c->sf_slots;
    const uint64_t consumed_target = must_wait ? frame_index - (uint64_t)c->sf_slots + 1 : 0;
    int rc = sf_lanes_ensure(c, c->sf_rank != 0);
    if (rc) return rc;
    const int lane = c->sf_lane;
    c->sf_lane = (c->sf_lane + 1) % VkrtContext::kSfLanes;
    cudaStream_t ls = c->sf_stream[lane];
    if (c->sf_rank == 0) {
        // root: its tiles go straight into the ring slot (local memory) — never on the context's own stream: that one
        // carries the in-order waits for every frame, and a render queued behind the wait for the peers' tiles of frame f
        // would keep the root from starting its share of frame f + 1. Consecutive frames rotate over the lanes: the next
        // frames' launches fill the SMs that the tail of this one (its longest rays) leaves idle.
        if (must_wait) CK(launch_flag_wait(sf_consumed(c), consumed_target, sf_timeouts(c), ls));
        if (tiles && c->sf_world > 1) {  // the peers do not ship what lies outside the cull rectangle
            PushClip clip;
            push_clip(c, cam, 0, clip);
            rc = sf_fill_outside(c, sf_slot(c, slot), clip, 0, ls);
            if (rc) return rc;
        }
        if (eb) CK(cudaEventRecord(eb, ls));
        rc = do_render(c, cam, un, tiles ? offsets : nullptr, tiles ? n : 0, false, 1, sf_slot(c, slot), nullptr, ls);
        if (rc) return rc;
        if (ee) CK(cudaEventRecord(ee, ls));
        CK(launch_flag_add(sf_arrive(c, slot), 1ull, ls));
        return VKRT_OK;
    }
    // peer: render into the lane's LOCAL frame, then ship the tiles (or the whole frame) to rank 0's ring slot on the copy
    // stream — whole tile rows in 16-byte stores over NVLink — overlapping this rank's next launches. The slot-reuse wait
    // sits on the copy stream, so rendering runs ahead of rank 0's consumption by up to kSfLanes frames.
    CK(cudaStreamWaitEvent(ls, c->sf_copied[lane], 0));  // the lane's local frame has left (kSfLanes frames back)
    if (eb) CK(cudaEventRecord(eb, ls));
    rc = do_render(c, cam, un, tiles ? offsets : nullptr, tiles ? n : 0, false, 1, c->sf_local[lane], nullptr, ls);
    if (rc) return rc;
    if (ee) CK(cudaEventRecord(ee, ls));
    CK(cudaEventRecord(c->sf_ready[lane], ls));
    CK(cudaStreamWaitEvent(c->copy_stream, c->sf_ready[lane], 0));
    if (must_wait) CK(launch_flag_wait(sf_consumed(c), consumed_target, sf_timeouts(c), c->copy_stream));
    if (tiles) {
        bool vec16 = (c->W % 2 == 0) && (c->params.tile_size % 2 == 0);
        for (int i = 0; i < n && vec16; ++i) vec16 = ((long long)offsets[i].x % 2) == 0 && offsets[i].x >= 0.0f;
        PushClip clip;
        push_clip(c, cam, 0, clip);
        CK(launch_push_tiles(c->sf_local[lane], sf_slot(c, slot), c->d_offsets, n, c->params.tile_size, c->W, c->H, vec16, c->copy_stream, 1, &clip));
    } else {
        CK(cudaMemcpyAsync(sf_slot(c, slot), c->sf_local[lane], sf_frame_bytes(c), cudaMemcpyDeviceToDevice, c->copy_stream));
    }
    CK(launch_flag_add(sf_arrive(c, slot), 1ull, c->copy_stream));  // after the transfer, system scope
    CK(cudaEventRecord(c->sf_copied[lane], c->copy_stream));
    return VKRT_OK;
}

int vkrt_sortfirst_render_tiles_batch(VkrtContext* c, const VkrtCameraUniform* cams, int n_frames, const VkrtUniform* un, const VkrtOffset* offsets, int n,
                                      uint64_t first_frame) {
    if (!c || !c->sf_base) return fail(VKRT_ERR_INVALID, "context is not in a sort-first group");
    if (!cams || !offsets || n < 1 || n_frames < 1 || n_frames > kMaxPushDst || n_frames > VKRT_MAX_BATCH)
        return fail(VKRT_ERR_INVALID, "tile batch needs 1..15 cameras and this rank's tiles");
    const int slot0 = (int)(first_frame % (uint64_t)c->sf_slots);
    if (slot0 + n_frames > c->sf_slots) return fail(VKRT_ERR_INVALID, "a batch must occupy consecutive ring slots (slots % batch == 0, first_frame % batch == 0)");
    CK(cudaSetDevice(c->device));
    cudaEvent_t eb = nullptr, ee = nullptr;
    if (!c->ring_begin.empty()) {
        eb = c->ring_begin[c->ring_next];
        ee = c->ring_end[c->ring_next];
        c->ring_next = (c->ring_next + 1) % c->ring_begin.size();
        if (c->ring_count < c->ring_begin.size()) ++c->ring_count;
    }
    // slot reuse: the frame that used the LAST of these slots before must have been consumed (frames are consumed in order)
    const uint64_t last = first_frame + (uint64_t)n_frames - 1;
    const bool must_wait = last >= (uint64_t)c->sf_slots;
    const uint64_t consumed_target = must_wait ? last - (uint64_t)c->sf_slots + 1 : 0;
    int rc = sf_lanes_ensure(c, c->sf_rank != 0, n_frames);
    if (rc) return rc;
    const int lane = c->sf_lane;
    c->sf_lane = (c->sf_lane + 1) % VkrtContext::kSfLanes;
    cudaStream_t ls = c->sf_stream[lane];
    PushDst arrive{};
    for (int f = 0; f < n_frames; ++f) arrive.ptr[f] = sf_arrive(c, slot0 + f);
    arrive.n = n_frames;
    // ONE launch for this rank's tiles of n_frames consecutive frames (grid.z = frame x tile): a single frame's share on
    // 1/N of the GPUs lasts as long as its longest rays whatever its pixel count (config 4: 72 of 576 tiles take 0.30 ms
    // where 576 take 0.78 ms), so the shares of several frames are batched exactly like whole frames are.
    if (c->sf_rank == 0) {
        if (must_wait) CK(launch_flag_wait(sf_consumed(c), consumed_target, sf_timeouts(c), ls));
        if (c->sf_world > 1) {
            PushClip clip;
            for (int f = 0; f < n_frames; ++f) {
                push_clip(c, cams + f, f, clip);
                rc = sf_fill_outside(c, sf_slot(c, slot0 + f), clip, f, ls);
                if (rc) return rc;
            }
        }
        if (eb) CK(cudaEventRecord(eb, ls));
        rc = do_render(c, cams, un, offsets, n, false, n_frames, sf_slot(c, slot0), nullptr, ls);
        if (rc) return rc;
        if (ee) CK(cudaEventRecord(ee, ls));
        CK(launch_flag_add_many(arrive, 1ull, ls));
        return VKRT_OK;
    }
    CK(cudaStreamWaitEvent(ls, c->sf_copied[lane], 0));  // the lane's local frames have left
    if (eb) CK(cudaEventRecord(eb, ls));
    rc = do_render(c, cams, un, offsets, n, false, n_frames, c->sf_local[lane], nullptr, ls);
    if (rc) return rc;
    if (ee) CK(cudaEventRecord(ee, ls));
    CK(cudaEventRecord(c->sf_ready[lane], ls));
    CK(cudaStreamWaitEvent(c->copy_stream, c->sf_ready[lane], 0));
    if (must_wait) CK(launch_flag_wait(sf_consumed(c), consumed_target, sf_timeouts(c), c->copy_stream));
    bool vec16 = (c->W % 2 == 0) && (c->params.tile_size % 2 == 0);
    for (int i = 0; i < n && vec16; ++i) vec16 = ((long long)offsets[i].x % 2) == 0 && offsets[i].x >= 0.0f;
    PushClip clip;
    for (int f = 0; f < n_frames; ++f) push_clip(c, cams + f, f, clip);
    CK(launch_push_tiles(c->sf_local[lane], sf_slot(c, slot0), c->d_offsets, n, c->params.tile_size, c->W, c->H, vec16, c->copy_stream, n_frames, &clip));
    CK(launch_flag_add_many(arrive, 1ull, c->copy_stream));  // after the transfer, system scope
    CK(cudaEventRecord(c->sf_copied[lane], c->copy_stream));
    return VKRT_OK;
}

int vkrt_sortfirst_render_batch(VkrtContext* c, const VkrtCameraUniform* cams, int n_frames, const VkrtUniform* un, uint64_t first_frame,
                                int flags) {
    if (!c || !c->sf_base) return fail(VKRT_ERR_INVALID, "context is not in a sort-first group");
    if (!cams || n_frames < 1 || n_frames > VKRT_MAX_BATCH) return fail(VKRT_ERR_INVALID, "batch needs 1..VKRT_MAX_BATCH cameras");
    const int slot0 = (int)(first_frame % (uint64_t)c->sf_slots);
    if (slot0 + n_frames > c->sf_slots) return fail(VKRT_ERR_INVALID, "a batch must occupy consecutive ring slots (slots % batch == 0, first_frame % batch == 0)");
    CK(cudaSetDevice(c->device));
    cudaEvent_t eb = nullptr, ee = nullptr;
    if (!c->ring_begin.empty()) {
        eb = c->ring_begin[c->ring_next];
        ee = c->ring_end[c->ring_next];
        c->ring_next = (c->ring_next + 1) % c->ring_begin.size();
        if (c->ring_count < c->ring_begin.size()) ++c->ring_count;
    }
    // slot reuse: the frame that used the LAST of these slots before must have been consumed (frames are consumed in order)
    const uint64_t last = first_frame + (uint64_t)n_frames - 1;
    const bool must_wait = last >= (uint64_t)c->sf_slots;
    const uint64_t consumed_target = must_wait ? last - (uint64_t)c->sf_slots + 1 : 0;
    if (c->sf_rank == 0) {
        // root: ONE launch, grid.z = frame, straight into its own ring slots slot0 .. slot0 + n - 1 — on the
        // SECOND stream: the context's stream carries the in-order waits for every frame (its own and the peers'),
        // and a render queued behind the wait for a peer's group would idle the root until that group arrives.
        cudaStream_t rs = c->copy_stream;
        if (flags & VKRT_SF_FLUSH_L2) { int rc = flush_l2_on(c, rs); if (rc) return rc; }
        if (must_wait) CK(launch_flag_wait(sf_consumed(c), consumed_target, sf_timeouts(c), rs));
        if (eb) CK(cudaEventRecord(eb, rs));
        int rc = do_render(c, cams, un, nullptr, 0, false, n_frames, sf_slot(c, slot0), nullptr, rs);
        if (rc) return rc;
        if (ee) CK(cudaEventRecord(ee, rs));
        for (int f = 0; f < n_frames; ++f) CK(launch_flag_add(sf_arrive(c, slot0 + f), 1ull, rs));
        return VKRT_OK;
    }
    if (flags & VKRT_SF_FLUSH_L2) { int rc = flush_l2_on(c, c->stream); if (rc) return rc; }
    // peer: render the group into a local buffer, then ONE copy-engine transfer of the n contiguous frames into
    // rank 0's ring over NVLink on the copy stream, overlapping this rank's next launch. (A single frame goes out
    // by the kernel's own peer stores, vkrt_sortfirst_render; for a group the 8-byte stores of 8x4-pixel warps —
    // 64-B row segments — cost the kernel 20-30 % at 4 GPUs, while the DMA moves 133 MB in ~0.2 ms off the
    // critical path.) The slot-reuse wait sits on the copy stream too, so rendering runs ahead of consumption.
    int rc = ensure_batch(c, n_frames);
    if (rc) return rc;
    const int b = c->sf_parity;
    c->sf_parity ^= 1;
    CK(cudaStreamWaitEvent(c->stream, c->ev_group_copied[b], 0));  // local buffer b has left (two groups back)
    if (eb) CK(cudaEventRecord(eb, c->stream));
    rc = do_render(c, cams, un, nullptr, 0, false, n_frames, c->batch_frames[b]);
    if (rc) return rc;
    if (ee) CK(cudaEventRecord(ee, c->stream));
    CK(cudaEventRecord(c->ev_group_ready[b], c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev_group_ready[b], 0));
    if (must_wait) CK(launch_flag_wait(sf_consumed(c), consumed_target, sf_timeouts(c), c->copy_stream));
    CK(cudaMemcpyAsync(sf_slot(c, slot0), c->batch_frames[b], (size_t)n_frames * sf_frame_bytes(c), cudaMemcpyDeviceToDevice, c->copy_stream));
    for (int f = 0; f < n_frames; ++f) CK(launch_flag_add(sf_arrive(c, slot0 + f), 1ull, c->copy_stream));  // after the copy, system scope
    CK(cudaEventRecord(c->ev_group_copied[b], c->copy_stream));
    return VKRT_OK;
}

int vkrt_sortfirst_wait(VkrtContext* c, uint64_t frame_index, uint64_t arrivals_target) {
    if (!c || !c->sf_base || c->sf_rank != 0) return fail(VKRT_ERR_INVALID, "only the root waits for frames");
    CK(cudaSetDevice(c->device));
    const int slot = (int)(frame_index % (uint64_t)c->sf_slots);
    CK(launch_flag_wait(sf_arrive(c, slot), arrivals_target, sf_timeouts(c), c->stream));
    c->frame = sf_slot(c, slot);
    return VKRT_OK;
}

int vkrt_sortfirst_timeouts(VkrtContext* c, uint64_t* out) {
    if (!c || !c->sf_base || !out) return fail(VKRT_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    unsigned long long v = 0;
    CK(cudaMemcpy(&v, sf_timeouts(c), sizeof v, cudaMemcpyDeviceToHost));
    *out = v;
    return VKRT_OK;
}

int vkrt_sortfirst_consume(VkrtContext* c, uint64_t frame_index, int do_present) {
    if (!c || !c->sf_base || c->sf_rank != 0) return fail(VKRT_ERR_INVALID, "only the root consumes frames");
    CK(cudaSetDevice(c->device));
    const int slot = (int)(frame_index % (uint64_t)c->sf_slots);
    c->frame = sf_slot(c, slot);
    if (do_present) CK(launch_present(c->frame, c->rgba8, c->W, c->H, c->stream));
    CK(launch_flag_set(sf_consumed(c), frame_index + 1, c->stream));
    return VKRT_OK;
}

// ---- sort-last: one brick of a partitioned grid per context ------------------------------------------
namespace {
int fill_partial_args(VkrtContext* c, const VkrtCameraUniform* cam, PartialArgs& A) {
    if (!c->windowed) return fail(VKRT_ERR_INVALID, "no windowed volume resident (vkrt_upload_window / vkrt_generate_synthetic_window)");
    const VkrtParams& P = c->params;
    if ((P.mode == VKRT_MODE_M0) != (c->kind == VOL_RGBA16F)) return fail(VKRT_ERR_INVALID, "mode does not match the resident volume kind");
    if (P.mode == VKRT_MODE_M0 && P.clear_color[3] != 0.0f) return fail(VKRT_ERR_UNSUPPORTED, "sort-last needs clear_color.a == 0");
    // the reference tests the threshold only after compositing a sample: with initial_alpha >= alpha_threshold it takes
    // exactly one sample, which the per-brick leaps and the a_in >= threshold shortcut would drop
    if (!(P.initial_alpha < P.alpha_threshold)) return fail(VKRT_ERR_UNSUPPORTED, "sort-last needs initial_alpha < alpha_threshold");
    memset(&A, 0, sizeof A);
    if (cam) memcpy(A.inv, cam->inv_proj, sizeof A.inv);
    A.W = c->W; A.H = c->H;
    A.gnx = c->gn[0]; A.gny = c->gn[1]; A.gnz = c->gn[2];
    A.fx = (float)A.gnx; A.fy = (float)A.gny; A.fz = (float)A.gnz;
    A.hx = A.fx / 2.0f; A.hy = A.fy / 2.0f; A.hz = A.fz / 2.0f;
    A.vol_a = c->lin_a; A.vol_b = c->lin_b;
    A.wx = c->win_lo[0]; A.wy = c->win_lo[1]; A.wz = c->win_lo[2];
    A.nx = c->nx; A.ny = c->ny; A.nz = c->nz;
    A.bricked = c->win_bricked ? 1 : 0;
    A.bnx = (c->nx + 7) >> 3; A.bny = (c->ny + 7) >> 3;
    for (int i = 0; i < 3; ++i) { A.own_lo[i] = c->own_lo[i]; A.own_hi[i] = c->own_hi[i]; }
    A.dist = c->dist;
    A.cox = c->cell_lo[0]; A.coy = c->cell_lo[1]; A.coz = c->cell_lo[2];
    A.cnx = c->cell_n[0]; A.cny = c->cell_n[1]; A.cnz = c->cell_n[2];
    const int gmax = A.gnx > A.gny ? (A.gnx > A.gnz ? A.gnx : A.gnz) : (A.gny > A.gnz ? A.gny : A.gnz);
    A.leap_eps = 16.0f * 1.1920929e-07f * (float)gmax;
    A.dt_scale = P.dt_scale; A.dt_floor = P.dt_floor; A.alpha_threshold = P.alpha_threshold;
    memcpy(A.clear, P.clear_color, sizeof A.clear);
    return VKRT_OK;
}

int install_window(VkrtContext* c, int kind, int dtype, const int gn[3], const int own_lo[3], const int own_hi[3]) {
    for (int i = 0; i < 3; ++i) {
        if (gn[i] <= 0 || gn[i] > 8192 || own_lo[i] < 0 || own_hi[i] > gn[i] || own_lo[i] >= own_hi[i]) return fail(VKRT_ERR_INVALID, "bad grid / brick range");
        if (own_lo[i] % 8 != 0 || (own_hi[i] % 8 != 0 && own_hi[i] != gn[i])) return fail(VKRT_ERR_INVALID, "brick bounds must be multiples of 8 (or the grid edge)");
        c->gn[i] = gn[i]; c->own_lo[i] = own_lo[i]; c->own_hi[i] = own_hi[i];
        c->win_lo[i] = own_lo[i] > 0 ? own_lo[i] - 1 : 0;  // one-voxel halo, clamped to the grid
        c->cell_lo[i] = own_lo[i] >> 3;
        c->cell_n[i] = ((own_hi[i] - 1) >> 3) - c->cell_lo[i] + 1;
    }
    c->nx = (own_hi[0] < gn[0] ? own_hi[0] + 1 : gn[0]) - c->win_lo[0];
    c->ny = (own_hi[1] < gn[1] ? own_hi[1] + 1 : gn[1]) - c->win_lo[1];
    c->nz = (own_hi[2] < gn[2] ? own_hi[2] + 1 : gn[2]) - c->win_lo[2];
    c->nbx = c->nby = c->nbz = 0;
    c->kind = kind;
    c->dtype = dtype;
    c->windowed = true;
    c->win_bricked = kind == VOL_SCALAR;
    return VKRT_OK;
}

int build_window_occupancy(VkrtContext* c) {
    PartialArgs A;
    const int mode = c->kind == VOL_RGBA16F ? VKRT_MODE_M0 : VKRT_MODE_M1;
    VkrtParams keep = c->params;
    vkrt_default_params(mode, &c->params);
    int rc = fill_partial_args(c, nullptr, A);
    c->params = keep;
    if (rc) return rc;
    const size_t cells = (size_t)A.cnx * A.cny * A.cnz;
    uint8_t* scratch = nullptr;
    CK(cudaMalloc(&c->dist, cells));
    CK(cudaMalloc(&scratch, cells));
    cudaError_t e = launch_window_occupancy(A, mode, c->dtype, c->dist, c->stream);
    // outside the own cells is other ranks' territory (or outside the grid): no-ops for this rank
    if (e == cudaSuccess) e = launch_distance_transform(c->dist, scratch, A.cnx, A.cny, A.cnz, 255, kMaxLeapBricks, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(scratch);
    if (e != cudaSuccess) return cuda_fail(e, "build_window_occupancy");
    return VKRT_OK;
}
}  // namespace

static int upload_window_impl(VkrtContext* c, const void* a, const void* b, int dtype, const int gn[3], const int own_lo[3], const int own_hi[3]) {
    if (!c || !a || !gn || !own_lo || !own_hi) return fail(VKRT_ERR_INVALID, "NULL argument");
    if (dtype < -1 || dtype > VKRT_F32) return fail(VKRT_ERR_INVALID, "dtype must be -1 (rgba16f pair) or a VkrtDtype");
    if (dtype == -1 && !b) return fail(VKRT_ERR_INVALID, "rgba16f window needs both arrays");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    free_volume(c);
    int rc = install_window(c, dtype == -1 ? VOL_RGBA16F : VOL_SCALAR, dtype == -1 ? 0 : dtype, gn, own_lo, own_hi);
    if (rc) return rc;
    const size_t eb = dtype == -1 ? 8 : (dtype == VKRT_U8 ? 1 : (dtype == VKRT_F16 ? 2 : 4));
    const size_t bytes = (size_t)c->nx * c->ny * c->nz * eb;
    if (dtype == -1) {
        CK(cudaMalloc(&c->lin_a, bytes));
        CK(cudaMemcpyAsync(c->lin_a, a, bytes, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMalloc(&c->lin_b, bytes));
        CK(cudaMemcpyAsync(c->lin_b, b, bytes, cudaMemcpyHostToDevice, c->stream));
    } else {
        void* lin = nullptr;
        const size_t padded = (size_t)((c->nx + 7) >> 3) * ((c->ny + 7) >> 3) * ((c->nz + 7) >> 3) * 512 * eb;
        CK(cudaMalloc(&lin, bytes));
        cudaError_t e = cudaMalloc(&c->lin_a, padded);
        if (e == cudaSuccess) e = cudaMemcpyAsync(lin, a, bytes, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = launch_brick_window(lin, c->lin_a, (int)eb, c->nx, c->ny, c->nz, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(lin);  // staging copy: released on every exit
        if (e != cudaSuccess) return cuda_fail(e, "vkrt_upload_window: staging / bricking the window");
    }
    return build_window_occupancy(c);
}

int vkrt_upload_window(VkrtContext* c, const void* a, const void* b, int dtype, const int gn[3], const int own_lo[3], const int own_hi[3]) { return volume_guard(c, upload_window_impl(c, a, b, dtype, gn, own_lo, own_hi)); }

static int generate_synthetic_window_impl(VkrtContext* c, int kind, int dtype, const int gn[3], const int own_lo[3], const int own_hi[3], uint32_t seed) {
    if (!c || !gn || !own_lo || !own_hi) return fail(VKRT_ERR_INVALID, "NULL argument");
    if (kind < 0 || kind > 3 || dtype < VKRT_U8 || dtype > VKRT_F32) return fail(VKRT_ERR_INVALID, "bad kind / dtype");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    free_volume(c);
    int rc = install_window(c, VOL_SCALAR, dtype, gn, own_lo, own_hi);
    if (rc) return rc;
    const size_t eb = dtype == VKRT_U8 ? 1 : (dtype == VKRT_F16 ? 2 : 4);
    CK(cudaMalloc(&c->lin_a, (size_t)((c->nx + 7) >> 3) * ((c->ny + 7) >> 3) * ((c->nz + 7) >> 3) * 512 * eb));  // bricked, padded
    CK(launch_synth(c->lin_a, kind, dtype, c->nx, c->ny, c->nz, c->win_lo[0], c->win_lo[1], c->win_lo[2], gn[0], gn[1], gn[2], seed, c->stream, 1));
    return build_window_occupancy(c);
}

int vkrt_generate_synthetic_window(VkrtContext* c, int kind, int dtype, const int gn[3], const int own_lo[3], const int own_hi[3], uint32_t seed) { return volume_guard(c, generate_synthetic_window_impl(c, kind, dtype, gn, own_lo, own_hi, seed)); }

int vkrt_window_info(VkrtContext* c, int win_lo[3], int win_n[3]) {
    if (!c || !c->windowed) return fail(VKRT_ERR_INVALID, "no windowed volume resident");
    for (int i = 0; i < 3; ++i) if (win_lo) win_lo[i] = c->win_lo[i];
    if (win_n) { win_n[0] = c->nx; win_n[1] = c->ny; win_n[2] = c->nz; }
    return VKRT_OK;
}

int vkrt_partial_alpha(VkrtContext* c, const VkrtCameraUniform* cam, const VkrtUniform* un, float* d_T) {
    if (!c || !cam || !un || !d_T) return fail(VKRT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    PartialArgs A;
    int rc = fill_partial_args(c, cam, A);
    if (rc) return rc;
    A.T_out = d_T;
    CK(launch_partial(A, c->params.mode, c->dtype, 1, c->stream));
    return VKRT_OK;
}

int vkrt_partial_ain(VkrtContext* c, const float* d_T_all, const int* ranks_before, int n_before, float* d_ain) {
    if (!c || !d_ain || n_before < 0 || (n_before > 0 && (!d_T_all || !ranks_before))) return fail(VKRT_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(c->device));
    if (!c->d_before) CK(cudaMalloc(&c->d_before, 1024 * sizeof(int)));
    if (n_before > 1024) return fail(VKRT_ERR_INVALID, "too many ranks");
    if (n_before > 0) CK(cudaMemcpyAsync(c->d_before, ranks_before, (size_t)n_before * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    const size_t n = (size_t)c->W * c->H;
    CK(launch_partial_ain(d_T_all, n, c->d_before, n_before, c->params.initial_alpha, d_ain, n, c->stream));
    return VKRT_OK;
}

int vkrt_partial_color(VkrtContext* c, const VkrtCameraUniform* cam, const VkrtUniform* un, const float* d_ain, float* d_rgba) {
    if (!c || !cam || !un || !d_ain || !d_rgba) return fail(VKRT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    PartialArgs A;
    int rc = fill_partial_args(c, cam, A);
    if (rc) return rc;
    A.a_in = d_ain;
    A.rgba_out = reinterpret_cast<float4*>(d_rgba);
    CK(launch_partial(A, c->params.mode, c->dtype, 2, c->stream));
    return VKRT_OK;
}

int vkrt_partial_relative(VkrtContext* c, const VkrtCameraUniform* cam, const VkrtUniform* un, float* d_rgba, float* d_T) {
    if (!c || !cam || !un || !d_rgba || !d_T) return fail(VKRT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    PartialArgs A;
    int rc = fill_partial_args(c, cam, A);
    if (rc) return rc;
    A.a_in = nullptr;  // relative pass
    A.rgba_out = reinterpret_cast<float4*>(d_rgba);
    A.T_out = d_T;
    CK(launch_partial(A, c->params.mode, c->dtype, 2, c->stream));
    return VKRT_OK;
}

int vkrt_partial_resolve(VkrtContext* c, const float* d_T_all, const int* ranks_before, int n_before, float* d_rgba, float* d_ain) {
    if (!c || !d_rgba || !d_ain || n_before < 0 || (n_before > 0 && (!d_T_all || !ranks_before))) return fail(VKRT_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(c->device));
    if (!c->d_before) CK(cudaMalloc(&c->d_before, 1024 * sizeof(int)));
    if (n_before > 1024) return fail(VKRT_ERR_INVALID, "too many ranks");
    if (n_before > 0) CK(cudaMemcpyAsync(c->d_before, ranks_before, (size_t)n_before * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    const size_t n = (size_t)c->W * c->H;
    CK(launch_partial_resolve(d_T_all, n, c->d_before, n_before, c->params.initial_alpha, c->params.alpha_threshold,
                              reinterpret_cast<float4*>(d_rgba), d_ain, n, c->stream));
    return VKRT_OK;
}

int vkrt_partial_finalize(VkrtContext* c, const VkrtCameraUniform* cam, const VkrtUniform* un, const float* d_sum_rgba) {
    if (!c || !cam || !un || !d_sum_rgba) return fail(VKRT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    PartialArgs A;
    int rc = fill_partial_args(c, cam, A);
    if (rc) return rc;
    CK(launch_partial_finalize(A, reinterpret_cast<const float4*>(d_sum_rgba), c->frame, c->params.mode, c->params.m1_srgb, c->stream));
    return VKRT_OK;
}

// ---- sort-last direct-send exchange (helpers: above, before the extern "C" block) ----
int vkrt_exchange_create(VkrtContext* c, int rank, int world, VkrtExchangeHandle* out) {
    if (!c || !out || world < 1 || world > VkrtContext::kXMaxWorld || rank < 0 || rank >= world) return fail(VKRT_ERR_INVALID, "bad exchange arguments (world <= 16)");
    if ((x_image_floats(c) & 3u) != 0) return fail(VKRT_ERR_UNSUPPORTED, "the exchange moves 16-byte vectors: W*H must be a multiple of 4");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    x_release(c);
    c->x_world = world;
    c->x_rank = rank;
    const size_t bytes = VkrtContext::kXMailbox + 2 * (size_t)world * x_image_floats(c) * sizeof(float);
    cudaError_t e = cudaMalloc(&c->x_base, bytes);
    if (e != cudaSuccess) { x_release(c); return cuda_fail(e, "vkrt_exchange_create"); }
    CK(cudaMemsetAsync(c->x_base, 0, VkrtContext::kXMailbox, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // completed before the handle leaves (see vkrt_sortfirst_create_root)
    c->x_peer[rank] = c->x_base;
    memset(out, 0, sizeof *out);
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->x_base));
    static_assert(sizeof(h) <= sizeof(out->ipc), "IPC handle size");
    memcpy(out->ipc, &h, sizeof h);
    out->width = c->W; out->height = c->H; out->world = world; out->rank = rank;
    return VKRT_OK;
}

int vkrt_exchange_open(VkrtContext* c, const VkrtExchangeHandle* all) {
    if (!c || !all || !c->x_base) return fail(VKRT_ERR_INVALID, "vkrt_exchange_create first");
    CK(cudaSetDevice(c->device));
    for (int r = 0; r < c->x_world; ++r) {
        if (r == c->x_rank) continue;
        if (all[r].width != c->W || all[r].height != c->H || all[r].world != c->x_world || all[r].rank != r)
            return fail(VKRT_ERR_INVALID, "exchange handle does not match this group (frame size / world / rank)");
        cudaIpcMemHandle_t ih;
        memcpy(&ih, all[r].ipc, sizeof ih);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
        c->x_peer[r] = (unsigned char*)p;
    }
    return VKRT_OK;
}

int vkrt_exchange_push(VkrtContext* c, const float* d_T, const int* ranks_behind, int n_behind, uint64_t frame) {
    if (!c || !c->x_base || !d_T || n_behind < 0 || (n_behind > 0 && !ranks_behind) || n_behind > kMaxPushDst) return fail(VKRT_ERR_INVALID, "bad exchange_push arguments");
    CK(cudaSetDevice(c->device));
    PushDst tables{}, arrive{}, done{};
    for (int k = 0; k < n_behind; ++k) {
        const int r = ranks_behind[k];
        if (r < 0 || r >= c->x_world || r == c->x_rank || !c->x_peer[r]) return fail(VKRT_ERR_INVALID, "exchange_push: no such peer (vkrt_exchange_open?)");
        tables.ptr[k] = x_table(c, c->x_peer[r], frame) + (size_t)c->x_rank * x_image_floats(c);
        arrive.ptr[k] = x_flag(c->x_peer[r], (int)(frame & 1u));
        done.ptr[k] = x_flag(c->x_peer[r], 2);
    }
    tables.n = arrive.n = done.n = n_behind;
    // the receivers must have resolved frame - 2 (same parity) before its table is overwritten
    if (frame >= 2) CK(launch_flag_wait_many(done, frame - 1, x_flag(c->x_base, 3), c->stream));
    CK(launch_push_many(d_T, tables, x_image_floats(c), c->stream));
    CK(launch_flag_add_many(arrive, 1ull, c->stream));
    return VKRT_OK;
}

int vkrt_exchange_wait(VkrtContext* c, uint64_t frame, int arrivals) {
    if (!c || !c->x_base || arrivals < 0) return fail(VKRT_ERR_INVALID, "bad exchange_wait arguments");
    CK(cudaSetDevice(c->device));
    c->x_expected[frame & 1u] += (unsigned long long)arrivals;
    if (arrivals > 0) CK(launch_flag_wait(x_flag(c->x_base, (int)(frame & 1u)), c->x_expected[frame & 1u], x_flag(c->x_base, 3), c->stream));
    return VKRT_OK;
}

const float* vkrt_exchange_table(VkrtContext* c, uint64_t frame) {
    if (!c || !c->x_base) return nullptr;
    return x_table(c, c->x_base, frame);
}

int vkrt_exchange_done(VkrtContext* c, uint64_t frame) {
    if (!c || !c->x_base) return fail(VKRT_ERR_INVALID, "no exchange");
    CK(cudaSetDevice(c->device));
    CK(launch_flag_set(x_flag(c->x_base, 2), (unsigned long long)frame + 1ull, c->stream));
    return VKRT_OK;
}

int vkrt_exchange_timeouts(VkrtContext* c, uint64_t* out) {
    if (!c || !out || !c->x_base) return fail(VKRT_ERR_INVALID, "no exchange");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    unsigned long long v = 0;
    CK(cudaMemcpy(&v, x_flag(c->x_base, 3), sizeof v, cudaMemcpyDeviceToHost));
    *out = v;
    return VKRT_OK;
}

int vkrt_exchange_close(VkrtContext* c) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    x_release(c);
    return VKRT_OK;
}

int vkrt_mark(VkrtContext* c, int idx) {
    if (!c || idx < 0 || idx >= 8) return fail(VKRT_ERR_INVALID, "mark index must be 0..7");
    CK(cudaSetDevice(c->device));
    if (!c->marks[idx]) CK(cudaEventCreate(&c->marks[idx]));
    CK(cudaEventRecord(c->marks[idx], c->stream));
    return VKRT_OK;
}

int vkrt_mark_elapsed(VkrtContext* c, int from, int to, float* ms) {
    if (!c || !ms || from < 0 || from >= 8 || to < 0 || to >= 8 || !c->marks[from] || !c->marks[to]) return fail(VKRT_ERR_INVALID, "bad marks");
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->marks[to]));
    CK(cudaEventElapsedTime(ms, c->marks[from], c->marks[to]));
    return VKRT_OK;
}

int vkrt_set_occupancy_brick(VkrtContext* c, int edge) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    int bs = -1;
    if (edge != 0) {
        for (bs = 0; bs <= 5 && (1 << bs) != edge; ++bs) {}
        if (bs > 5) return fail(VKRT_ERR_INVALID, "occupancy brick edge must be 0 (automatic), 1, 2, 4, 8, 16 or 32");
    }
    c->obs_request = bs;
    return VKRT_OK;
}

int vkrt_volume_info(VkrtContext* c, int* kind, int* dtype, int dims[3], uint64_t* bricks_total, uint64_t* bricks_occupied) {
    if (!c) return fail(VKRT_ERR_INVALID, "ctx is NULL");
    if (kind) *kind = c->kind;
    if (dtype) *dtype = c->dtype;
    if (dims) { dims[0] = c->nx; dims[1] = c->ny; dims[2] = c->nz; }
    const size_t cells = c->windowed ? (size_t)c->cell_n[0] * c->cell_n[1] * c->cell_n[2] : (size_t)c->obx * c->oby * c->obz;  // the grid c->dist covers
    if (bricks_total) *bricks_total = cells;
    if (bricks_occupied) {
        *bricks_occupied = 0;
        if (c->dist) {
            CK(cudaSetDevice(c->device));
            CK(cudaStreamSynchronize(c->stream));
            std::vector<uint8_t> h(cells);
            CK(cudaMemcpy(h.data(), c->dist, cells, cudaMemcpyDeviceToHost));
            uint64_t n = 0;
            for (uint8_t d : h) n += d == 0;
            *bricks_occupied = n;
        }
    }
    return VKRT_OK;
}

uint32_t vkrt_dispatch_optimal(uint32_t len, uint32_t subgroup_size) {
    // src/utils/mod.rs:15-18
    if (subgroup_size == 0) return 0;
    const uint32_t padded = (subgroup_size - len % subgroup_size) % subgroup_size;
    return (len + padded) / subgroup_size;
}

}  // extern "C"
