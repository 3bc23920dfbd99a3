// vokselis.hpp — C++ host-side mirror of the reference's Rust interface for the raycast path.
//
// The reference host is Rust (src/lib.rs:13-18 re-exports). This image has no Rust toolchain, so
// the host layer above the C ABI (include/vokselis_rt.h) is C++ with the reference's names, argument
// meaning and error behaviour; INTEGRATION.md shows the Rust `extern "C"` binding a maintainer adds.
//
//   reference (Rust)                                   here (C++)
//   Camera, CameraUniform      src/camera.rs           vokselis::Camera, VkrtCameraUniform
//   Uniform                    src/context/global_ubo.rs  VkrtUniform, vokselis::default_uniform()
//   HdrBackBuffer              src/context/hdr_backbuffer.rs  vokselis::HdrBackBuffer (size constants)
//   Context                    src/context.rs          vokselis::Context (device, stream, frame, camera, uniform)
//   XorCompute                 examples/xor/xor_compute.rs    vokselis::XorCompute
//   VolumeTexture              src/context/volume_texture.rs  vokselis::VolumeTexture
//   RaycastPipeline            examples/xor/raycast.rs        vokselis::RaycastPipeline ("single" | "tile")
//   trait Demo, run()          src/lib.rs:37-45        vokselis::Demo, vokselis::run_headless()
//   mouse / wheel -> camera    src/lib.rs:64-66,150-176  vokselis::OrbitInput
//   dispatch_optimal           src/utils/mod.rs:15-18  vokselis::dispatch_optimal
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vokselis_rt.h"

namespace vokselis {

struct VKRT_API Mat4 {
    float c[16];  // column-major, like glam::Mat4::to_cols_array
    static Mat4 identity();
    static Mat4 look_at_rh(const float eye[3], const float center[3], const float up[3]);
    static Mat4 perspective_rh(float fov_y, float aspect, float z_near, float z_far);
    Mat4 operator*(const Mat4& rhs) const;
    Mat4 inverse() const;
};

// src/camera.rs:75-171
struct VKRT_API Camera {
    static constexpr float ZFAR = 100.0f;                                 // :88
    static constexpr float ZNEAR = 0.1f;                                  // :89
    static constexpr float FOVY = 3.14159265358979323846f / 2.0f;         // :90
    float zoom, pitch, yaw, aspect;
    float target[3], eye[3], up[3];
    bool updated;

    Camera(float zoom, float pitch, float yaw, const float target[3], float aspect);
    Mat4 build_projection_view_matrix() const;
    void set_zoom(float zoom);
    void add_zoom(float delta);
    void set_pitch(float pitch);
    void add_pitch(float delta);
    void set_yaw(float yaw);
    void add_yaw(float delta);
    void set_aspect(unsigned width, unsigned height);
    VkrtCameraUniform get_proj_view_matrix() const;

   private:
    void fix_eye();
};

// The event -> camera mapping of `run` (src/lib.rs:64-66,150-176): dragging with the mouse button held turns the orbit by
// 0.0025 rad per pixel (yaw against the x motion, pitch with the y motion), the wheel zooms by 0.002 per line / pixel
// (scrolling up moves in). Windowing itself is out of scope; a host that has events feeds them here.
struct OrbitInput {
    static constexpr float ROTATE_SPEED = 0.0025f;  // src/lib.rs:65
    static constexpr float ZOOM_SPEED = 0.002f;     // src/lib.rs:66
    bool mouse_dragged = false;                     // src/lib.rs:64
    void button(bool pressed) { mouse_dragged = pressed; }                                                   // :150-159
    void mouse_wheel_lines(Camera& cam, float scroll) const { cam.add_zoom(-(scroll * 1.0f) * ZOOM_SPEED); }  // :160-168, LineDelta
    void mouse_wheel_pixels(Camera& cam, double scroll_y) const { cam.add_zoom(-(float)scroll_y * ZOOM_SPEED); }  // PixelDelta
    void mouse_motion(Camera& cam, double dx, double dy) const {                                             // :169-174
        if (!mouse_dragged) return;
        cam.add_yaw(-(float)dx * ROTATE_SPEED);
        cam.add_pitch((float)dy * ROTATE_SPEED);
    }
};

// `impl Default for Uniform` (src/context/global_ubo.rs:67-81)
inline VkrtUniform default_uniform() {
    VkrtUniform u{};
    u.resolution[0] = 1920.0f;
    u.resolution[1] = 780.0f;
    u.time_delta = 1.0f / 60.0f;
    return u;
}

struct HdrBackBuffer {
    static constexpr unsigned DEFAULT_WIDTH = 1280, DEFAULT_HEIGHT = 720;  // hdr_backbuffer.rs:11
};

inline uint32_t dispatch_optimal(uint32_t len, uint32_t subgroup_size) { return vkrt_dispatch_optimal(len, subgroup_size); }

// The reference panics/unwraps on failure (examples/xor/raycast.rs:41) or returns eyre::Result
// (src/context.rs:71-75); the C++ mirror throws vokselis::Error carrying vkrt_last_error().
struct Error : std::runtime_error {
    int code;
    Error(int code_, const std::string& what) : std::runtime_error(what), code(code_) {}
};
inline void check(int rc) {
    if (rc != VKRT_OK) throw Error(rc, vkrt_last_error());
}

// src/context.rs:39-69 — what the raycast path needs from `Context`.
class Context {
   public:
    Context(int device, unsigned width = HdrBackBuffer::DEFAULT_WIDTH, unsigned height = HdrBackBuffer::DEFAULT_HEIGHT,
            const Camera* cam = nullptr)
        : camera(cam ? *cam : default_camera(width, height)), global_uniform(default_uniform()), width_(width), height_(height) {
        check(vkrt_create(device, (int)width, (int)height, &ctx_));
        // The reference leaves the GPU camera buffer as identity until the first `updated`;
        // a headless host uploads explicitly, so mark it dirty.
        camera.updated = true;
        update();
    }
    ~Context() { vkrt_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;

    // src/context.rs:225-236 `update`: refresh the per-frame uniforms; camera only when dirty.
    void update() {
        global_uniform.frame += 1;
        if (camera.updated) {
            camera_uniform = camera.get_proj_view_matrix();
            camera.updated = false;
        }
    }
    // src/context.rs:238-249 `resize`
    void resize(unsigned width, unsigned height) {
        check(vkrt_resize(ctx_, (int)width, (int)height));
        width_ = width;
        height_ = height;
        camera.set_aspect(width, height);
    }
    // src/context.rs:251-297 `render`: the present pass (tonemap into the Rgba8 capture target)
    void render() { check(vkrt_present(ctx_)); }
    // No reference counterpart: brick edge (voxels) of the occupancy grid behind exact empty-space skipping, for volumes
    // created after the call; 0 = the library's choice. Frames do not depend on it.
    void set_occupancy_brick(int edge) { check(vkrt_set_occupancy_brick(ctx_, edge)); }
    // src/context.rs:299-306 `capture_frame`
    std::vector<uint8_t> capture_frame() {
        std::vector<uint8_t> px((size_t)width_ * height_ * 4);
        check(vkrt_readback_rgba8(ctx_, px.data()));
        return px;
    }
    // A whole camera sweep for a host consumer — what the recorder does frame by frame (src/lib.rs:132-140,
    // src/utils/recorder.rs:79-127) — in one call: groups of frames per launch, present fused, D2H pipelined.
    // Returns cams.size() tightly packed RGBA8 frames.
    std::vector<uint8_t> capture_sweep(const std::vector<VkrtCameraUniform>& cams) {
        std::vector<uint8_t> px((size_t)width_ * height_ * 4 * cams.size());
        if (!cams.empty()) check(vkrt_frames_host(ctx_, cams.data(), (int)cams.size(), &global_uniform, px.data(), 0));
        return px;
    }
    std::vector<uint16_t> read_backbuffer() {
        std::vector<uint16_t> px((size_t)width_ * height_ * 4);
        check(vkrt_readback(ctx_, px.data()));
        return px;
    }
    VkrtContext* raw() const { return ctx_; }
    unsigned width() const { return width_; }
    unsigned height() const { return height_; }

    Camera camera;
    VkrtCameraUniform camera_uniform{};
    VkrtUniform global_uniform;

   private:
    static Camera default_camera(unsigned w, unsigned h) {
        const float t[3] = {0.0f, 0.0f, 0.0f};
        return Camera(3.0f, -0.5f, 1.0f, t, (float)w / (float)h);  // examples/xor/main.rs:273-279
    }
    VkrtContext* ctx_ = nullptr;
    unsigned width_, height_;
};

// examples/xor/xor_compute.rs — the two rgba16f volumes and the generator dispatch.
class XorCompute {
   public:
    explicit XorCompute(unsigned n = 256, int which = 0) : n_(n), which_(which) {}
    // `record` (:188-200): dispatch cs_main with the context's global uniform
    void record(Context& ctx) const { check(vkrt_generate_xor(ctx.raw(), &ctx.global_uniform, (int)n_, which_)); }

   private:
    unsigned n_;
    int which_;
};

// src/context/volume_texture.rs — a tightly packed scalar grid (the bonsai loader's layout contract).
class VolumeTexture {
   public:
    VolumeTexture(Context& ctx, const void* data, VkrtDtype dtype, unsigned nx, unsigned ny, unsigned nz) {
        check(vkrt_upload_scalar(ctx.raw(), data, dtype, (int)nx, (int)ny, (int)nz));
    }
};

// examples/xor/raycast.rs — pipeline object with an entry point name.
class RaycastPipeline {
   public:
    explicit RaycastPipeline(const std::string& entry_point) : tile_(entry_point == "tile") {
        if (entry_point != "single" && entry_point != "tile") throw Error(VKRT_ERR_INVALID, "unknown entry point: " + entry_point);
    }
    // What `Xor::render` records (examples/xor/main.rs:223-254).
    void record(Context& ctx) const {
        if (!tile_) {
            check(vkrt_render(ctx.raw(), &ctx.camera_uniform, &ctx.global_uniform, nullptr));
            return;
        }
        VkrtParams p;
        check(vkrt_get_params(ctx.raw(), &p));
        const int n = vkrt_tile_table((int)ctx.width(), (int)ctx.height(), p.tile_size, nullptr, 0);
        std::vector<VkrtOffset> table((size_t)n);
        vkrt_tile_table((int)ctx.width(), (int)ctx.height(), p.tile_size, table.data(), n);
        check(vkrt_render_tiles(ctx.raw(), &ctx.camera_uniform, &ctx.global_uniform, table.data(), n));
    }

   private:
    bool tile_;
};

// src/lib.rs:37-43 `trait Demo`
struct Demo {
    virtual ~Demo() = default;
    virtual void init(Context&) = 0;
    virtual void update(Context&) {}
    virtual void render(Context&) = 0;
};

// Headless counterpart of `run` (src/lib.rs:45-208): no window, no input; `frames` iterations of
// update -> demo.update -> demo.render -> ctx.render, with an optional per-frame camera callback.
template <class OnFrame>
inline void run_headless(Demo& demo, Context& ctx, unsigned frames, OnFrame&& on_frame) {
    demo.init(ctx);
    for (unsigned i = 0; i < frames; ++i) {
        on_frame(ctx, i);
        ctx.update();
        demo.update(ctx);
        demo.render(ctx);
        ctx.render();
    }
    check(vkrt_sync(ctx.raw()));
}

}  // namespace vokselis
