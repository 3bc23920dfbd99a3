// headless.cpp — offscreen harness next to the reference's windowed presenter (src/lib.rs:45-208).
// Runs the xor example's Demo (examples/xor/main.rs) without a window: generate the volume once,
// then render `frames` frames of an orbit, print the mean frame time, optionally dump the last frame.
// Mode `sweep` hands the whole orbit to the library in chunks of 24 cameras (Context::capture_sweep ->
// vkrt_frames_host): several frames per launch, present fused, D2H pipelined — the offline-recording path.
// Mode `drag` replays a scripted mouse drag + wheel through OrbitInput (the event mapping of src/lib.rs:150-176) and
// renders ONE frame with the resulting camera: 40 motion events of (+8, -3) pixels with the button held, 5 without,
// two wheel lines up.
//   usage: headless [frames=360] [W=1280] [H=720] [single|tile|sweep|drag] [out.rgba8]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "vokselis.hpp"

using namespace vokselis;

struct Xor : Demo {  // examples/xor/main.rs:34-48,50-162,210-262
    XorCompute xor_texture{256, 0};
    RaycastPipeline raycast_single{"single"}, raycast_tile{"tile"};
    bool tile_mode = false;
    void init(Context& ctx) override { xor_texture.record(ctx); }
    void render(Context& ctx) override { (tile_mode ? raycast_tile : raycast_single).record(ctx); }
};

int main(int argc, char** argv) {
    const unsigned frames = argc > 1 ? (unsigned)atoi(argv[1]) : 360u;
    const unsigned W = argc > 2 ? (unsigned)atoi(argv[2]) : HdrBackBuffer::DEFAULT_WIDTH;
    const unsigned H = argc > 3 ? (unsigned)atoi(argv[3]) : HdrBackBuffer::DEFAULT_HEIGHT;
    try {
        Context ctx(0, W, H);
        Xor demo;
        demo.tile_mode = argc > 4 && strcmp(argv[4], "tile") == 0;
        if (argc > 4 && strcmp(argv[4], "sweep") == 0) {
            demo.init(ctx);
            std::vector<VkrtCameraUniform> cams;
            std::vector<uint8_t> last;
            const auto t0 = std::chrono::steady_clock::now();
            for (unsigned i = 0; i < frames; ++i) {
                ctx.camera.set_yaw(1.0f + 6.2831853f * (float)i / (float)(frames ? frames : 1));
                cams.push_back(ctx.camera.get_proj_view_matrix());
                if (cams.size() == 24 || i + 1 == frames) {
                    last = ctx.capture_sweep(cams);  // presented RGBA8 frames, in order
                    cams.clear();
                }
            }
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("%u frames %ux%u (sweep, frames delivered to host memory) in %.3f s -> %.1f frames/s\n", frames, W, H, s, frames / s);
            if (argc > 5 && !last.empty()) {
                const size_t one = (size_t)W * H * 4;
                FILE* f = fopen(argv[5], "wb");
                if (f) { fwrite(last.data() + last.size() - one, 1, one, f); fclose(f); }
            }
            return 0;
        }
        if (argc > 4 && strcmp(argv[4], "drag") == 0) {
            OrbitInput input;
            input.button(true);
            for (int i = 0; i < 40; ++i) input.mouse_motion(ctx.camera, 8.0, -3.0);
            input.button(false);
            for (int i = 0; i < 5; ++i) input.mouse_motion(ctx.camera, 100.0, 100.0);  // not dragging: ignored
            input.mouse_wheel_lines(ctx.camera, 2.0f);
            run_headless(demo, ctx, 1, [](Context&, unsigned) {});
            printf("drag: yaw %.9g pitch %.9g zoom %.9g\n", ctx.camera.yaw, ctx.camera.pitch, ctx.camera.zoom);
            if (argc > 5) {
                const auto px = ctx.capture_frame();
                FILE* f = fopen(argv[5], "wb");
                if (f) { fwrite(px.data(), 1, px.size(), f); fclose(f); }
            }
            return 0;
        }
        const auto t0 = std::chrono::steady_clock::now();
        run_headless(demo, ctx, frames, [&](Context& c, unsigned i) {
            c.camera.set_yaw(1.0f + 6.2831853f * (float)i / (float)(frames ? frames : 1));
        });
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("%u frames %ux%u (%s) in %.3f s -> %.1f frames/s (includes volume generation)\n", frames, W, H,
               demo.tile_mode ? "tile" : "single", s, frames / s);
        if (argc > 5) {
            const auto px = ctx.capture_frame();
            FILE* f = fopen(argv[5], "wb");
            if (f) { fwrite(px.data(), 1, px.size(), f); fclose(f); }
        }
    } catch (const Error& e) {
        fprintf(stderr, "vokselis error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
