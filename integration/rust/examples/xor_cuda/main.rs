// UNBUILT SOURCE (no Rust toolchain in this image): the xor example driving the CUDA raycaster; mirrors
// examples/xor/main.rs:34-262 of the reference. See INTEGRATION.md.
struct XorCuda { rt: CudaRaycast, offsets: Vec<Offset>, mode: Mode }

impl Demo for XorCuda {
    fn init(ctx: &mut vokselis::Context) -> Self {
        let (w, h) = HdrBackBuffer::DEFAULT_RESOLUTION;
        let rt = CudaRaycast::new(0, w, h).unwrap();          // like the reference's .unwrap() at raycast.rs:41
        rt.generate_xor(&ctx.global_uniform).unwrap();        // examples/xor/main.rs:135-146 (time = 0)
        let mut offsets = vec![];                              // examples/xor/main.rs:80-95
        for y in 0..(h / TILE_SIZE) + 1 { for x in 0..(w / TILE_SIZE) + 1 {
            offsets.push(Offset { x: (x * TILE_SIZE) as f32, y: (y * TILE_SIZE) as f32 });
        }}
        Self { rt, offsets, mode: Mode::SinglePass }
    }
    fn render(&mut self, ctx: &vokselis::Context) {
        let cam = ctx.camera.get_proj_view_matrix();          // src/camera.rs:164-171 (upload explicitly)
        let offsets = match self.mode { Mode::SinglePass => None, Mode::Tile => Some(&self.offsets[..]) };
        self.rt.record(&cam, &ctx.global_uniform, offsets).unwrap();
        // present: either rt.capture_frame() -> queue.write_texture(ctx.render_backbuffer ...), or keep
        // wgpu's present pass reading an imported external-memory texture.
    }
}
