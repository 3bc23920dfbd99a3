// UNBUILT SOURCE (no Rust toolchain in this image): the binding of libvokselis_rt.so a maintainer adds to the
// reference crate as src/cuda_rt.rs. See INTEGRATION.md.
// Link: println!("cargo:rustc-link-lib=dylib=vokselis_rt"); println!("cargo:rustc-link-search=native=<dir>");
use crate::{camera::CameraUniform, context::Uniform};
use std::{ffi::CStr, os::raw::{c_char, c_int, c_void}};

#[repr(C)] #[derive(Clone, Copy)] pub struct Offset { pub x: f32, pub y: f32 }          // examples/xor/main.rs:20-25
#[repr(C)] #[derive(Clone, Copy)] pub struct VkrtParams {
    pub struct_size: u32, pub mode: i32, pub dt_scale: f32, pub dt_floor: f32, pub alpha_threshold: f32,
    pub initial_alpha: f32, pub clear_color: [f32; 4], pub tile_size: i32, pub layout: i32,
    pub skip_empty: i32, pub count_samples: i32, pub m1_srgb: i32, pub reserved: [i32; 7],
}
#[repr(C)] pub struct VkrtContext { _private: [u8; 0] }

extern "C" {
    fn vkrt_create(device: c_int, width: c_int, height: c_int, out: *mut *mut VkrtContext) -> c_int;
    fn vkrt_destroy(ctx: *mut VkrtContext) -> c_int;
    fn vkrt_last_error() -> *const c_char;
    fn vkrt_default_params(mode: c_int, out: *mut VkrtParams);
    fn vkrt_set_params(ctx: *mut VkrtContext, p: *const VkrtParams) -> c_int;
    fn vkrt_generate_xor(ctx: *mut VkrtContext, un: *const Uniform, n: c_int, which: c_int) -> c_int;
    fn vkrt_upload_rgba16f(ctx: *mut VkrtContext, color: *const u16, normal: *const u16, nx: c_int, ny: c_int, nz: c_int) -> c_int;
    fn vkrt_upload_scalar(ctx: *mut VkrtContext, data: *const c_void, dtype: c_int, nx: c_int, ny: c_int, nz: c_int) -> c_int;
    fn vkrt_render(ctx: *mut VkrtContext, cam: *const CameraUniform, un: *const Uniform, offset: *const Offset) -> c_int;
    fn vkrt_render_tiles(ctx: *mut VkrtContext, cam: *const CameraUniform, un: *const Uniform, offsets: *const Offset, n: c_int) -> c_int;
    fn vkrt_present(ctx: *mut VkrtContext) -> c_int;
    // present stretched onto a window of another size (src/context.rs:285-289 + present_pipeline.rs:110-118)
    fn vkrt_present_scaled(ctx: *mut VkrtContext, out_w: c_int, out_h: c_int, rgba8: *mut u8) -> c_int;
    fn vkrt_readback(ctx: *mut VkrtContext, rgba16f: *mut u16) -> c_int;
    fn vkrt_readback_rgba8(ctx: *mut VkrtContext, rgba8: *mut u8) -> c_int;
    fn vkrt_sync(ctx: *mut VkrtContext) -> c_int;
    // a sweep: n cameras -> n presented RGBA8 frames in host memory, groups of `group` frames per launch
    fn vkrt_frames_host(ctx: *mut VkrtContext, cams: *const CameraUniform, n: c_int, un: *const Uniform, rgba8: *mut u8, group: c_int) -> c_int;
    fn vkrt_alloc_host(bytes: usize, out: *mut *mut c_void) -> c_int;   // page-locked memory: keeps the D2H copies asynchronous
    fn vkrt_free_host(ptr: *mut c_void) -> c_int;
    // page-lock memory the caller owns (e.g. a shared-memory segment mapped by one process per GPU)
    fn vkrt_host_register(ptr: *mut c_void, bytes: usize) -> c_int;
    fn vkrt_host_unregister(ptr: *mut c_void) -> c_int;
    // tuning knob without a reference counterpart: brick edge of the occupancy grid (0 = automatic); frames do not depend on it
    fn vkrt_set_occupancy_brick(ctx: *mut VkrtContext, edge: c_int) -> c_int;
}

pub struct CudaRaycast { ctx: *mut VkrtContext, pub width: u32, pub height: u32 }

fn check(rc: c_int) -> color_eyre::eyre::Result<()> {
    if rc == 0 { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(vkrt_last_error()) }.to_string_lossy().into_owned();
    Err(color_eyre::eyre::eyre!("vokselis_rt error {rc}: {msg}"))
}

impl CudaRaycast {
    pub fn new(device: i32, width: u32, height: u32) -> color_eyre::eyre::Result<Self> {
        let mut ctx = std::ptr::null_mut();
        check(unsafe { vkrt_create(device, width as _, height as _, &mut ctx) })?;
        Ok(Self { ctx, width, height })
    }
    /// Brick edge (voxels) of the occupancy grid for volumes created after this call; 0 = the library's choice.
    pub fn set_occupancy_brick(&self, edge: i32) -> color_eyre::eyre::Result<()> {
        check(unsafe { vkrt_set_occupancy_brick(self.ctx, edge as c_int) })
    }
    pub fn generate_xor(&self, un: &Uniform) -> color_eyre::eyre::Result<()> {
        check(unsafe { vkrt_generate_xor(self.ctx, un, 256, 0) })
    }
    /// `single`: offset = None. `tile`: one call with the whole offset table.
    pub fn record(&self, cam: &CameraUniform, un: &Uniform, offsets: Option<&[Offset]>) -> color_eyre::eyre::Result<()> {
        check(unsafe { match offsets {
            None => vkrt_render(self.ctx, cam, un, std::ptr::null()),
            Some(t) => vkrt_render_tiles(self.ctx, cam, un, t.as_ptr(), t.len() as _),
        }})
    }
    pub fn capture_frame(&self) -> color_eyre::eyre::Result<Vec<u8>> {
        let mut px = vec![0u8; (self.width * self.height * 4) as usize];
        check(unsafe { vkrt_present(self.ctx) })?;
        check(unsafe { vkrt_readback_rgba8(self.ctx, px.as_mut_ptr()) })?;
        Ok(px)
    }
}
impl CudaRaycast {
    /// The recorder path (src/lib.rs:132-140, src/utils/recorder.rs:79-127) for a whole camera sweep: one call,
    /// frames come back presented (ACES + sRGB) as tightly packed RGBA8, frame i at i * width * height * 4.
    pub fn capture_sweep(&self, cams: &[CameraUniform], un: &Uniform, out: &mut [u8]) -> color_eyre::eyre::Result<()> {
        assert!(out.len() >= cams.len() * (self.width * self.height * 4) as usize);
        check(unsafe { vkrt_frames_host(self.ctx, cams.as_ptr(), cams.len() as _, un, out.as_mut_ptr(), 0) })
    }
}
impl CudaRaycast {
    /// What `Context::render` shows in a window that differs from the backbuffer (src/context.rs:285-289).
    pub fn present_to(&self, out_w: u32, out_h: u32) -> color_eyre::eyre::Result<Vec<u8>> {
        let mut px = vec![0u8; (out_w * out_h * 4) as usize];
        check(unsafe { vkrt_present_scaled(self.ctx, out_w as _, out_h as _, px.as_mut_ptr()) })?;
        Ok(px)
    }
}

/// The event -> camera mapping of `run` (src/lib.rs:64-66,150-176), kept as it is in the crate: the CUDA path only
/// consumes `camera.get_proj_view_matrix()`. Listed here because a headless host (no winit) has to feed it itself;
/// vokselis_b200/host/vokselis.hpp (`OrbitInput`) and vokselis_b200/rt.py carry the same mapping and tests.
pub struct OrbitInput { pub mouse_dragged: bool }
impl OrbitInput {
    pub const ROTATE_SPEED: f32 = 0.0025; // src/lib.rs:65
    pub const ZOOM_SPEED: f32 = 0.002;    // src/lib.rs:66
    pub fn button(&mut self, pressed: bool) { self.mouse_dragged = pressed; }
    pub fn mouse_wheel_lines(&self, cam: &mut crate::Camera, scroll: f32) { cam.add_zoom(-(scroll * 1.0) * Self::ZOOM_SPEED); }
    pub fn mouse_wheel_pixels(&self, cam: &mut crate::Camera, scroll_y: f64) { cam.add_zoom(-(scroll_y as f32) * Self::ZOOM_SPEED); }
    pub fn mouse_motion(&self, cam: &mut crate::Camera, dx: f64, dy: f64) {
        if self.mouse_dragged {
            cam.add_yaw(-dx as f32 * Self::ROTATE_SPEED);
            cam.add_pitch(dy as f32 * Self::ROTATE_SPEED);
        }
    }
}

impl Drop for CudaRaycast { fn drop(&mut self) { unsafe { vkrt_destroy(self.ctx); } } }
