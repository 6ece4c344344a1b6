//! Thin `extern "C"` layer over include/stacker_cuda.h (ABI version 5).  One declaration per entry point the
//! Rust wrappers use; see the header for ownership and threading rules.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const STK_OK: c_int = 0;
pub const STK_ERR_BAD_ARG: c_int = 1;
pub const STK_ERR_CUDA: c_int = 2;
pub const STK_ERR_NOT_ENOUGH: c_int = 3;
pub const STK_ERR_ECC_NOCONV: c_int = 4;
pub const STK_ERR_ECC_NAN: c_int = 5;
pub const STK_ERR_CRITERIA: c_int = 6;
pub const STK_ERR_STATE: c_int = 7;
pub const STK_ERR_UNSUPPORTED: c_int = 8;
pub const STK_ERR_NOMEM: c_int = 9;

#[repr(C)]
pub struct stk_ecc_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct stk_ecc_config {
    pub width: i32,
    pub height: i32,
    pub channels: i32,
    pub motion_type: i32,
    pub criteria_type: i32,
    pub max_count: i32,
    pub epsilon: f64,
    pub gauss_filt_size: i32,
    pub device: i32,
    pub lanes: i32,
    pub seed_reference: i32,
    pub align: i32,
    pub ecc_width: i32,
    pub ecc_height: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct stk_frame_result {
    pub tag: i64,
    pub warp: [f32; 9],
    pub rho: f64,
    pub iterations: i32,
    pub status: i32,
}

unsafe extern "C" {
    pub fn stk_abi_version() -> c_int;
    pub fn stk_last_error() -> *const c_char;
    pub fn stk_device_count(count: *mut c_int) -> c_int;
    pub fn stk_scaled_size(width: c_int, height: c_int, scale_down: f32, sw: *mut c_int, sh: *mut c_int) -> c_int;
    pub fn stk_pinned_alloc(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn stk_pinned_free(ptr: *mut c_void) -> c_int;

    pub fn stk_ecc_create(cfg: *const stk_ecc_config, out: *mut *mut stk_ecc_ctx) -> c_int;
    pub fn stk_ecc_destroy(ctx: *mut stk_ecc_ctx) -> c_int;
    pub fn stk_ecc_set_reference(ctx: *mut stk_ecc_ctx, bgr: *const u8, pitch: usize) -> c_int;
    /// ABI v5: order every later `*_device` submission behind `cuda_stream` (a `cudaStream_t` of the context's device).
    pub fn stk_ecc_set_input_stream(ctx: *mut stk_ecc_ctx, cuda_stream: *mut c_void, enabled: c_int) -> c_int;
    pub fn stk_ecc_submit_frame(ctx: *mut stk_ecc_ctx, bgr: *const u8, pitch: usize, tag: i64) -> c_int;
    pub fn stk_ecc_submit_frame_pinned(ctx: *mut stk_ecc_ctx, bgr: *const u8, pitch: usize, tag: i64) -> c_int;
    pub fn stk_ecc_acquire_frame_buffer(ctx: *mut stk_ecc_ctx, buf: *mut *mut u8, pitch: *mut usize) -> c_int;
    pub fn stk_ecc_submit_acquired(ctx: *mut stk_ecc_ctx, buf: *mut u8, tag: i64) -> c_int;
    pub fn stk_ecc_release_frame_buffer(ctx: *mut stk_ecc_ctx, buf: *mut u8) -> c_int;
    pub fn stk_ecc_submit_warp(
        ctx: *mut stk_ecc_ctx, bgr: *const u8, pitch: usize, h: *const f64, border_mode: c_int,
        border_value: *const f64, tag: i64,
    ) -> c_int;
    pub fn stk_ecc_submit_warp_affine(
        ctx: *mut stk_ecc_ctx, bgr: *const u8, pitch: usize, m: *const f64, border_mode: c_int,
        border_value: *const f64, tag: i64,
    ) -> c_int;
    pub fn stk_ecc_sync(ctx: *mut stk_ecc_ctx) -> c_int;
    pub fn stk_ecc_results(ctx: *mut stk_ecc_ctx, out: *mut stk_frame_result, capacity: c_int, count: *mut c_int) -> c_int;
    pub fn stk_ecc_finish(ctx: *mut stk_ecc_ctx, divisor: c_int, out: *mut f32, out_pitch: usize) -> c_int;
    pub fn stk_ecc_partial(ctx: *mut stk_ecc_ctx, d_partial: *mut *mut f32, n_floats: *mut usize) -> c_int;
    pub fn stk_ecc_finish_from(ctx: *mut stk_ecc_ctx, d_sum: *const f32, divisor: c_int, out: *mut f32, out_pitch: usize) -> c_int;
    pub fn stk_ecc_reset(ctx: *mut stk_ecc_ctx) -> c_int;

    // multi-GPU exchange over NVLink peer memory, all contexts in this process (include/stacker_cuda.h)
    pub fn stk_ecc_peer_connect_local(ctxs: *const *mut stk_ecc_ctx, world: c_int) -> c_int;
    pub fn stk_ecc_peer_reduce(ctx: *mut stk_ecc_ctx, divisor: c_int, d_out: *mut *const f32) -> c_int;
    pub fn stk_ecc_peer_reduce_scatter(
        ctx: *mut stk_ecc_ctx, divisor: c_int, d_slice: *mut *const f32, begin: *mut usize, count: *mut usize,
    ) -> c_int;
    pub fn stk_ecc_peer_slice_to_host(ctx: *mut stk_ecc_ctx, out: *mut f32) -> c_int;
    pub fn stk_ecc_peer_disconnect(ctx: *mut stk_ecc_ctx) -> c_int;

    pub fn stk_tenengrad(
        img: *const u8, pitch: usize, width: c_int, height: c_int, channels: c_int, ksize: c_int,
        device: c_int, out: *mut f64,
    ) -> c_int;
    pub fn stk_sharpness_all(
        img: *const u8, pitch: usize, width: c_int, height: c_int, channels: c_int, device: c_int, out: *mut f64,
    ) -> c_int;
}

/// RAII handle; `Send + Sync` because every entry point taking a context is internally locked.
pub struct Ctx(pub *mut stk_ecc_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { stk_ecc_destroy(self.0) };
    }
}

pub fn last_error() -> String {
    unsafe {
        let p = stk_last_error();
        if p.is_null() { String::new() } else { std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned() }
    }
}
