//! Raw bindings of include/zenwebp_b200.h.  The only `unsafe` surface of the Rust side.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct zw_ctx { _private: [u8; 0] }

#[repr(C)]
pub struct zw_limits { pub max_device_bytes: usize, pub persistent_warps_per_sm: c_int, pub reserved: [c_int; 5] }

#[repr(C)]
pub struct zw_image { pub data: *const u8, pub len: usize, pub width: u32, pub height: u32, pub color: u32, pub reserved: u32 }

#[repr(C)]
pub struct zw_output { pub data: *mut u8, pub cap: usize, pub len: usize, pub status: c_int, pub reserved: c_int }

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct zw_timing {
    pub h2d_ms: f32, pub yuv_ms: f32, pub analysis_ms: f32, pub pass1_ms: f32, pub stats_ms: f32, pub pass2_ms: f32,
    pub token_ms: f32, pub boolcode_ms: f32, pub assemble_ms: f32, pub d2h_ms: f32, pub device_total_ms: f32, pub wall_ms: f32,
    pub kernel_launches: u64, pub h2d_bytes: u64, pub d2h_bytes: u64, pub pixels: u64,
    pub chroma1_ms: f32, pub chroma2_ms: f32,
}

pub const ZW_COLOR_RGB8: u32 = 2;
pub const ZW_COLOR_RGBA8: u32 = 3;

extern "C" {
    pub fn zw_create(device: c_int, limits: *const zw_limits) -> *mut zw_ctx;
    pub fn zw_destroy(ctx: *mut zw_ctx);
    pub fn zw_last_error() -> c_int;
    pub fn zw_strerror(code: c_int) -> *const c_char;
    pub fn zw_free(p: *mut c_void);
    pub fn zw_max_output_size(width: u32, height: u32) -> usize;
    pub fn zw_encode_vp8_batch(ctx: *mut zw_ctx, imgs: *const zw_image, n: usize, quality: c_int, method: c_int,
                               outs: *mut zw_output, timing: *mut zw_timing) -> c_int;
    pub fn zw_encode_webp_batch(ctx: *mut zw_ctx, imgs: *const zw_image, n: usize, quality: c_int, method: c_int,
                                outs: *mut zw_output, timing: *mut zw_timing) -> c_int;
    pub fn zw_stage_batch(ctx: *mut zw_ctx, imgs: *const zw_image, n: usize) -> c_int;
    pub fn zw_encode_resident(ctx: *mut zw_ctx, quality: c_int, method: c_int, timing: *mut zw_timing) -> c_int;
    pub fn zw_download(ctx: *mut zw_ctx, outs: *mut zw_output, n: usize, container: c_int, timing: *mut zw_timing) -> c_int;
    pub fn zw_dump_stage(ctx: *mut zw_ctx, index: usize, stage: *const c_char, dst: *mut c_void, cap: usize, len: *mut usize) -> c_int;
    pub fn zw_version() -> *const c_char;
    pub fn zw_measure_int_peak(ctx: *mut zw_ctx, int_instr_per_s: *mut f64) -> c_int;
}
