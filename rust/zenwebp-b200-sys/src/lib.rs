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
    pub symbols: u64,
}

/// What `zw_wait` hands back: the files of one batch inside the slot's pinned arena (valid until `zw_release`).
#[repr(C)]
pub struct zw_batch_view { pub arena: *const u8, pub n: usize, pub offsets: *const u64, pub lens: *const u32, pub status: *const i32 }

#[repr(C)]
pub struct zw_multi { _private: [u8; 0] }
#[repr(C)]
#[derive(Clone, Copy)]
pub struct zw_blob { pub data: *const u8, pub len: usize }
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct zw_decode_info {
    pub status: i32, pub width: u32, pub height: u32, pub filter_type: u32, pub filter_level: u32, pub sharpness: u32,
    pub num_partitions: u32, pub segments_enabled: u32, pub sse_rgb: u64, pub psnr_rgb: f64,
}

pub const ZW_ERR_BUSY: c_int = 7;
pub const ZW_ERR_TOO_LARGE: c_int = 8;
pub const ZW_COLOR_L8: u32 = 0;
pub const ZW_COLOR_LA8: u32 = 1;
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
    pub fn zw_submit(ctx: *mut zw_ctx, imgs: *const zw_image, n: usize, quality: c_int, method: c_int, ticket: *mut c_int) -> c_int;
    pub fn zw_wait(ctx: *mut zw_ctx, ticket: c_int, container: c_int, view: *mut zw_batch_view, timing: *mut zw_timing) -> c_int;
    pub fn zw_release(ctx: *mut zw_ctx, ticket: c_int) -> c_int;
    pub fn zw_multi_create(devices: *const c_int, n_devices: c_int, limits: *const zw_limits) -> *mut zw_multi;
    pub fn zw_multi_destroy(m: *mut zw_multi);
    pub fn zw_multi_device_count(m: *const zw_multi) -> c_int;
    pub fn zw_multi_encode(m: *mut zw_multi, imgs: *const zw_image, n: usize, quality: c_int, method: c_int, container: c_int,
                           outs: *mut zw_output, per_device: *mut zw_timing) -> c_int;
    pub fn zw_stage_batch(ctx: *mut zw_ctx, imgs: *const zw_image, n: usize) -> c_int;
    pub fn zw_encode_resident(ctx: *mut zw_ctx, quality: c_int, method: c_int, timing: *mut zw_timing) -> c_int;
    pub fn zw_download(ctx: *mut zw_ctx, outs: *mut zw_output, n: usize, container: c_int, timing: *mut zw_timing) -> c_int;
    pub fn zw_dump_stage(ctx: *mut zw_ctx, index: usize, stage: *const c_char, dst: *mut c_void, cap: usize, len: *mut usize) -> c_int;
    pub fn zw_version() -> *const c_char;
    pub fn zw_measure_int_peak(ctx: *mut zw_ctx, int_instr_per_s: *mut f64) -> c_int;
    pub fn zw_decode_batch(ctx: *mut zw_ctx, files: *const zw_blob, n: usize, upsampling: c_int, rgb_outs: *mut zw_output,
                           sources: *const zw_image, infos: *mut zw_decode_info, device_ms: *mut f32) -> c_int;
    pub fn zw_verify(ctx: *mut zw_ctx, ticket: c_int, upsampling: c_int, infos: *mut zw_decode_info, device_ms: *mut f32) -> c_int;
    pub fn zw_encode_lossless_batch(ctx: *mut zw_ctx, imgs: *const zw_image, n: usize, use_predictor_transform: c_int, container: c_int,
                                    outs: *mut zw_output, timing: *mut zw_timing) -> c_int;
    pub fn zw_encode_alpha_batch(ctx: *mut zw_ctx, imgs: *const zw_image, n: usize, outs: *mut zw_output, timing: *mut zw_timing) -> c_int;
    pub fn zw_params_default() -> zw_params;
    pub fn zw_encode_batch(ctx: *mut zw_ctx, imgs: *const zw_image, n: usize, params: *const zw_params, meta: *const zw_metadata,
                           outs: *mut zw_output, timing: *mut zw_timing) -> c_int;
    pub fn zw_lossless_dump_stage(ctx: *mut zw_ctx, index: usize, stage: *const c_char, dst: *mut c_void, cap: usize, len: *mut usize) -> c_int;
    pub fn zw_decode_dump_stage(ctx: *mut zw_ctx, index: usize, stage: *const c_char, dst: *mut c_void, cap: usize, len: *mut usize) -> c_int;
}
