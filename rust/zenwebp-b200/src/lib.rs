//! Safe wrapper with the reference's names: `WebPEncoder`, `EncoderParams`, `ColorType`,
//! `EncodingError` (zenwebp 0.2.0 src/encoder/api.rs) + `encode_batch`.
//! NOTE: written but NOT compiled in the build image (no cargo/rustc there); see INTEGRATION.md.
#![forbid(unsafe_op_in_unsafe_fn)]
use zenwebp_b200_sys as sys;

#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum ColorType { L8, La8, Rgb8, Rgba8 }

#[derive(Debug, thiserror::Error)]
#[non_exhaustive]
pub enum EncodingError {
    #[error("Invalid dimensions")]
    InvalidDimensions,
    #[error("Invalid buffer size: {0}")]
    InvalidBufferSize(String),
    #[error("device error {0}")]
    Device(i32),
}

#[non_exhaustive]
#[derive(Clone, Debug)]
pub struct EncoderParams { pub use_predictor_transform: bool, pub use_lossy: bool, pub lossy_quality: u8, pub method: u8 }

impl Default for EncoderParams {
    fn default() -> Self { Self { use_predictor_transform: true, use_lossy: false, lossy_quality: 95, method: 4 } }
}
impl EncoderParams {
    pub fn lossless() -> Self { Self::default() }
    pub fn lossy(quality: u8) -> Self { Self { use_lossy: true, lossy_quality: quality, ..Self::default() } }
    /// Additive builder (the reference's README shows `.method(m)`; its struct only has the field).
    pub fn method(mut self, m: u8) -> Self { self.method = m; self }
}

pub struct ImageRef<'a> { pub data: &'a [u8], pub width: u32, pub height: u32, pub color: ColorType }

pub struct Context { h: *mut sys::zw_ctx }
impl Context {
    pub fn new(device: i32) -> Result<Self, EncodingError> {
        let h = unsafe { sys::zw_create(device, core::ptr::null()) };
        if h.is_null() { Err(EncodingError::Device(unsafe { sys::zw_last_error() })) } else { Ok(Self { h }) }
    }
    /// Batch entry point: one `.webp` per image, byte-identical to `WebPEncoder::encode` of the reference.
    pub fn encode_batch(&mut self, imgs: &[ImageRef<'_>], p: &EncoderParams) -> Vec<Result<Vec<u8>, EncodingError>> {
        let cimgs: Vec<sys::zw_image> = imgs.iter().map(|i| sys::zw_image {
            data: i.data.as_ptr(), len: i.data.len(), width: i.width, height: i.height,
            color: match i.color { ColorType::Rgb8 => sys::ZW_COLOR_RGB8, ColorType::Rgba8 => sys::ZW_COLOR_RGBA8, ColorType::L8 => 0, ColorType::La8 => 1 },
            reserved: 0 }).collect();
        let mut outs: Vec<sys::zw_output> = imgs.iter().map(|_| sys::zw_output { data: core::ptr::null_mut(), cap: 0, len: 0, status: 0, reserved: 0 }).collect();
        let rc = unsafe { sys::zw_encode_webp_batch(self.h, cimgs.as_ptr(), cimgs.len(), p.lossy_quality as i32, p.method as i32, outs.as_mut_ptr(), core::ptr::null_mut()) };
        outs.iter().map(|o| {
            let st = if rc != 0 { rc } else { o.status };
            let r = match st {
                0 => Ok(unsafe { core::slice::from_raw_parts(o.data, o.len) }.to_vec()),
                1 => Err(EncodingError::InvalidDimensions),
                2 => Err(EncodingError::InvalidBufferSize("width/height doesn't match data length".into())),
                c => Err(EncodingError::Device(c)),
            };
            unsafe { sys::zw_free(o.data as *mut _) };
            r
        }).collect()
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::zw_destroy(self.h) } } }

/// Same shape as the reference: `WebPEncoder::new(&mut out); set_params(..); encode(data, w, h, color)`.
pub struct WebPEncoder<'a> { writer: &'a mut Vec<u8>, params: EncoderParams, ctx: Context }
impl<'a> WebPEncoder<'a> {
    pub fn new(w: &'a mut Vec<u8>) -> Self { Self { writer: w, params: EncoderParams::default(), ctx: Context::new(0).expect("no CUDA device (there is no CPU fallback)") } }
    pub fn set_params(&mut self, params: EncoderParams) { self.params = params; }
    pub fn encode(mut self, data: &[u8], width: u32, height: u32, color: ColorType) -> Result<(), EncodingError> {
        if width > 65535 || height > 65535 { return Err(EncodingError::InvalidDimensions); }
        let out = self.ctx.encode_batch(&[ImageRef { data, width, height, color }], &self.params).pop().unwrap()?;
        self.writer.extend_from_slice(&out);
        Ok(())
    }
}

// SAFETY: a context is only ever used by the thread that currently owns it (the C ABI is
// thread-safe across contexts; one context must not be used from two threads at once).
unsafe impl Send for Context {}

/// Streaming batch entry: `depth` contexts on one GPU, one worker thread each, so that the H2D
/// copy, the D2H copy and the host RIFF assembly of one batch run under the kernels of the next
/// (same design as `BatchPipeline` in the Python / C++ mirrors).  `encode_batches` returns the
/// per-batch results in order.
pub struct BatchPipeline { ctxs: Vec<Context> }
impl BatchPipeline {
    pub fn new(device: i32, depth: usize) -> Result<Self, EncodingError> {
        Ok(Self { ctxs: (0..depth.max(1)).map(|_| Context::new(device)).collect::<Result<_, _>>()? })
    }
    pub fn encode_batches(&mut self, batches: &[Vec<ImageRef<'_>>], p: &EncoderParams) -> Vec<Vec<Result<Vec<u8>, EncodingError>>> {
        let depth = self.ctxs.len();
        let mut out: Vec<Option<Vec<Result<Vec<u8>, EncodingError>>>> = (0..batches.len()).map(|_| None).collect();
        std::thread::scope(|s| {
            let handles: Vec<_> = self.ctxs.iter_mut().enumerate().map(|(k, ctx)| {
                s.spawn(move || (k..batches.len()).step_by(depth).map(|i| (i, ctx.encode_batch(&batches[i], p))).collect::<Vec<_>>())
            }).collect();
            for h in handles { for (i, r) in h.join().unwrap() { out[i] = Some(r); } }
        });
        out.into_iter().map(|o| o.unwrap()).collect()
    }
}
