//! Safe wrapper with the reference's names: `WebPEncoder`, `EncoderParams`, `ColorType`,
//! `EncodingError` (zenwebp 0.2.0 src/encoder/api.rs) + `encode_batch`.
//! NOTE: written but NOT compiled in the build image (no cargo/rustc there); see INTEGRATION.md.
#![forbid(unsafe_op_in_unsafe_fn)]
use zenwebp_b200_sys as sys;

#[derive(Copy, Clone, Debug, PartialEq, Eq)]
#[repr(u32)]
pub enum ColorType { L8 = 0, La8 = 1, Rgb8 = 2, Rgba8 = 3 }

#[derive(Debug, thiserror::Error)]
#[non_exhaustive]
pub enum EncodingError {
    #[error("Invalid dimensions")]
    InvalidDimensions,
    #[error("Invalid buffer size: {0}")]
    InvalidBufferSize(String),
    /// Lossless (VP8L) parameters, or an alpha colour type in the simple container (needs VP8X + ALPH): not built.
    #[error("unsupported by the GPU path: {0}")]
    Unsupported(&'static str),
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

#[derive(Default, Clone)]
pub struct Metadata { pub icc_profile: Vec<u8>, pub exif: Vec<u8>, pub xmp: Vec<u8> }

pub struct ImageRef<'a> { pub data: &'a [u8], pub width: u32, pub height: u32, pub color: ColorType }

pub struct Context { h: *mut sys::zw_ctx }
impl Context {
    pub fn new(device: i32) -> Result<Self, EncodingError> {
        let h = unsafe { sys::zw_create(device, core::ptr::null()) };
        if h.is_null() { Err(EncodingError::Device(unsafe { sys::zw_last_error() })) } else { Ok(Self { h }) }
    }
    /// Batch entry point: one `.webp` per image, byte-identical to `WebPEncoder::encode` of the reference.
    pub fn encode_batch(&mut self, imgs: &[ImageRef<'_>], p: &EncoderParams) -> Vec<Result<Vec<u8>, EncodingError>> {
        self.encode_batch_with_metadata(imgs, p, &[])
    }
    /// Same with ICC / EXIF / XMP per image (`meta`: empty, or one entry per image).  Lossy ("VP8 ") or lossless ("VP8L")
    /// frame, simple or extended (VP8X + ICCP + ALPH + frame + EXIF + XMP) container -- everything `WebPEncoder::encode` writes.
    pub fn encode_batch_with_metadata(&mut self, imgs: &[ImageRef<'_>], p: &EncoderParams, meta: &[Metadata]) -> Vec<Result<Vec<u8>, EncodingError>> {
        assert!(meta.is_empty() || meta.len() == imgs.len());
        let zp = sys::zw_params { use_predictor_transform: p.use_predictor_transform as i32, use_lossy: p.use_lossy as i32,
                                  lossy_quality: p.lossy_quality as i32, method: p.method as i32 };
        let cmeta: Vec<sys::zw_metadata> = meta.iter().map(|m| sys::zw_metadata {
            icc_profile: m.icc_profile.as_ptr(), icc_len: m.icc_profile.len(), exif: m.exif.as_ptr(), exif_len: m.exif.len(),
            xmp: m.xmp.as_ptr(), xmp_len: m.xmp.len() }).collect();
        let cimgs: Vec<sys::zw_image> = imgs.iter().map(|i| sys::zw_image {
            data: i.data.as_ptr(), len: i.data.len(), width: i.width, height: i.height,
            color: match i.color { ColorType::Rgb8 => sys::ZW_COLOR_RGB8, ColorType::Rgba8 => sys::ZW_COLOR_RGBA8, ColorType::L8 => 0, ColorType::La8 => 1 },
            reserved: 0 }).collect();
        let mut outs: Vec<sys::zw_output> = imgs.iter().map(|_| sys::zw_output { data: core::ptr::null_mut(), cap: 0, len: 0, status: 0, reserved: 0 }).collect();
        let rc = unsafe { sys::zw_encode_batch(self.h, cimgs.as_ptr(), cimgs.len(), &zp, if cmeta.is_empty() { core::ptr::null() } else { cmeta.as_ptr() },
                                               outs.as_mut_ptr(), core::ptr::null_mut()) };
        outs.iter().map(|o| {
            let st = if rc != 0 { rc } else { o.status };
            let r = match st {
                0 => Ok(unsafe { core::slice::from_raw_parts(o.data, o.len) }.to_vec()),
                1 => Err(EncodingError::InvalidDimensions),
                2 => Err(EncodingError::InvalidBufferSize("width/height doesn't match data length".into())),
                3 => Err(EncodingError::Unsupported("bad parameter (quality > 100: the reference panics)")),
                c => Err(EncodingError::Device(c)),
            };
            if !o.data.is_null() { unsafe { sys::zw_free(o.data as *mut _) }; }
            r
        }).collect()
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::zw_destroy(self.h) } } }

thread_local! {
    /// One cached context per (thread, device 0): building a context allocates streams, events and tables, which a
    /// drop-in single-image caller must not pay per image.
    static DEFAULT_CTX: std::cell::RefCell<Option<Context>> = const { std::cell::RefCell::new(None) };
}

/// Same shape as the reference: `WebPEncoder::new(&mut out); set_params(..); encode(data, w, h, color)`.
pub struct WebPEncoder<'a> { writer: &'a mut Vec<u8>, params: EncoderParams, meta: Metadata }
impl<'a> WebPEncoder<'a> {
    pub fn new(w: &'a mut Vec<u8>) -> Self { Self { writer: w, params: EncoderParams::default(), meta: Metadata::default() } }
    pub fn set_params(&mut self, params: EncoderParams) { self.params = params; }
    pub fn set_icc_profile(&mut self, v: Vec<u8>) { self.meta.icc_profile = v; }
    pub fn set_exif_metadata(&mut self, v: Vec<u8>) { self.meta.exif = v; }
    pub fn set_xmp_metadata(&mut self, v: Vec<u8>) { self.meta.xmp = v; }
    pub fn encode(self, data: &[u8], width: u32, height: u32, color: ColorType) -> Result<(), EncodingError> {
        let out = DEFAULT_CTX.with(|c| -> Result<Vec<u8>, EncodingError> {
            let mut c = c.borrow_mut();
            if c.is_none() { *c = Some(Context::new(0)?); }  // no CUDA device -> Device(code): there is no CPU fallback
            let any = !(self.meta.icc_profile.is_empty() && self.meta.exif.is_empty() && self.meta.xmp.is_empty());
            let meta = if any { vec![self.meta.clone()] } else { Vec::new() };
            c.as_mut().unwrap().encode_batch_with_metadata(&[ImageRef { data, width, height, color }], &self.params, &meta).pop().unwrap()
        })?;
        self.writer.extend_from_slice(&out);
        Ok(())
    }
}

// SAFETY: a context is only ever used by the thread that currently owns it (the C ABI is
// thread-safe across contexts; one context must not be used from two threads at once).
unsafe impl Send for Context {}

/// Streaming batch entry on ONE context (`zw_submit` / `zw_wait` / `zw_release`): up to `depth` batches in flight,
/// the copies of one hidden behind the kernels of its neighbours.  `encode_batches` returns the per-batch results in order.
pub struct BatchPipeline { ctx: Context, depth: usize }
impl BatchPipeline {
    pub fn new(device: i32, depth: usize) -> Result<Self, EncodingError> {
        let lim = sys::zw_limits { max_device_bytes: 0, persistent_warps_per_sm: 0, reserved: [depth.clamp(1, 8) as i32, 0, 0, 0, 0] };
        let h = unsafe { sys::zw_create(device, &lim) };
        if h.is_null() { return Err(EncodingError::Device(unsafe { sys::zw_last_error() })); }
        Ok(Self { ctx: Context { h }, depth: depth.clamp(1, 8) })
    }
    fn collect(&mut self, ticket: i32) -> Vec<Result<Vec<u8>, EncodingError>> {
        let mut v = sys::zw_batch_view { arena: core::ptr::null(), n: 0, offsets: core::ptr::null(), lens: core::ptr::null(), status: core::ptr::null() };
        let rc = unsafe { sys::zw_wait(self.ctx.h, ticket, 1, &mut v, core::ptr::null_mut()) };
        let res = if rc != 0 { Vec::new() } else {
            (0..v.n).map(|i| unsafe {
                match *v.status.add(i) {
                    0 => Ok(core::slice::from_raw_parts(v.arena.add(*v.offsets.add(i) as usize), *v.lens.add(i) as usize).to_vec()),
                    1 => Err(EncodingError::InvalidDimensions),
                    2 => Err(EncodingError::InvalidBufferSize("width/height doesn't match data length".into())),
                    c => Err(EncodingError::Device(c)),
                }
            }).collect()
        };
        unsafe { sys::zw_release(self.ctx.h, ticket) };
        res
    }
    pub fn encode_batches(&mut self, batches: &[Vec<ImageRef<'_>>], p: &EncoderParams) -> Vec<Vec<Result<Vec<u8>, EncodingError>>> {
        let mut out = Vec::with_capacity(batches.len());
        let mut inflight: std::collections::VecDeque<(i32, Vec<sys::zw_image>)> = Default::default();
        for b in batches {
            if inflight.len() == self.depth { let (t, _keep) = inflight.pop_front().unwrap(); out.push(self.collect(t)); }
            let cimgs: Vec<sys::zw_image> = b.iter().map(|i| sys::zw_image { data: i.data.as_ptr(), len: i.data.len(), width: i.width, height: i.height,
                color: i.color as u32, reserved: 0 }).collect();
            let mut ticket = -1;
            let rc = unsafe { sys::zw_submit(self.ctx.h, cimgs.as_ptr(), cimgs.len(), p.lossy_quality as i32, p.method as i32, &mut ticket) };
            if rc != 0 { out.push(b.iter().map(|_| Err(EncodingError::Device(rc))).collect()); continue; }
            inflight.push_back((ticket, cimgs));
        }
        while let Some((t, _keep)) = inflight.pop_front() { out.push(self.collect(t)); }
        out
    }
}
