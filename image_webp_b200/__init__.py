"""image_webp_b200 -- B200-native WebP encoder core (lossy VP8 key frames, lossless VP8L, every container
WebPEncoder::encode writes) and VP8 decoder / verifier behind the C ABI of include/zenwebp_b200.h; a drop-in for the
WebPEncoder / EncoderParams / Encoder / EncoderConfig path of imazen/image-webp (`zenwebp` 0.2.0).
Hand-written CUDA for sm_100a, no CPU fallback."""
from .encoder import (BatchPipeline, ColorType, Context, DeviceError, Encoder, EncoderConfig, EncoderParams, EncodingError,
                      InvalidBufferSize, InvalidDimensions, MultiContext, PendingBatch, Preset, WebPEncoder, default_context,
                      encode_batch)

from .decoder import DecodingError, UpsamplingMethod, WebPDecoder, decode_batch, decode_rgb, verify_pending

__all__ = ["DecodingError", "UpsamplingMethod", "WebPDecoder", "decode_batch", "decode_rgb", "verify_pending", "BatchPipeline", "ColorType", "Context", "DeviceError", "Encoder", "EncoderConfig", "EncoderParams", "EncodingError", "InvalidBufferSize", "Preset",
           "InvalidDimensions", "MultiContext", "PendingBatch", "WebPEncoder", "default_context", "encode_batch"]
