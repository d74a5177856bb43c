//! Raw bindings to `include/lc3b.h`.  Each function names the reference method it replaces.
#![no_std]
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const LC3B_OK: c_int = 0;
/// `Lc3DecoderError::Only16BitsPerAudioSampleSupported` (src/decoder/lc3_decoder.rs:80)
pub const LC3B_ERR_BITS_PER_SAMPLE: c_int = 1;
/// where the reference panics (bad channel index :228, slice-length asserts)
pub const LC3B_ERR_INVALID_ARG: c_int = 2;
pub const LC3B_ERR_CUDA: c_int = 3;
pub const LC3B_ERR_WORKSPACE: c_int = 4;

#[repr(C)]
pub struct lc3b_decoder {
    _private: [u8; 0],
}
#[repr(C)]
pub struct lc3b_encoder {
    _private: [u8; 0],
}
/// `Lc3Config` (src/common/config.rs:18-39)
#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct lc3b_config {
    pub fs_ind: i32,
    pub fs: i32,
    pub ne: i32,
    pub nb: i32,
    pub nf: i32,
    pub z: i32,
    pub n_ms: i32,
}

extern "C" {
    /// `Lc3Config::new` (src/common/config.rs:42)
    pub fn lc3b_config_new(sampling_frequency: c_int, frame_duration: c_int, out: *mut lc3b_config) -> c_int;
    pub fn lc3b_last_cuda_error() -> c_int;
    pub fn lc3b_version() -> *const c_char;

    /// `Lc3Decoder::calc_working_buffer_lengths` (src/decoder/lc3_decoder.rs:236)
    pub fn lc3b_decoder_workspace_bytes(n_streams: c_int, frame_duration: c_int, sampling_frequency: c_int,
                                        max_nbytes: c_int, device_bytes: *mut usize) -> c_int;
    /// `Lc3Decoder::new` (src/decoder/lc3_decoder.rs:181)
    pub fn lc3b_decoder_init(out: *mut *mut lc3b_decoder, n_streams: c_int, frame_duration: c_int,
                             sampling_frequency: c_int, max_nbytes: c_int, device: c_int, dev_workspace: *mut c_void,
                             workspace_bytes: usize, cuda_stream: *mut c_void) -> c_int;
    /// one `Lc3Decoder::decode_frame` (src/decoder/lc3_decoder.rs:217) per stream, device buffers
    pub fn lc3b_decode_frames(h: *mut lc3b_decoder, bits_per_sample: c_int, frames: *const u8, frame_nbytes: *const i32,
                              nbytes: c_int, frame_stride: usize, pcm_out: *mut i16, pcm_stride: usize,
                              status_out: *mut i32, cuda_stream: *mut c_void) -> c_int;
    /// same, host buffers (pinned => asynchronous)
    pub fn lc3b_decode_frames_host(h: *mut lc3b_decoder, bits_per_sample: c_int, frames: *const u8,
                                   frame_nbytes: *const i32, nbytes: c_int, frame_stride: usize, pcm_out: *mut i16,
                                   pcm_stride: usize, status_out: *mut i32, cuda_stream: *mut c_void) -> c_int;
    pub fn lc3b_decoder_set_host_pipelining(h: *mut lc3b_decoder, on: c_int) -> c_int;
    pub fn lc3b_decoder_host_fence(h: *mut lc3b_decoder, cuda_stream: *mut c_void) -> c_int;
    /// the frame loop of examples/decode.rs:85-123 as one call: `n_frames` frames of every stream (time-parallel)
    pub fn lc3b_decoder_multi_scratch_bytes(h: *const lc3b_decoder, n_frames: c_int, device_bytes: *mut usize) -> c_int;
    pub fn lc3b_decode_stream_frames(h: *mut lc3b_decoder, bits_per_sample: c_int, frames: *const u8, frame_nbytes: *const i32,
                                     nbytes: c_int, frame_stride: usize, n_frames: c_int, pcm_out: *mut i16,
                                     status_out: *mut i32, scratch: *mut c_void, scratch_bytes: usize,
                                     cuda_stream: *mut c_void) -> c_int;
    pub fn lc3b_decoder_destroy(h: *mut lc3b_decoder);

    /// `Lc3Encoder::calc_working_buffer_lengths` (src/encoder/lc3_encoder.rs:194)
    pub fn lc3b_encoder_workspace_bytes(n_streams: c_int, frame_duration: c_int, sampling_frequency: c_int,
                                        max_nbytes: c_int, device_bytes: *mut usize) -> c_int;
    /// `Lc3Encoder::new` (src/encoder/lc3_encoder.rs:117)
    pub fn lc3b_encoder_init(out: *mut *mut lc3b_encoder, n_streams: c_int, frame_duration: c_int,
                             sampling_frequency: c_int, max_nbytes: c_int, device: c_int, dev_workspace: *mut c_void,
                             workspace_bytes: usize, cuda_stream: *mut c_void) -> c_int;
    /// one `Lc3Encoder::encode_frame` (src/encoder/lc3_encoder.rs:175) per stream, device buffers
    pub fn lc3b_encode_frames(h: *mut lc3b_encoder, pcm_in: *const i16, pcm_stride: usize, frames_out: *mut u8,
                              nbytes: c_int, frame_stride: usize, cuda_stream: *mut c_void) -> c_int;
    /// same, host buffers
    pub fn lc3b_encode_frames_host(h: *mut lc3b_encoder, pcm_in: *const i16, pcm_stride: usize, frames_out: *mut u8,
                                   nbytes: c_int, frame_stride: usize, cuda_stream: *mut c_void) -> c_int;
    pub fn lc3b_encoder_set_host_pipelining(h: *mut lc3b_encoder, on: c_int) -> c_int;
    pub fn lc3b_encoder_destroy(h: *mut lc3b_encoder);
}
