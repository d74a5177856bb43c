//! Batched drop-in for the hot path of ninjasource/lc3-codec:
//! `Lc3Decoder::decode_frame` (src/decoder/lc3_decoder.rs:217) and `Lc3Encoder::encode_frame`
//! (src/encoder/lc3_encoder.rs:175), one call per frame period for ALL streams.
//!
//! Preallocated-buffer style as in the reference: ask for the working-buffer size, allocate it yourself (device
//! memory, 256-byte aligned), lend it to `new` for `'a`.  Nothing is allocated on the device afterwards.
#![no_std]
use core::ffi::c_void;
use core::marker::PhantomData;
use core::ptr;
use lc3b_sys as sys;

/// `SamplingFrequency` (src/common/config.rs:2); discriminants are the C ABI's.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum SamplingFrequency {
    Hz8000 = 0,
    Hz16000 = 1,
    Hz24000 = 2,
    Hz32000 = 3,
    Hz44100 = 4,
    Hz48000 = 5,
}
/// `FrameDuration` (src/common/config.rs:12)
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum FrameDuration {
    SevenPointFiveMs = 0,
    TenMs = 1,
}

/// `Lc3DecoderError` (src/decoder/lc3_decoder.rs:36-41); bitstream errors are concealed like the reference (:138-141).
#[derive(Debug, PartialEq, Eq)]
pub enum Lc3DecoderError {
    Only16BitsPerAudioSampleSupported,
}
/// `Lc3EncoderError` is empty in the reference (src/encoder/lc3_encoder.rs:30): encoding cannot fail.
#[derive(Debug)]
pub enum Lc3EncoderError {}

/// A caller-owned device allocation (any CUDA binding can produce the pointer).
pub struct DeviceBuf<'a> {
    pub ptr: *mut c_void,
    pub bytes: usize,
    pub _owner: PhantomData<&'a mut [u8]>,
}

/// Where a call's I/O buffers live.
#[derive(Clone, Copy)]
pub enum Residency {
    /// device pointers: asynchronous on the stream, no copies
    Device,
    /// host pointers (pinned => asynchronous): the library stages H2D / D2H itself
    Host,
}

fn check(rc: i32, what: &str) {
    // LC3B_ERR_INVALID_ARG stands where the reference panics; CUDA / workspace failures have no reference analogue
    if rc != sys::LC3B_OK {
        panic!("lc3b: {} failed with status {} (cuda error {})", what, rc, unsafe { sys::lc3b_last_cuda_error() });
    }
}

pub struct Lc3BatchDecoder<'a> {
    h: *mut sys::lc3b_decoder,
    n_streams: usize,
    nf: usize,
    stream: *mut c_void,
    _ws: PhantomData<&'a mut [u8]>,
}

impl<'a> Lc3BatchDecoder<'a> {
    /// `Lc3Decoder::calc_working_buffer_lengths` (:236): bytes of device working memory for `num_streams` channels.
    pub fn calc_working_buffer_lengths(num_streams: usize, duration: FrameDuration, freq: SamplingFrequency,
                                       max_nbytes: usize) -> usize {
        let mut n = 0usize;
        check(unsafe { sys::lc3b_decoder_workspace_bytes(num_streams as i32, duration as i32, freq as i32, max_nbytes as i32, &mut n) },
              "lc3b_decoder_workspace_bytes");
        n
    }

    /// `Lc3Decoder::new` (:181): borrows the working buffer for `'a`; state starts zeroed.
    pub fn new(num_streams: usize, duration: FrameDuration, freq: SamplingFrequency, working: DeviceBuf<'a>,
               max_nbytes: usize, device: i32, cuda_stream: *mut c_void) -> Self {
        let mut cfg = sys::lc3b_config::default();
        check(unsafe { sys::lc3b_config_new(freq as i32, duration as i32, &mut cfg) }, "lc3b_config_new");
        let mut h = ptr::null_mut();
        check(unsafe { sys::lc3b_decoder_init(&mut h, num_streams as i32, duration as i32, freq as i32, max_nbytes as i32,
                                              device, working.ptr, working.bytes, cuda_stream) }, "lc3b_decoder_init");
        Self { h, n_streams: num_streams, nf: cfg.nf as usize, stream: cuda_stream, _ws: PhantomData }
    }

    /// `decode_frame` for every stream: `frames[s*stride .. s*stride + nbytes]` -> `samples_out[s*nf .. (s+1)*nf]`.
    /// `frame_nbytes` (optional, same residency) gives per-stream lengths; 0 = lost frame (concealed).
    ///
    /// # Safety
    /// The pointers must be valid for the stated extents in the stated residency until the stream has drained.
    pub unsafe fn decode_frames(&mut self, num_bits_per_audio_sample: usize, residency: Residency, frames: *const u8,
                                frame_nbytes: *const i32, nbytes: usize, stride: usize, samples_out: *mut i16)
                                -> Result<(), Lc3DecoderError> {
        let f = match residency {
            Residency::Device => sys::lc3b_decode_frames,
            Residency::Host => sys::lc3b_decode_frames_host,
        };
        match f(self.h, num_bits_per_audio_sample as i32, frames, frame_nbytes, nbytes as i32, stride, samples_out, self.nf,
                ptr::null_mut(), self.stream) {
            sys::LC3B_OK => Ok(()),
            sys::LC3B_ERR_BITS_PER_SAMPLE => Err(Lc3DecoderError::Only16BitsPerAudioSampleSupported),
            rc => { check(rc, "lc3b_decode_frames"); unreachable!() }
        }
    }

    /// Let the PCM copy of call i overlap the kernels of call i+1 (host residency only); see `host_fence`.
    pub fn set_host_pipelining(&mut self, on: bool) {
        check(unsafe { sys::lc3b_decoder_set_host_pipelining(self.h, on as i32) }, "lc3b_decoder_set_host_pipelining");
    }
    /// Makes the stream wait for every outstanding pipelined PCM copy.
    pub fn host_fence(&mut self) {
        check(unsafe { sys::lc3b_decoder_host_fence(self.h, self.stream) }, "lc3b_decoder_host_fence");
    }
    pub fn num_streams(&self) -> usize { self.n_streams }
    pub fn samples_per_frame(&self) -> usize { self.nf }
}
impl Drop for Lc3BatchDecoder<'_> {
    fn drop(&mut self) { unsafe { sys::lc3b_decoder_destroy(self.h) } }
}

pub struct Lc3BatchEncoder<'a> {
    h: *mut sys::lc3b_encoder,
    n_streams: usize,
    nf: usize,
    stream: *mut c_void,
    _ws: PhantomData<&'a mut [u8]>,
}

impl<'a> Lc3BatchEncoder<'a> {
    /// `Lc3Encoder::calc_working_buffer_lengths` (src/encoder/lc3_encoder.rs:194).  Panics at 8 kHz like `Lc3Encoder::new`
    /// (bandwidth_detector.rs:42-56).
    pub fn calc_working_buffer_lengths(num_streams: usize, duration: FrameDuration, freq: SamplingFrequency,
                                       max_nbytes: usize) -> usize {
        let mut n = 0usize;
        check(unsafe { sys::lc3b_encoder_workspace_bytes(num_streams as i32, duration as i32, freq as i32, max_nbytes as i32, &mut n) },
              "lc3b_encoder_workspace_bytes");
        n
    }
    /// `Lc3Encoder::new` (:117)
    pub fn new(num_streams: usize, duration: FrameDuration, freq: SamplingFrequency, working: DeviceBuf<'a>,
               max_nbytes: usize, device: i32, cuda_stream: *mut c_void) -> Self {
        let mut cfg = sys::lc3b_config::default();
        check(unsafe { sys::lc3b_config_new(freq as i32, duration as i32, &mut cfg) }, "lc3b_config_new");
        let mut h = ptr::null_mut();
        check(unsafe { sys::lc3b_encoder_init(&mut h, num_streams as i32, duration as i32, freq as i32, max_nbytes as i32,
                                              device, working.ptr, working.bytes, cuda_stream) }, "lc3b_encoder_init");
        Self { h, n_streams: num_streams, nf: cfg.nf as usize, stream: cuda_stream, _ws: PhantomData }
    }
    /// `encode_frame` (:175) for every stream: `samples_in[s*nf ..]` -> `buf_out[s*stride .. s*stride + nbytes]`.
    ///
    /// # Safety
    /// As for `Lc3BatchDecoder::decode_frames`.
    pub unsafe fn encode_frames(&mut self, residency: Residency, samples_in: *const i16, buf_out: *mut u8, nbytes: usize,
                                stride: usize) -> Result<(), Lc3EncoderError> {
        let f = match residency {
            Residency::Device => sys::lc3b_encode_frames,
            Residency::Host => sys::lc3b_encode_frames_host,
        };
        check(f(self.h, samples_in, self.nf, buf_out, nbytes as i32, stride, self.stream), "lc3b_encode_frames");
        Ok(())
    }
    pub fn num_streams(&self) -> usize { self.n_streams }
    pub fn samples_per_frame(&self) -> usize { self.nf }
}
impl Drop for Lc3BatchEncoder<'_> {
    fn drop(&mut self) { unsafe { sys::lc3b_encoder_destroy(self.h) } }
}
