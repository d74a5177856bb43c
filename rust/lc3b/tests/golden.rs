//! The reference's own golden tests, re-stated against this crate (run with `cargo test` on a B200 box once a Rust
//! toolchain is available: `LC3B_LIB_DIR=../../lc3_codec_b200 cargo test`).  NOT COMPILED in the build image (no
//! cargo / rustc there); the same vectors are asserted through the same C ABI by tests/test_decoder_gpu.py,
//! tests/test_encoder_gpu.py and tests/test_abi_cpu.py.
//!
//! Of the reference's 39 `#[test]`s, the ones whose subject is reachable through the drop-in boundary are here:
//!   lc3_decode_channel      src/decoder/lc3_decoder.rs:374      end to end, PCM within +-1 LSB (own FFT factorisation)
//!   lc3_encode_channel      src/encoder/lc3_encoder.rs:314      end to end, bytes exact
//!   arithmetic_decode       src/decoder/arithmetic_codec.rs:415 seed / TNS orders / coefficients via the inspection record
//!   decode_noise_filling    src/decoder/noise_filling.rs:65     its integer spectrum input = the entropy decoder's output
//!   simple_config           src/common/config.rs:109            derived sizes
//!   only-16-bit error       src/decoder/lc3_decoder.rs:80
//! The remaining ones assert on functions that exist only INSIDE the kernels here (buffer reader / writer, global gain,
//! TNS, SNS, PLC, MDCT, LTPF, the encoder stages, kissfft, DCT-IV): their vectors pin the CPU oracle the kernels are
//! compared with (tests/test_oracle_golden.py replays all 39), which is where they can be asserted exactly.
use core::ffi::c_void;
use core::marker::PhantomData;
use core::ptr;

use lc3b::{DeviceBuf, FrameDuration, Lc3BatchDecoder, Lc3BatchDecoderN, Lc3BatchEncoder, Lc3DecoderError, Residency, SamplingFrequency};
use lc3b_sys as sys;

mod golden_vectors;
use golden_vectors::*;

#[link(name = "cudart")]
extern "C" {
    fn cudaMalloc(p: *mut *mut c_void, bytes: usize) -> i32;
    fn cudaFree(p: *mut c_void) -> i32;
    fn cudaMemcpy(dst: *mut c_void, src: *const c_void, bytes: usize, kind: i32) -> i32;
    fn cudaDeviceSynchronize() -> i32;
}
const D2H: i32 = 2;

struct Dev {
    p: *mut c_void,
    bytes: usize,
}
impl Dev {
    fn new(bytes: usize) -> Self {
        let mut p = ptr::null_mut();
        assert_eq!(unsafe { cudaMalloc(&mut p, bytes) }, 0);
        Dev { p, bytes }
    }
    fn buf(&self) -> DeviceBuf<'_> { DeviceBuf { ptr: self.p, bytes: self.bytes, _owner: PhantomData } }
}
impl Drop for Dev {
    fn drop(&mut self) { unsafe { cudaFree(self.p); } }
}

const STREAMS: usize = 33; // the vector replicated over more than one warp of streams

#[test]
fn lc3_decode_channel() {
    let n = Lc3BatchDecoder::calc_working_buffer_lengths(STREAMS, FrameDuration::TenMs, SamplingFrequency::Hz48000, 150);
    let ws = Dev::new(n);
    let mut dec = Lc3BatchDecoder::new(STREAMS, FrameDuration::TenMs, SamplingFrequency::Hz48000, ws.buf(), 150, 0, ptr::null_mut());
    let mut frames = vec![0u8; STREAMS * 150];
    for s in 0..STREAMS { frames[s * 150..(s + 1) * 150].copy_from_slice(&DECODE_FRAME); }
    let mut pcm = vec![0i16; STREAMS * 480];
    let mut status = vec![-1i32; STREAMS];
    unsafe {
        dec.decode_frames_with_status(16, Residency::Host, frames.as_ptr(), ptr::null(), 150, 150, pcm.as_mut_ptr(), status.as_mut_ptr()).unwrap();
        assert_eq!(cudaDeviceSynchronize(), 0);
    }
    assert!(status.iter().all(|&s| s == 0));
    for s in 0..STREAMS {
        for (got, exp) in pcm[s * 480..(s + 1) * 480].iter().zip(DECODE_PCM.iter()) {
            assert!((*got as i32 - *exp as i32).abs() <= 1);
        }
        assert_eq!(&pcm[s * 480..(s + 1) * 480], &pcm[0..480]); // every stream decodes identically
    }
}

#[test]
fn arithmetic_decode_and_noise_filling_input() {
    let n = Lc3BatchDecoder::calc_working_buffer_lengths(1, FrameDuration::TenMs, SamplingFrequency::Hz48000, 150);
    let ws = Dev::new(n);
    let mut dec = Lc3BatchDecoder::new(1, FrameDuration::TenMs, SamplingFrequency::Hz48000, ws.buf(), 150, 0, ptr::null_mut());
    let trace = Dev::new(4 * sys::LC3B_TRACE_WORDS);
    let x = Dev::new(4 * 400);
    assert_eq!(unsafe { sys::lc3b_decoder_set_trace(dec.raw(), trace.p as *mut i32, x.p as *mut i32) }, 0);
    let mut pcm = [0i16; 480];
    let (mut tr, mut xi) = ([0i32; sys::LC3B_TRACE_WORDS], [0i32; 400]);
    unsafe {
        dec.decode_frames(16, Residency::Host, DECODE_FRAME.as_ptr(), ptr::null(), 150, 150, pcm.as_mut_ptr()).unwrap();
        assert_eq!(cudaDeviceSynchronize(), 0);
        assert_eq!(cudaMemcpy(tr.as_mut_ptr() as *mut c_void, trace.p, 4 * sys::LC3B_TRACE_WORDS, D2H), 0);
        assert_eq!(cudaMemcpy(xi.as_mut_ptr() as *mut c_void, x.p, 4 * 400, D2H), 0);
    }
    // arithmetic_codec.rs:458-473 (indices: LC3B_TR_* of include/lc3b.h)
    assert_eq!(tr[0], 1);                       // LC3B_TR_OK
    assert_eq!(tr[40], 56909);                  // LC3B_TR_SEED: noise_filling_seed
    assert_eq!(&tr[21..23], &[8, 0]);           // LC3B_TR_RC_ORDER0/1: reflect_coef_order
    assert_eq!(&tr[23..39], &[6, 10, 7, 8, 7, 9, 7, 7, 0, 0, 0, 0, 0, 0, 0, 0]);   // reflect_coef_ints
    assert_eq!(tr[39], 45);                     // LC3B_TR_NRES: residual_bits.len()
    assert_eq!(tr[41], 0);                      // LC3B_TR_IS_ZERO
    assert_eq!(xi, ARITH_X);                    // noise_filling.rs:65 takes exactly this integer spectrum
}

#[test]
fn lc3_encode_channel() {
    let n = Lc3BatchEncoder::calc_working_buffer_lengths(STREAMS, FrameDuration::TenMs, SamplingFrequency::Hz48000, 150);
    let ws = Dev::new(n);
    let mut enc = Lc3BatchEncoder::new(STREAMS, FrameDuration::TenMs, SamplingFrequency::Hz48000, ws.buf(), 150, 0, ptr::null_mut());
    let mut pcm = vec![0i16; STREAMS * 480];
    for s in 0..STREAMS { pcm[s * 480..(s + 1) * 480].copy_from_slice(&ENCODE_PCM); }
    let mut out = vec![0u8; STREAMS * 150];
    unsafe {
        enc.encode_frames(Residency::Host, pcm.as_ptr(), out.as_mut_ptr(), 150, 150).unwrap();
        assert_eq!(cudaDeviceSynchronize(), 0);
    }
    for s in 0..STREAMS { assert_eq!(&out[s * 150..(s + 1) * 150], &ENCODE_FRAME[..]); }
}

#[test]
fn const_generic_twin_decodes_the_same_vector() {
    // the reference's no-alloc API (src/decoder/lc3_decoder.rs:247-310): stream count in the type
    let n = Lc3BatchDecoderN::<2>::calc_working_buffer_lengths(FrameDuration::TenMs, SamplingFrequency::Hz48000, 150);
    let ws = Dev::new(n);
    let mut dec = Lc3BatchDecoderN::<2>::new(FrameDuration::TenMs, SamplingFrequency::Hz48000, ws.buf(), 150, 0, ptr::null_mut());
    let frames = [DECODE_FRAME, DECODE_FRAME];
    let mut pcm = [[0i16; 480]; 2];
    dec.decode_frames_host(16, &frames, &mut pcm).unwrap();
    unsafe { assert_eq!(cudaDeviceSynchronize(), 0); }
    for ch in pcm.iter() {
        for (got, exp) in ch.iter().zip(DECODE_PCM.iter()) { assert!((*got as i32 - *exp as i32).abs() <= 1); }
    }
}

#[test]
fn only_16_bits_per_audio_sample() {
    // src/decoder/lc3_decoder.rs:80 - the one error the reference returns
    let n = Lc3BatchDecoder::calc_working_buffer_lengths(1, FrameDuration::TenMs, SamplingFrequency::Hz48000, 150);
    let ws = Dev::new(n);
    let mut dec = Lc3BatchDecoder::new(1, FrameDuration::TenMs, SamplingFrequency::Hz48000, ws.buf(), 150, 0, ptr::null_mut());
    let mut pcm = [0i16; 480];
    let r = unsafe { dec.decode_frames(24, Residency::Host, DECODE_FRAME.as_ptr(), ptr::null(), 150, 150, pcm.as_mut_ptr()) };
    assert_eq!(r, Err(Lc3DecoderError::Only16BitsPerAudioSampleSupported));
}

#[test]
fn simple_config() {
    // src/common/config.rs:109
    let mut c = sys::lc3b_config::default();
    assert_eq!(unsafe { sys::lc3b_config_new(SamplingFrequency::Hz48000 as i32, FrameDuration::TenMs as i32, &mut c) }, 0);
    assert_eq!((c.fs_ind, c.fs, c.ne, c.nb, c.nf, c.z), (4, 48000, 400, 64, 480, 180));
    assert_eq!(unsafe { sys::lc3b_config_new(SamplingFrequency::Hz44100 as i32, FrameDuration::TenMs as i32, &mut c) }, 0);
    assert_eq!((c.fs_ind, c.fs, c.nf), (4, 44100, 480));   // 44.1 kHz shares every size with 48 kHz (config.rs:48-49)
}
