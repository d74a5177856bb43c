"""Mixed-rate batches (BASELINE config 4): streams of different (sampling frequency, frame duration) in one call.

The reference builds one `Lc3Decoder` per configuration (src/decoder/lc3_decoder.rs:181 takes ONE frame duration and
ONE sampling frequency for all its channels); a mixed population is therefore a set of decoders.  `Lc3MixedBatchDecoder`
is the ctypes mirror of that set behind the C ABI (`lc3b_mixed_*`, include/lc3b.h): streams are bucketed by
configuration (stable order) and one call is one entropy launch over every bucket, one dequantisation launch per frame
duration and one synthesis + post-filter launch per configuration, issued as one CUDA graph.
Callers lay their per-stream rows out in bucket order (`order` maps row -> original stream id), so every bucket reads
and writes a contiguous row range of the caller's buffers: no gather, no copy.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import native
from .decoder import Lc3DecoderError, _ptr
from .native import FrameDuration, Lc3bError, SamplingFrequency


class Lc3MixedBatchDecoder:
    def __init__(self, stream_configs: list[tuple[SamplingFrequency, FrameDuration]], max_nbytes: int = 400,
                 device: str | torch.device = "cuda"):
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        n = len(stream_configs)
        self._sf = np.ascontiguousarray([int(sf) for sf, _ in stream_configs], np.int32)
        self._fd = np.ascontiguousarray([int(fd) for _, fd in stream_configs], np.int32)
        lib = native.lib()
        order = np.zeros(max(n, 1), np.int32)
        buckets = (native.MixedBucket * 12)()
        nb, elems = C.c_int32(0), C.c_uint64(0)
        rc = lib.lc3b_mixed_decoder_layout(n, self._sf.ctypes.data, self._fd.ctypes.data, order.ctypes.data, buckets,
                                           C.byref(nb), C.byref(elems))
        if rc:
            raise Lc3bError(rc, "lc3b_mixed_decoder_layout")
        self.num_streams = n
        self.order = order[:n].tolist()                                            # row -> original stream id
        self.buckets = [((b.sampling_frequency, b.frame_duration), b.first_row, b.n_rows) for b in buckets[:nb.value]]
        self.nf = [b.nf for b in buckets[:nb.value]]
        self.host_pcm_offsets = [int(b.host_pcm_offset) for b in buckets[:nb.value]]
        self.host_pcm_elems = int(elems.value)
        self.max_nf = max(self.nf) if self.nf else 0
        self.max_nbytes = max_nbytes
        size = C.c_size_t(0)
        rc = lib.lc3b_mixed_decoder_workspace_bytes(n, self._sf.ctypes.data, self._fd.ctypes.data, max_nbytes, C.byref(size))
        if rc:
            raise Lc3bError(rc, "lc3b_mixed_decoder_workspace_bytes")
        self.workspace = torch.empty(size.value, dtype=torch.uint8, device=self.device)   # lent for the handle's lifetime
        self._h = C.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = lib.lc3b_mixed_decoder_init(C.byref(self._h), n, self._sf.ctypes.data, self._fd.ctypes.data, max_nbytes,
                                         self.device.index, _ptr(self.workspace), self.workspace.numel(), C.c_void_p(stream))
        if rc:
            raise Lc3bError(rc, "lc3b_mixed_decoder_init")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            native.lib().lc3b_mixed_decoder_destroy(h)

    def _check(self, frames, frame_nbytes, want_cuda, status_out):
        if not (frames.is_cuda == want_cuda and frames.dtype == torch.uint8 and frames.dim() == 2
                and frames.shape[0] == self.num_streams and frames.stride(1) == 1):
            raise Lc3bError(2, "frames: wrong device/dtype/shape")
        for t, what in ((frame_nbytes, "frame_nbytes"), (status_out, "status_out")):
            if t is not None and not (t.is_cuda == want_cuda and t.dtype == torch.int32 and t.numel() == self.num_streams
                                      and t.is_contiguous()):
                raise Lc3bError(2, f"{what}: wrong device/dtype/shape")

    def decode_frames(self, num_bits_per_audio_sample: int, frames: torch.Tensor, frame_nbytes: torch.Tensor,
                      pcm_out: torch.Tensor, status_out: torch.Tensor | None = None) -> None:
        """frames [S, stride] u8, frame_nbytes [S] i32, pcm_out [S, >= max nf] i16 - all CUDA, rows in bucket order.
        Bucket b writes nf_b samples per row; the rest of each row is left untouched."""
        if num_bits_per_audio_sample != 16:
            raise Lc3DecoderError("Only16BitsPerAudioSampleSupported")
        self._check(frames, frame_nbytes, True, status_out)
        if not (pcm_out.is_cuda and pcm_out.dtype == torch.int16 and pcm_out.dim() == 2 and pcm_out.shape[0] == self.num_streams
                and pcm_out.stride(1) == 1 and pcm_out.shape[1] >= self.max_nf):
            raise Lc3bError(2, "pcm_out: wrong device/dtype/shape")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = native.lib().lc3b_mixed_decode_frames(self._h, 16, _ptr(frames), _ptr(frame_nbytes), min(frames.shape[1], self.max_nbytes),
                                                       frames.stride(0), _ptr(pcm_out), pcm_out.stride(0), _ptr(status_out),
                                                       C.c_void_p(stream))
        if rc:
            raise Lc3bError(rc, "lc3b_mixed_decode_frames")

    # ------------------------------------------------------------------ host buffers
    def alloc_host_pcm(self) -> torch.Tensor:
        """One pinned int16 buffer holding every bucket's rows DENSELY ([rows_b, nf_b] at host_pcm_offsets[b]): the read-back
        is a single linear copy.  host_pcm_views() slices it per bucket."""
        return torch.empty(self.host_pcm_elems, dtype=torch.int16).pin_memory()

    def host_pcm_views(self, buf: torch.Tensor) -> list[torch.Tensor]:
        return [buf[off:off + count * nf].view(count, nf)
                for off, (_, _, count), nf in zip(self.host_pcm_offsets, self.buckets, self.nf)]

    def set_host_pipelining(self, on: bool) -> None:
        rc = native.lib().lc3b_mixed_decoder_set_host_pipelining(self._h, int(bool(on)))
        if rc:
            raise Lc3bError(rc, "lc3b_mixed_decoder_set_host_pipelining")

    def set_dequant_mode(self, mode: int) -> None:
        rc = native.lib().lc3b_mixed_decoder_set_dequant_mode(self._h, mode)
        if rc:
            raise Lc3bError(rc, "lc3b_mixed_decoder_set_dequant_mode")

    def set_graph_mode(self, on: bool) -> None:
        rc = native.lib().lc3b_mixed_decoder_set_graph_mode(self._h, int(bool(on)))
        if rc:
            raise Lc3bError(rc, "lc3b_mixed_decoder_set_graph_mode")

    def decode_frames_host(self, num_bits_per_audio_sample: int, frames: torch.Tensor, frame_nbytes: torch.Tensor,
                           pcm_out: torch.Tensor, status_out: torch.Tensor | None = None) -> None:
        """Same call with HOST tensors (pinned): frames [S, stride] u8 and frame_nbytes [S] i32 in bucket order, pcm_out the
        dense buffer of alloc_host_pcm().  With host pipelining on, the PCM copy overlaps the next call; join with host_fence()."""
        if num_bits_per_audio_sample != 16:
            raise Lc3DecoderError("Only16BitsPerAudioSampleSupported")
        self._check(frames, frame_nbytes, False, status_out)
        if not (not pcm_out.is_cuda and pcm_out.dtype == torch.int16 and pcm_out.is_contiguous()
                and pcm_out.numel() >= self.host_pcm_elems):
            raise Lc3bError(2, "pcm_out: want the dense host buffer of alloc_host_pcm()")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = native.lib().lc3b_mixed_decode_frames_host(self._h, 16, _ptr(frames), _ptr(frame_nbytes), min(frames.shape[1], self.max_nbytes),
                                                            frames.stride(0), _ptr(pcm_out), _ptr(status_out), C.c_void_p(stream))
        if rc:
            raise Lc3bError(rc, "lc3b_mixed_decode_frames_host")

    def host_fence(self) -> None:
        """Make the current stream wait for every outstanding pipelined PCM copy."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = native.lib().lc3b_mixed_decoder_host_fence(self._h, C.c_void_p(stream))
        if rc:
            raise Lc3bError(rc, "lc3b_mixed_decoder_host_fence")
