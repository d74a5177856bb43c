"""Batched mirror of the reference's `Lc3Encoder` (src/encoder/lc3_encoder.rs:33, :116-210).

Reference                                   here
------------------------------------------  -----------------------------------------------------------------
Lc3Encoder::calc_working_buffer_lengths     Lc3BatchEncoder.calc_working_buffer_lengths  (bytes of device workspace)
Lc3Encoder::new(num_channels, ..bufs)       Lc3BatchEncoder(num_streams, .., workspace)
encode_frame(channel, samples_in, buf_out)  encode_frames(pcm_in, frames_out)            one frame for EVERY stream
Result<(), Lc3EncoderError> (empty enum)    returns None; cannot fail for valid shapes
assert_eq!(samples_in.len(), nf) panics     raises Lc3bError(LC3B_ERR_INVALID_ARG)
Lc3Encoder::new panics at 8 kHz             raises Lc3bError(LC3B_ERR_INVALID_ARG)
"""
from __future__ import annotations

import ctypes as C

import torch

from . import native
from .native import FrameDuration, Lc3bError, SamplingFrequency


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Lc3BatchEncoder:
    @staticmethod
    def calc_working_buffer_lengths(num_streams: int, frame_duration: FrameDuration,
                                    sampling_frequency: SamplingFrequency, max_nbytes: int = 400) -> int:
        n = C.c_size_t(0)
        rc = native.lib().lc3b_encoder_workspace_bytes(num_streams, int(frame_duration), int(sampling_frequency),
                                                       max_nbytes, C.byref(n))
        if rc:
            raise Lc3bError(rc, "lc3b_encoder_workspace_bytes")
        return n.value

    def __init__(self, num_streams: int, frame_duration: FrameDuration, sampling_frequency: SamplingFrequency,
                 workspace: torch.Tensor, max_nbytes: int = 400):
        if not workspace.is_cuda or workspace.dtype != torch.uint8:
            raise Lc3bError(2, "workspace must be a CUDA uint8 tensor")
        self.num_streams = num_streams
        self.config = native.config(sampling_frequency, frame_duration)
        self.nf, self.ne = self.config.nf, self.config.ne
        self.max_nbytes = max_nbytes
        self.workspace = workspace
        self.device = workspace.device
        self._h = C.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = native.lib().lc3b_encoder_init(C.byref(self._h), num_streams, int(frame_duration), int(sampling_frequency),
                                            max_nbytes, self.device.index or 0, _ptr(workspace), workspace.numel(),
                                            C.c_void_p(stream))
        if rc:
            raise Lc3bError(rc, "lc3b_encoder_init")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            native.lib().lc3b_encoder_destroy(h)

    def encode_frames(self, pcm_in: torch.Tensor, frames_out: torch.Tensor) -> None:
        """pcm_in: CUDA int16 [num_streams, nf]; frames_out: CUDA uint8 [num_streams, nbytes] (nbytes = buf_out.len())."""
        self._call(native.lib().lc3b_encode_frames, True, pcm_in, frames_out)

    def encode_frames_host(self, pcm_in: torch.Tensor, frames_out: torch.Tensor) -> None:
        self._call(native.lib().lc3b_encode_frames_host, False, pcm_in, frames_out)

    def _call(self, fn, want_cuda, pcm_in, frames_out):
        for t, dt, what in ((pcm_in, torch.int16, "pcm_in"), (frames_out, torch.uint8, "frames_out")):
            if t.is_cuda != want_cuda or t.dtype != dt or t.dim() != 2 or t.shape[0] != self.num_streams or t.stride(1) != 1:
                raise Lc3bError(2, f"{what}: wrong device/dtype/shape")
        if pcm_in.shape[1] != self.nf:
            raise Lc3bError(2, f"pcm_in: {pcm_in.shape[1]} samples per stream, need nf = {self.nf}")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = fn(self._h, _ptr(pcm_in), pcm_in.stride(0), _ptr(frames_out), frames_out.shape[1], frames_out.stride(0),
                    C.c_void_p(stream))
        if rc:
            raise Lc3bError(rc, fn.__name__)

    def set_host_pipelining(self, on: bool) -> None:
        """Let the PCM upload of encode_frames_host call i+1 overlap the kernels of call i (include/lc3b.h)."""
        rc = native.lib().lc3b_encoder_set_host_pipelining(self._h, 1 if on else 0)
        if rc != 0:
            raise Lc3bError(rc, "lc3b_encoder_set_host_pipelining")

    def set_graph_mode(self, on: bool) -> None:
        """Issue every call as ONE cached CUDA graph launch (True) or one launch per kernel (False); include/lc3b.h."""
        rc = native.lib().lc3b_encoder_set_graph_mode(self._h, int(bool(on)))
        if rc:
            raise Lc3bError(rc, "lc3b_encoder_set_graph_mode")

    def set_stage_mask(self, mask: int) -> None:
        """Profiling hook (lc3b_encoder_set_stage_mask): bit 0 = MDCT kernel, bit 1 = attack / LTPF analysis kernel,
        bit 2 = SNS kernel, bit 3 = TNS kernel, bit 4 = quantise kernel, bit 5 = bitstream stage; default 63 (all)."""
        rc = native.lib().lc3b_encoder_set_stage_mask(self._h, mask)
        if rc:
            raise Lc3bError(rc, "lc3b_encoder_set_stage_mask")

    def debug_read(self):
        """(xf, e_b, hand, xq) device tensors of the last encode's intermediates (test hook)."""
        S, ne = self.num_streams, self.ne
        xf = torch.empty((S, ne), dtype=torch.float32, device=self.device)
        eb = torch.empty((S, 64), dtype=torch.float32, device=self.device)
        hand = torch.empty((S, 8), dtype=torch.int32, device=self.device)
        xq = torch.empty((S, ne), dtype=torch.int16, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = native.lib().lc3b_encoder_debug_read(self._h, _ptr(xf), _ptr(eb), _ptr(hand), _ptr(xq), C.c_void_p(stream))
        if rc:
            raise Lc3bError(rc, "lc3b_encoder_debug_read")
        return xf, eb, hand, xq
