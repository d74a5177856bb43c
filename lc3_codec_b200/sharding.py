"""One batch over the GPUs of a box: streams are independent (src/decoder/lc3_decoder.rs:62-69), so
shard g of G owns the contiguous block of stream ids s with floor(s * G / total) == g.  No collective touches the data
path.

`Lc3ShardedBatchDecoder` / `Lc3ShardedBatchEncoder` mirror the C ABI's lc3b_sharded_* handles (include/lc3b.h): one
process, one host thread + CUDA stream + device workspace per GPU, host (pinned) buffers for the whole batch in, whole
batch out.  `shard_range` / `owner_of` state the same partition for the one-process-per-GPU launch bench.py uses under
torchrun."""
from __future__ import annotations

import ctypes as C

import torch

from . import native
from .decoder import Lc3DecoderError, _ptr
from .native import FrameDuration, Lc3bError, SamplingFrequency


def shard_range(total_streams: int, rank: int, world: int) -> tuple[int, int]:
    """(first stream id, count) owned by `rank`; blocks differ by at most one stream and tile [0, total)."""
    if not (0 <= rank < world) or total_streams < 0:
        raise ValueError("bad rank/world/total")
    start = -(-rank * total_streams // world)            # ceil(rank * total / world)
    end = -(-(rank + 1) * total_streams // world)
    return start, end - start


def owner_of(stream: int, total_streams: int, world: int) -> int:
    if total_streams <= 0 or world <= 0 or not (0 <= stream < total_streams):
        raise ValueError("bad stream/total/world")
    return stream * world // total_streams


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Device time of a multi-rank step is the slowest rank's (bench.py timing rule)."""
    if dist is None or not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def pinned_empty(shape, dtype) -> torch.Tensor:
    """Page-locked host tensor usable from every device (cudaHostAllocPortable)."""
    return torch.empty(shape, dtype=dtype).pin_memory()


class _Sharded:
    _kind = "decoder"

    def __init__(self, num_streams: int, frame_duration: FrameDuration, sampling_frequency: SamplingFrequency,
                 max_nbytes: int = 400, devices: list[int] | None = None):
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        self.num_streams, self.max_nbytes, self.devices = num_streams, max_nbytes, list(devices)
        self.config = native.config(sampling_frequency, frame_duration)
        self.nf = self.config.nf
        arr = (C.c_int32 * len(devices))(*devices)
        self._h = C.c_void_p()
        rc = getattr(native.lib(), f"lc3b_sharded_{self._kind}_create")(C.byref(self._h), num_streams, int(frame_duration),
                                                                         int(sampling_frequency), max_nbytes, arr, len(devices))
        if rc:
            raise Lc3bError(rc, f"lc3b_sharded_{self._kind}_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            getattr(native.lib(), f"lc3b_sharded_{self._kind}_destroy")(h)

    def shards(self) -> list[tuple[int, int, int]]:
        """[(device, first stream, stream count)] per shard."""
        lib = native.lib()
        out = []
        for g in range(getattr(lib, f"lc3b_sharded_{self._kind}_n_shards")(self._h)):
            d, f, n = C.c_int32(0), C.c_int32(0), C.c_int32(0)
            rc = getattr(lib, f"lc3b_sharded_{self._kind}_shard")(self._h, g, C.byref(d), C.byref(f), C.byref(n))
            if rc:
                raise Lc3bError(rc, f"lc3b_sharded_{self._kind}_shard")
            out.append((d.value, f.value, n.value))
        return out

    def wait(self) -> None:
        """Join every outstanding call on every GPU; outputs are valid afterwards."""
        rc = getattr(native.lib(), f"lc3b_sharded_{self._kind}_wait")(self._h)
        if rc:
            raise Lc3bError(rc, f"lc3b_sharded_{self._kind}_wait")


class Lc3ShardedBatchDecoder(_Sharded):
    """decode_frame for every stream of a batch spread over several GPUs (lc3b_sharded_decoder_*)."""
    _kind = "decoder"

    def set_min_nbytes(self, min_nbytes: int) -> None:
        """lc3b_decoder_set_min_nbytes on every shard (joins outstanding calls first)."""
        rc = native.lib().lc3b_sharded_decoder_set_min_nbytes(self._h, int(min_nbytes))
        if rc:
            raise Lc3bError(rc, "lc3b_sharded_decoder_set_min_nbytes")

    def decode_frames_host(self, num_bits_per_audio_sample: int, frames: torch.Tensor, pcm_out: torch.Tensor,
                           frame_nbytes: torch.Tensor | None = None, nbytes: int | None = None,
                           status_out: torch.Tensor | None = None) -> None:
        """HOST tensors for the whole batch: frames [S, stride] u8, pcm_out [S, >= nf] i16 (pinned).  Returns once every
        GPU's thread has the call; wait() joins."""
        if num_bits_per_audio_sample != 16:
            raise Lc3DecoderError("Only16BitsPerAudioSampleSupported")
        for t, dt, what in ((frames, torch.uint8, "frames"), (pcm_out, torch.int16, "pcm_out")):
            if t.is_cuda or t.dtype != dt or t.dim() != 2 or t.shape[0] != self.num_streams or t.stride(1) != 1:
                raise Lc3bError(2, f"{what}: wrong device/dtype/shape")
        if pcm_out.shape[1] < self.nf:
            raise Lc3bError(2, "pcm_out: too few samples per stream")
        for t, what in ((frame_nbytes, "frame_nbytes"), (status_out, "status_out")):
            if t is not None and (t.is_cuda or t.dtype != torch.int32 or t.numel() != self.num_streams or not t.is_contiguous()):
                raise Lc3bError(2, f"{what}: wrong device/dtype/shape")
        nb = frames.shape[1] if nbytes is None else nbytes
        rc = native.lib().lc3b_sharded_decode_frames_host(self._h, 16, _ptr(frames), _ptr(frame_nbytes), nb, frames.stride(0),
                                                          _ptr(pcm_out), pcm_out.stride(0), _ptr(status_out))
        if rc:
            raise Lc3bError(rc, "lc3b_sharded_decode_frames_host")


class Lc3ShardedBatchEncoder(_Sharded):
    """encode_frame for every stream of a batch spread over several GPUs (lc3b_sharded_encoder_*)."""
    _kind = "encoder"

    def encode_frames_host(self, pcm_in: torch.Tensor, frames_out: torch.Tensor) -> None:
        for t, dt, what in ((pcm_in, torch.int16, "pcm_in"), (frames_out, torch.uint8, "frames_out")):
            if t.is_cuda or t.dtype != dt or t.dim() != 2 or t.shape[0] != self.num_streams or t.stride(1) != 1:
                raise Lc3bError(2, f"{what}: wrong device/dtype/shape")
        if pcm_in.shape[1] != self.nf:
            raise Lc3bError(2, "pcm_in: wrong number of samples per stream")
        rc = native.lib().lc3b_sharded_encode_frames_host(self._h, _ptr(pcm_in), pcm_in.stride(0), _ptr(frames_out),
                                                          frames_out.shape[1], frames_out.stride(0))
        if rc:
            raise Lc3bError(rc, "lc3b_sharded_encode_frames_host")
