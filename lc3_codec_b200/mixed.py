"""Mixed-rate batches (BASELINE config 4): streams of different (sampling frequency, frame duration) in one call.

The reference builds one `Lc3Decoder` per configuration (src/decoder/lc3_decoder.rs:181 takes ONE frame duration and
ONE sampling frequency for all its channels); a mixed population is therefore a set of decoders.  This front end does
the same: streams are bucketed by configuration (stable order), each bucket owns an `Lc3BatchDecoder` and a CUDA
stream, and one `decode_frames` call fans out over the buckets concurrently and joins them back on the caller's stream.
Callers lay their per-stream rows out in bucket order (`order` maps sorted position -> original stream id), so every
bucket reads and writes a contiguous row range of the caller's buffers: no gather, no copy.
"""
from __future__ import annotations

import torch

from .decoder import Lc3BatchDecoder
from .native import FrameDuration, SamplingFrequency, config as native_config


class Lc3MixedBatchDecoder:
    def __init__(self, stream_configs: list[tuple[SamplingFrequency, FrameDuration]], max_nbytes: int = 400,
                 device: str | torch.device = "cuda"):
        self.device = torch.device(device)
        keys = [(int(sf), int(fd)) for sf, fd in stream_configs]
        self.order = sorted(range(len(keys)), key=lambda s: (keys[s], s))        # sorted position -> original stream id
        self.buckets = []                                                        # (key, first row, row count)
        pos = 0
        while pos < len(self.order):
            k = keys[self.order[pos]]
            end = pos
            while end < len(self.order) and keys[self.order[end]] == k:
                end += 1
            self.buckets.append((k, pos, end - pos))
            pos = end
        self.decoders, self.streams, self.workspaces, self.nf = [], [], [], []
        for (sf, fd), _, count in self.buckets:
            n = Lc3BatchDecoder.calc_working_buffer_lengths(count, FrameDuration(fd), SamplingFrequency(sf), max_nbytes)
            ws = torch.empty(n, dtype=torch.uint8, device=self.device)
            self.workspaces.append(ws)
            self.decoders.append(Lc3BatchDecoder(count, FrameDuration(fd), SamplingFrequency(sf), ws, max_nbytes))
            self.streams.append(torch.cuda.Stream(device=self.device))
            self.nf.append(native_config(sf, fd).nf)
        self.num_streams = len(keys)
        self.max_nf = max(self.nf) if self.nf else 0

    def decode_frames(self, num_bits_per_audio_sample: int, frames: torch.Tensor, frame_nbytes: torch.Tensor,
                      pcm_out: torch.Tensor, status_out: torch.Tensor | None = None) -> None:
        """frames [S, stride] u8, frame_nbytes [S] i32, pcm_out [S, >= max nf] i16 - all CUDA, rows in bucket order.
        Bucket b writes nf_b samples per row; the rest of each row is left untouched."""
        cur = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(cur)
        for dec, st, (_, first, count) in zip(self.decoders, self.streams, self.buckets):
            st.wait_event(start)
            with torch.cuda.stream(st):
                dec.decode_frames(num_bits_per_audio_sample, frames[first:first + count], pcm_out[first:first + count],
                                  frame_nbytes=frame_nbytes[first:first + count], nbytes=frames.shape[1],
                                  status_out=None if status_out is None else status_out[first:first + count])
            done = torch.cuda.Event()
            done.record(st)
            cur.wait_event(done)

    # ------------------------------------------------------------------ host buffers
    def alloc_host_pcm(self) -> list[torch.Tensor]:
        """One dense pinned int16 [rows_b, nf_b] tensor per bucket (dense rows keep every PCM copy a single linear one)."""
        return [torch.empty((count, nf), dtype=torch.int16).pin_memory() for (_, _, count), nf in zip(self.buckets, self.nf)]

    def set_host_pipelining(self, on: bool) -> None:
        for dec in self.decoders:
            dec.set_host_pipelining(on)

    def decode_frames_host(self, num_bits_per_audio_sample: int, frames: torch.Tensor, frame_nbytes: torch.Tensor,
                           pcm_out: list[torch.Tensor]) -> None:
        """Same call with HOST tensors (pinned): frames [S, stride] u8 and frame_nbytes [S] i32 in bucket order, pcm_out one
        dense [rows_b, nf_b] tensor per bucket (alloc_host_pcm).  Buckets run on their own CUDA streams; with host
        pipelining on, PCM copies overlap the other buckets' kernels and the next call.  Join with host_fence()."""
        cur = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(cur)
        for dec, st, (_, first, count), out in zip(self.decoders, self.streams, self.buckets, pcm_out):
            st.wait_event(start)
            with torch.cuda.stream(st):
                dec.decode_frames_host(num_bits_per_audio_sample, frames[first:first + count], out,
                                       frame_nbytes=frame_nbytes[first:first + count], nbytes=frames.shape[1])

    def host_fence(self) -> None:
        """Make the current stream wait for every bucket's outstanding work (kernels and pipelined PCM copies)."""
        cur = torch.cuda.current_stream(self.device)
        for dec, st in zip(self.decoders, self.streams):
            with torch.cuda.stream(st):
                dec.host_fence()
            done = torch.cuda.Event()
            done.record(st)
            cur.wait_event(done)
