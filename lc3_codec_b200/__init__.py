"""lc3_codec_b200 - batched LC3 decode/encode on NVIDIA B200 behind the reference's decode_frame/encode_frame surface.

The compute path is the in-tree CUDA library (csrc/ -> liblc3b.so, C ABI in include/lc3b.h).  This package is
the host-side mirror of the reference's interface for that path; torch is used only for device memory and
streams.  There is no CPU fallback: importing the native module without the built library raises.
"""
from .native import FrameDuration, Lc3bError, SamplingFrequency, lib, lib_path  # noqa: F401
from .decoder import Lc3BatchDecoder, Lc3DecoderError  # noqa: F401
from .encoder import Lc3BatchEncoder  # noqa: F401
from .mixed import Lc3MixedBatchDecoder  # noqa: F401
from .sharding import Lc3ShardedBatchDecoder, Lc3ShardedBatchEncoder  # noqa: F401
