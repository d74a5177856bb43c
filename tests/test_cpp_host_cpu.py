"""C++ host mirror (include/lc3b.hpp) and file adapters (include/lc3b_file.hpp): build + the reference's WAV test.

No GPU needed: `lc3b_codec_file wavtest` replays src/common/wav.rs:130-148 (can_read_pcm_wav_header) and the header
round trip; the GPU-side file workflows are in test_file_workflow_gpu.py.
"""
import pathlib
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_examples_build_and_wav_header_golden():
    subprocess.run(["make", "-C", str(ROOT / "lc3_codec_b200" / "csrc")], check=True, capture_output=True)
    subprocess.run(["make", "-C", str(ROOT / "examples")], check=True, capture_output=True)
    r = subprocess.run([str(ROOT / "examples" / "_build" / "lc3b_codec_file"), "wavtest"], capture_output=True, text=True)
    assert r.returncode == 0 and "wavtest ok" in r.stdout, r.stdout + r.stderr


def test_cpp_mirror_names_follow_the_reference():
    """The mirror keeps the reference's operator names (lc3_decoder.rs:181-245, lc3_encoder.rs:117-210)."""
    text = (ROOT / "include" / "lc3b.hpp").read_text()
    for name in ("calc_working_buffer_lengths", "decode_frames", "encode_frames", "Only16BitsPerAudioSampleSupported",
                 "SamplingFrequency", "FrameDuration", "Lc3Config"):
        assert name in text
