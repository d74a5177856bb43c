// lc3b - container / file adapters either side of the hot path, C++ host side.
//
//   wav::read_header / wav::write_header    src/common/wav.rs:45-123: the 44-byte canonical PCM WAV header, same field
//                                           set (WavHeader) and the same error cases (WavError)
//   Lc3File                                 the headerless framing the reference's examples use
//                                           (examples/encode.rs:73-116, examples/decode.rs:93-123): per frame period,
//                                           `nbytes` bytes for channel 0, then channel 1, ...
// Pure host code, no CUDA.
#pragma once
#include <cstdint>
#include <cstring>
#include <optional>
#include <vector>

namespace lc3b {
namespace wav {

enum class WavError {                    // wav.rs:8-17
    WriteHeaderBufferTooSmall,
    ReadHeaderInvalidHeaderLength,
    ReadHeaderChunkIdNotRIFF,
    ReadHeaderFormatNotWAVE,
    ReadHeaderSubChunk1IdNotFmt,
    ReadHeaderInvalidPcmHeaderLength,
    ReadHeaderAudioFormatNotPcm,
    ReadHeaderMissingDataSection,
};

constexpr size_t RIFF_HEADER_ONLY_LEN = 8;
constexpr size_t FULL_WAV_HEADER_LEN = 44;

struct WavHeader {                       // wav.rs:25-43
    size_t num_channels = 0;
    size_t sample_rate = 0;
    size_t byte_rate = 0;                // SampleRate * NumChannels * BitsPerSample / 8
    size_t block_align = 0;              // NumChannels * BitsPerSample / 8
    size_t bits_per_sample = 0;
    size_t data_size = 0;                // NumSamples * NumChannels * BitsPerSample / 8
    size_t data_start_position = 0;      // position of the first byte of data
    size_t data_with_header_size = 0;    // bytes of the entire file excluding the first 8
};

namespace detail {
inline void put_u16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
inline void put_u32(uint8_t* p, uint32_t v) { put_u16(p, v & 0xffffu); put_u16(p + 2, v >> 16); }
inline uint32_t get_u16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t get_u32(const uint8_t* p) { return get_u16(p) | (get_u16(p + 2) << 16); }
}  // namespace detail

// wav.rs:45-68.  Returns the number of bytes written (44) or the error.
inline std::optional<WavError> write_header(const WavHeader& h, uint8_t* buf, size_t len, size_t* written) {
    if (len < FULL_WAV_HEADER_LEN) return WavError::WriteHeaderBufferTooSmall;
    std::memcpy(buf, "RIFF", 4);
    detail::put_u32(buf + 4, (uint32_t)h.data_with_header_size);
    std::memcpy(buf + 8, "WAVE", 4);
    std::memcpy(buf + 12, "fmt ", 4);
    detail::put_u32(buf + 16, 16);                       // PCM header length
    detail::put_u16(buf + 20, 1);                        // PCM
    detail::put_u16(buf + 22, (uint32_t)h.num_channels);
    detail::put_u32(buf + 24, (uint32_t)h.sample_rate);
    detail::put_u32(buf + 28, (uint32_t)h.byte_rate);
    detail::put_u16(buf + 32, (uint32_t)h.block_align);
    detail::put_u16(buf + 34, (uint32_t)h.bits_per_sample);
    std::memcpy(buf + 36, "data", 4);
    detail::put_u32(buf + 40, (uint32_t)h.data_size);
    if (written) *written = FULL_WAV_HEADER_LEN;
    return std::nullopt;
}

// wav.rs:70-123, same checks in the same order; a "LIST" chunk in place of "data" moves the data start by 4 bytes
// exactly like the reference (it does not skip the LIST payload).
inline std::optional<WavError> read_header(const uint8_t* buf, size_t len, WavHeader* out) {
    if (len < FULL_WAV_HEADER_LEN) return WavError::ReadHeaderInvalidHeaderLength;
    if (std::memcmp(buf, "RIFF", 4) != 0) return WavError::ReadHeaderChunkIdNotRIFF;
    if (std::memcmp(buf + 8, "WAVE", 4) != 0) return WavError::ReadHeaderFormatNotWAVE;
    if (std::memcmp(buf + 12, "fmt ", 4) != 0) return WavError::ReadHeaderSubChunk1IdNotFmt;
    if (detail::get_u32(buf + 16) != 16) return WavError::ReadHeaderInvalidPcmHeaderLength;
    if (detail::get_u16(buf + 20) != 1) return WavError::ReadHeaderAudioFormatNotPcm;
    WavHeader h;
    h.data_with_header_size = detail::get_u32(buf + 4);
    h.num_channels = detail::get_u16(buf + 22);
    h.sample_rate = detail::get_u32(buf + 24);
    h.byte_rate = detail::get_u32(buf + 28);
    h.block_align = detail::get_u16(buf + 32);
    h.bits_per_sample = detail::get_u16(buf + 34);
    if (std::memcmp(buf + 36, "data", 4) == 0) {
        h.data_size = detail::get_u32(buf + 40);
        h.data_start_position = FULL_WAV_HEADER_LEN;
    } else if (std::memcmp(buf + 36, "LIST", 4) == 0) {
        h.data_size = detail::get_u32(buf + 40);
        h.data_start_position = FULL_WAV_HEADER_LEN + 4;
    } else {
        return WavError::ReadHeaderMissingDataSection;
    }
    *out = h;
    return std::nullopt;
}

}  // namespace wav

// Headerless .lc3 framing of the reference's examples: frame period p, channel c at byte (p * num_channels + c) * nbytes.
struct Lc3File {
    size_t num_channels, nbytes;
    // examples/decode.rs:93-123 stops as soon as a channel frame would END AT or beyond the end of the file
    // (`if to_index >= buf_in_full.len() { return }`), i.e. it never decodes the file's very last channel frame.
    // `reference_quirk` reproduces that count; otherwise every complete frame period is counted.
    size_t frame_periods(size_t file_len, bool reference_quirk = false) const {
        const size_t period = num_channels * nbytes;
        if (period == 0) return 0;
        size_t n = file_len / period;
        if (reference_quirk && n > 0 && n * period == file_len) n -= 1;
        return n;
    }
    // gather one frame period into the stream-major layout the batched decoder takes: out[c][nbytes]
    const uint8_t* channel_frame(const uint8_t* file, size_t period, size_t channel) const {
        return file + (period * num_channels + channel) * nbytes;
    }
};

// i16 PCM <-> interleaved little-endian bytes (examples/encode.rs:88-103, examples/decode.rs:112-118)
inline void deinterleave(const uint8_t* le_bytes, size_t frames_available, size_t nf, size_t num_channels, int16_t* by_channel) {
    for (size_t ch = 0; ch < num_channels; ch++)
        for (size_t i = 0; i < nf; i++) {
            int16_t v = 0;                                   // zero padding past the end of the file (encode.rs:93-97)
            if (i < frames_available) {
                const uint8_t* p = le_bytes + 2 * (i * num_channels + ch);
                v = (int16_t)((uint16_t)p[0] | ((uint16_t)p[1] << 8));
            }
            by_channel[ch * nf + i] = v;
        }
}
inline void interleave(const int16_t* by_channel, size_t nf, size_t num_channels, uint8_t* le_bytes) {
    for (size_t i = 0; i < nf; i++)
        for (size_t ch = 0; ch < num_channels; ch++) {
            const uint16_t v = (uint16_t)by_channel[ch * nf + i];
            le_bytes[2 * (i * num_channels + ch)] = (uint8_t)v;
            le_bytes[2 * (i * num_channels + ch) + 1] = (uint8_t)(v >> 8);
        }
}

}  // namespace lc3b
