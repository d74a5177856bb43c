// File-level callers of the hot path, after the reference's examples, on the C++ host mirror:
//   lc3b_codec_file encode in.wav out.lc3 <fs> <ms> <nbytes_per_channel>     examples/encode.rs:37-116
//   lc3b_codec_file decode in.lc3 out.wav <fs> <ms> <nbytes_per_channel> <channels>   examples/decode.rs:37-123
//   lc3b_codec_file wavtest                                                  wav.rs:130-148 (no GPU needed)
// Channels of the file are the streams of one batched handle; one call per frame period.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "../include/lc3b.hpp"
#include "../include/lc3b_file.hpp"

using namespace lc3b;

static std::vector<uint8_t> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static SamplingFrequency fs_of(int hz) {
    switch (hz) {
        case 8000: return SamplingFrequency::Hz8000;
        case 16000: return SamplingFrequency::Hz16000;
        case 24000: return SamplingFrequency::Hz24000;
        case 32000: return SamplingFrequency::Hz32000;
        case 44100: return SamplingFrequency::Hz44100;
        case 48000: return SamplingFrequency::Hz48000;
    }
    std::fprintf(stderr, "unsupported sampling frequency %d\n", hz);
    std::exit(2);
}
static FrameDuration dur_of(const std::string& ms) { return ms == "7.5" ? FrameDuration::SevenPointFiveMs : FrameDuration::TenMs; }
#define CUDA_OK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(3); } } while (0)

static int wavtest() {
    const uint8_t buffer[] = {0x52, 0x49, 0x46, 0x46, 0x16, 0x29, 0x0B, 0x00, 0x57, 0x41, 0x56, 0x45, 0x66, 0x6D, 0x74, 0x20, 0x10, 0x00,
                              0x00, 0x00, 0x01, 0x00, 0x02, 0x00, 0x44, 0xAC, 0x00, 0x00, 0x10, 0xB1, 0x02, 0x00, 0x04, 0x00, 0x10, 0x00,
                              0x64, 0x61, 0x74, 0x61, 0x70, 0x28, 0x0B, 0x00, 0x00};
    wav::WavHeader h;
    if (wav::read_header(buffer, sizeof(buffer), &h)) return 1;
    if (h.num_channels != 2 || h.sample_rate != 44100 || h.byte_rate != 176400 || h.block_align != 4 || h.bits_per_sample != 16 ||
        h.data_size != 731248 || h.data_start_position != 44)
        return 1;
    uint8_t out[44];
    size_t n = 0;
    if (wav::write_header(h, out, sizeof(out), &n) || n != 44) return 1;
    for (int i = 0; i < 44; i++) if (out[i] != buffer[i]) return 1;
    if (!wav::write_header(h, out, 10, &n)) return 1;                    // WriteHeaderBufferTooSmall
    if (wav::read_header(buffer, 20, &h) != wav::WavError::ReadHeaderInvalidHeaderLength) return 1;
    std::puts("wavtest ok");
    return 0;
}

static int encode(const char* wav_name, const char* lc3_name, int hz, const std::string& ms, size_t nbytes) {
    const std::vector<uint8_t> in = slurp(wav_name);
    wav::WavHeader wh;
    if (auto e = wav::read_header(in.data(), in.size(), &wh)) { std::fprintf(stderr, "wav error %d\n", (int)*e); return 2; }
    const size_t nch = wh.num_channels;
    const SamplingFrequency f = fs_of(hz);
    const FrameDuration d = dur_of(ms);
    const size_t ws_bytes = Lc3BatchEncoder::calc_working_buffer_lengths(nch, d, f, nbytes);
    void* ws = nullptr;
    CUDA_OK(cudaMalloc(&ws, ws_bytes));
    Lc3BatchEncoder enc(nch, d, f, ws, ws_bytes, nbytes);
    const size_t nf = (size_t)enc.config.nf, bytes_per_frame = nf * nch * 2;
    int16_t* samples = nullptr;
    uint8_t* bits = nullptr;
    CUDA_OK(cudaMallocHost((void**)&samples, nf * nch * sizeof(int16_t)));
    CUDA_OK(cudaMallocHost((void**)&bits, nch * nbytes));
    std::ofstream out(lc3_name, std::ios::binary);
    size_t periods = 0;
    for (size_t cur = wh.data_start_position; cur < in.size(); cur += bytes_per_frame, periods++) {   // encode.rs:73
        const size_t avail = std::min(in.size() - cur, bytes_per_frame) / (2 * nch);
        deinterleave(in.data() + cur, avail, nf, nch, samples);
        enc.encode_frames(Residency::Host, samples, bits, nbytes, nbytes);
        CUDA_OK(cudaStreamSynchronize(nullptr));
        out.write((const char*)bits, (std::streamsize)(nch * nbytes));
    }
    std::printf("encoded %zu frame periods x %zu channels\n", periods, nch);
    cudaFreeHost(samples); cudaFreeHost(bits); cudaFree(ws);
    return 0;
}

static int decode(const char* lc3_name, const char* wav_name, int hz, const std::string& ms, size_t nbytes, size_t nch) {
    const std::vector<uint8_t> in = slurp(lc3_name);
    const SamplingFrequency f = fs_of(hz);
    const FrameDuration d = dur_of(ms);
    const size_t ws_bytes = Lc3BatchDecoder::calc_working_buffer_lengths(nch, d, f, nbytes);
    void* ws = nullptr;
    CUDA_OK(cudaMalloc(&ws, ws_bytes));
    Lc3BatchDecoder dec(nch, d, f, ws, ws_bytes, nbytes);
    const size_t nf = (size_t)dec.config.nf;
    const Lc3File lf{nch, nbytes};
    const size_t periods = lf.frame_periods(in.size());
    // the whole file in ONE time-parallel call: channel-major staging [nch][periods][nbytes]
    std::vector<uint8_t> staged(nch * periods * nbytes);
    for (size_t p = 0; p < periods; p++)
        for (size_t c = 0; c < nch; c++)
            std::memcpy(staged.data() + (c * periods + p) * nbytes, lf.channel_frame(in.data(), p, c), nbytes);
    uint8_t* d_frames = nullptr;
    int16_t* d_pcm = nullptr;
    void* d_scratch = nullptr;
    const size_t scratch_bytes = periods ? dec.multi_scratch_bytes(periods) : 0;
    std::vector<int16_t> by_channel(nch * periods * nf);
    std::vector<uint8_t> pcm(periods * nf * nch * 2);
    if (periods) {
        CUDA_OK(cudaMalloc((void**)&d_frames, staged.size()));
        CUDA_OK(cudaMalloc((void**)&d_pcm, by_channel.size() * sizeof(int16_t)));
        CUDA_OK(cudaMalloc(&d_scratch, scratch_bytes));
        CUDA_OK(cudaMemcpy(d_frames, staged.data(), staged.size(), cudaMemcpyHostToDevice));
        if (dec.decode_stream_frames(16, d_frames, nbytes, periods, d_pcm, d_scratch, scratch_bytes)) {
            std::fprintf(stderr, "decoder error\n");
            return 2;
        }
        CUDA_OK(cudaMemcpy(by_channel.data(), d_pcm, by_channel.size() * sizeof(int16_t), cudaMemcpyDeviceToHost));
        std::vector<int16_t> period(nch * nf);
        for (size_t p = 0; p < periods; p++) {
            for (size_t c = 0; c < nch; c++) std::memcpy(period.data() + c * nf, by_channel.data() + (c * periods + p) * nf, nf * sizeof(int16_t));
            interleave(period.data(), nf, nch, pcm.data() + p * nf * nch * 2);
        }
        cudaFree(d_frames); cudaFree(d_pcm); cudaFree(d_scratch);
    }
    wav::WavHeader wh;
    wh.num_channels = nch;
    wh.sample_rate = (size_t)dec.config.fs;
    wh.bits_per_sample = 16;
    wh.block_align = nch * 2;
    wh.byte_rate = wh.sample_rate * wh.block_align;
    wh.data_size = pcm.size();                                            // (the reference's example leaves the sizes at 0)
    wh.data_start_position = wav::FULL_WAV_HEADER_LEN;
    wh.data_with_header_size = pcm.size() + wav::FULL_WAV_HEADER_LEN - wav::RIFF_HEADER_ONLY_LEN;
    uint8_t hdr[wav::FULL_WAV_HEADER_LEN];
    wav::write_header(wh, hdr, sizeof(hdr), nullptr);
    std::ofstream out(wav_name, std::ios::binary);
    out.write((const char*)hdr, sizeof(hdr));
    out.write((const char*)pcm.data(), (std::streamsize)pcm.size());
    std::printf("decoded %zu frame periods x %zu channels in one call\n", periods, nch);
    cudaFree(ws);
    return 0;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "";
    try {
        if (mode == "wavtest") return wavtest();
        if (mode == "encode" && argc == 7) return encode(argv[2], argv[3], std::atoi(argv[4]), argv[5], (size_t)std::atoi(argv[6]));
        if (mode == "decode" && argc == 8)
            return decode(argv[2], argv[3], std::atoi(argv[4]), argv[5], (size_t)std::atoi(argv[6]), (size_t)std::atoi(argv[7]));
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 4;
    }
    std::fprintf(stderr, "usage: %s encode in.wav out.lc3 fs ms nbytes | decode in.lc3 out.wav fs ms nbytes channels | wavtest\n", argv[0]);
    return 2;
}
