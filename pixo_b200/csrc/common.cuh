// common.cuh — context object, error plumbing and small device helpers shared by the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/pixo_b200.h"

namespace pixo {

constexpr int kHistWords = 536;  // dc_lum[12] dc_chrom[12] ac_lum[256] ac_chrom[256]

// Per-table divisor / reciprocal pairs, passed by value as a __grid_constant__ kernel
// parameter so that every use is a constant-bank operand with a static offset.
// r[i] = RN(1/d[i]); q = fma(fma(-x*r, d, x), r, x*r) == RN(x/d) for every d in 1..255 and
// every binary32 x (proved exhaustively by tools/verify_div.c).
struct QuantTab {
    float lum_d[64], lum_r[64], chr_d[64], chr_r[64];
};

struct Scratch {
    void *ptr = nullptr;
    size_t cap = 0;
};

// A few persistent host threads per context for the one host-side job that is worth spreading:
// copying a caller's ordinary (pageable) memory into / out of the pinned staging ring while the DMA
// engine drains it.  Created on first use; the threads sleep on a condition variable between calls.
class HostPool {
public:
    explicit HostPool(int nthreads);
    ~HostPool();
    // fn(job) for job in [0, njobs), on the pool's threads and the caller; returns when all are done
    void run(int njobs, const std::function<void(int)> &fn);
    int size() const { return (int)threads_.size(); }

private:
    void worker();
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> next_{0};
    int njobs_ = 0, active_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

}  // namespace pixo

struct pixo_b200_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of the next group of frames
    cudaStream_t d2h_stream = nullptr;   // D2H of finished scan bytes
    int sm_count = 0;
    int host_threads = 0;
    uint64_t launches = 0;
    uint64_t host_fallbacks = 0;   // frames finished by the host entropy coder (see encode_frames)
    size_t scan_cap_override = 0;  // device scan bytes per frame; 0 = the built-in heuristic
    bool gpu_retry = true;         // re-run k_huff with the exact size when the heuristic was too small
    std::string err;
    // reusable scratch (device + pinned host)
    pixo::Scratch d_in, d_y, d_cb, d_cr, d_misc, d_out, d_ent, d_coef, d_retry, d_raw;
    pixo::Scratch h_in, h_out, h_misc;
    std::vector<cudaEvent_t> events;
    std::vector<cudaEvent_t> stage_events;  // one per pinned staging slot of h2d_copy
    pixo::HostPool *pool = nullptr;         // see HostPool
    bool no_segments = false;               // entropy stage: never cut images into segments (retry path)
    // How the bands coded by pixo_b200_jpeg_band_entropy_dev were cut into segments, keyed by the
    // caller's raw buffer (which holds the segments' strings, bit counts and tails until the splice).
    struct BandInfo {
        uint32_t segments = 1, bpm = 0, y_per_mcu = 1;
        bool has_chroma = true;
        uint64_t mcus = 0;
    };
    std::unordered_map<const void *, BandInfo> bands;
    uint32_t last_band_segments = 1;        // of the latest launch_jpeg_entropy(raw) call
};

namespace pixo {

int set_error(pixo_b200_ctx *ctx, int code, const char *fmt, ...);
int cuda_fail(pixo_b200_ctx *ctx, cudaError_t e, const char *what);
int ensure_dev(pixo_b200_ctx *ctx, Scratch &s, size_t bytes);
int ensure_pinned(pixo_b200_ctx *ctx, Scratch &s, size_t bytes);

#define PIXO_CUDA(ctx, call)                                            \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return ::pixo::cuda_fail(ctx, e__, #call); \
    } while (0)

#define PIXO_TRY(expr)                 \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != 0) return rc__;    \
    } while (0)

// ---- launchers implemented in the .cu files ----
int launch_jpeg_transform(pixo_b200_ctx *ctx, const uint8_t *d_pixels, size_t pixel_stride,
                          uint32_t n_images, uint32_t w, uint32_t h, uint32_t color_type,
                          uint32_t subsampling, const float *lum_q, const float *chr_q,
                          int16_t *d_y, size_t y_stride, int16_t *d_cb, int16_t *d_cr,
                          size_t c_stride, uint32_t flags);
int launch_jpeg_histogram(pixo_b200_ctx *ctx, const int16_t *d_y, size_t y_stride,
                          const int16_t *d_cb, const int16_t *d_cr, size_t c_stride,
                          uint32_t n_images, size_t ny, size_t nc, uint32_t blocks_y_per_mcu,
                          uint32_t restart_interval, bool zigzag_in, uint64_t *d_hist,
                          const int *dc_seed = nullptr);
int launch_png_filter(pixo_b200_ctx *ctx, const uint8_t *d_data, size_t in_stride,
                      uint32_t n_images, uint32_t width, uint32_t height, size_t row_bytes,
                      uint32_t bpp, uint32_t strategy, uint8_t *d_out, size_t out_stride,
                      uint32_t *d_adler);
int launch_png_filter_rows(pixo_b200_ctx *ctx, const uint8_t *d_data, size_t in_stride, uint32_t n_images,
                           uint32_t width, uint32_t height, size_t row_bytes, uint32_t bpp, uint32_t strategy,
                           uint8_t *d_out, size_t out_stride, uint32_t *d_adler, const uint8_t *d_above,
                           uint32_t rule_height);
int launch_adler32(pixo_b200_ctx *ctx, const uint8_t *d_data, size_t len, uint32_t *d_out);

struct FrameGeometry;
struct HuffTables;
size_t entropy_scratch_bytes(uint32_t n, const FrameGeometry &g, uint32_t restart_interval);
int launch_jpeg_entropy(pixo_b200_ctx *ctx, const int16_t *d_y, size_t y_stride, const int16_t *d_cb,
                        const int16_t *d_cr, size_t c_stride, uint32_t n, const FrameGeometry &g,
                        const HuffTables &t, uint32_t restart_interval, uint8_t *d_scratch, uint8_t *d_out,
                        uint64_t out_cap, uint64_t **d_out_len, uint32_t **d_overflow,
                        const int *dc_seed = nullptr, uint64_t **d_raw_tail = nullptr);
size_t splice_scratch_bytes(uint64_t nbits);
int launch_band_splice_segments(pixo_b200_ctx *ctx, const uint8_t *d_raw, uint64_t base_bit, uint32_t base_tail,
                                bool last, uint8_t *d_scratch, uint8_t *d_out, uint64_t out_cap,
                                uint64_t **d_out_len, uint32_t **d_overflow);
int launch_band_entropy_async(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb, const int16_t *d_cr,
                              const FrameGeometry &g, const HuffTables &t, const int *d_seed, uint8_t *d_raw,
                              uint64_t raw_cap, uint64_t *d_bits_tail, uint32_t *d_flags);
int launch_band_splice_async(pixo_b200_ctx *ctx, const uint8_t *d_raw, const uint64_t *d_offset, uint8_t *d_out,
                             uint64_t out_cap, uint64_t *d_out_len, uint32_t *d_flags);
// bytes a band's raw buffer needs so that the band can be coded in segments (0: never segmented)
size_t band_raw_bytes_segmented(const FrameGeometry &g);
int launch_splice(pixo_b200_ctx *ctx, const uint8_t *d_raw, uint64_t nbits, uint32_t phase, uint32_t tail_in,
                  bool last, uint8_t *d_scratch, uint8_t *d_out, uint64_t out_cap, uint64_t **d_out_len,
                  uint32_t **d_overflow);

}  // namespace pixo
