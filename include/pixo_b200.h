/*
 * pixo_b200.h — C ABI of libpixo_b200.so: the B200 (sm_100a) replacement for the data-parallel
 * stages of leerob/pixo's JPEG and PNG encoders.
 *
 * pixo (v0.4.1 @ 437bf63) has no FFI of its own; its only dispatch seam is src/simd/mod.rs
 * (row-granular, 15 KB calls — far too small for a GPU).  This boundary therefore sits one
 * level up, at the reference functions that already take a whole image.
 * Each entry point cites the reference interface it replaces (file:line in the pixo tree).
 * INTEGRATION.md shows the Rust `extern "C"` binding and the call-site patch.
 *
 * Conventions
 *  - plain pointers and sizes only; the caller owns every buffer; the library never retains
 *    or frees caller memory.  `_dev` variants take device pointers and are asynchronous on the
 *    context's stream; the others take host pointers and return after the result is in `out`.
 *  - every function returns a pixo_b200_status (0 = ok).  pixo_b200_last_error(ctx) gives the
 *    message a Rust shim would wrap in Error::CompressionError(String) (src/error.rs:41).
 *    Validation errors mirror the reference's own checks (src/jpeg/mod.rs:333-373,
 *    src/png/mod.rs:442-467).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    PIXO_B200_ERR_CUDA.
 *  - a context owns one CUDA stream plus reusable device/pinned scratch (mirrors the
 *    encode_into buffer-reuse convention, src/jpeg/mod.rs:375-376).  One context per host
 *    thread; contexts are independent and may target different GPUs.
 */
#ifndef PIXO_B200_H
#define PIXO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIXO_B200_VERSION 0x000100

typedef struct pixo_b200_ctx pixo_b200_ctx;

typedef enum {
    PIXO_B200_OK = 0,
    PIXO_B200_ERR_INVALID_QUALITY = 1,     /* Error::InvalidQuality          src/error.rs */
    PIXO_B200_ERR_INVALID_DIMENSIONS = 2,  /* Error::InvalidDimensions */
    PIXO_B200_ERR_IMAGE_TOO_LARGE = 3,     /* Error::ImageTooLarge */
    PIXO_B200_ERR_UNSUPPORTED_COLOR = 4,   /* Error::UnsupportedColorType */
    PIXO_B200_ERR_INVALID_DATA_LENGTH = 5, /* Error::InvalidDataLength */
    PIXO_B200_ERR_INVALID_RESTART = 6,     /* Error::InvalidRestartInterval */
    PIXO_B200_ERR_INVALID_ARGUMENT = 7,    /* null pointer, unknown enum value ... */
    PIXO_B200_ERR_OUTPUT_TOO_SMALL = 8,    /* caller's output capacity insufficient */
    PIXO_B200_ERR_UNSUPPORTED = 9,         /* option outside the hot path (progressive, trellis) */
    PIXO_B200_ERR_CUDA = 10,               /* CUDA runtime/driver failure (no device, launch) */
    PIXO_B200_ERR_OOM = 11                 /* device or pinned allocation failed */
} pixo_b200_status;

/* pixo::ColorType repr(u8) — src/color.rs:8-18 */
enum { PIXO_B200_GRAY = 0, PIXO_B200_GRAY_ALPHA = 1, PIXO_B200_RGB = 2, PIXO_B200_RGBA = 3 };
/* pixo::jpeg::Subsampling — src/jpeg/mod.rs:96-102 */
enum { PIXO_B200_S444 = 0, PIXO_B200_S420 = 1 };
/* pixo::png::FilterStrategy, declaration order — src/png/mod.rs:345-364 */
enum {
    PIXO_B200_FILTER_NONE = 0, PIXO_B200_FILTER_SUB = 1, PIXO_B200_FILTER_UP = 2,
    PIXO_B200_FILTER_AVERAGE = 3, PIXO_B200_FILTER_PAETH = 4, PIXO_B200_FILTER_MINSUM = 5,
    PIXO_B200_FILTER_ADAPTIVE = 6, PIXO_B200_FILTER_ADAPTIVE_FAST = 7,
    PIXO_B200_FILTER_BIGRAMS = 8
};

/* flags for pixo_b200_jpeg_coefficients* */
#define PIXO_B200_COEF_ZIGZAG 1u /* emit each block in zig-zag order (quantize.rs:107-113) */

/* ---- context ---------------------------------------------------------------------------- */
int pixo_b200_version(void);
/* number of visible CUDA devices (0 when none / no driver) */
int pixo_b200_device_count(void);
int pixo_b200_ctx_create(int device, pixo_b200_ctx **out);
void pixo_b200_ctx_destroy(pixo_b200_ctx *ctx);
/* last error message of this context (thread's last error when ctx == NULL); never NULL */
const char *pixo_b200_last_error(const pixo_b200_ctx *ctx);
/* adopt an external CUDA stream (cudaStream_t) for all work of this context; NULL restores
 * the context's own stream */
int pixo_b200_ctx_set_stream(pixo_b200_ctx *ctx, void *cuda_stream);
void *pixo_b200_ctx_stream(pixo_b200_ctx *ctx);
int pixo_b200_ctx_sync(pixo_b200_ctx *ctx);
/* kernels launched by this context since creation (every launch is one of this library's own
 * kernels; memcpy/memset are not counted) */
uint64_t pixo_b200_ctx_launch_count(const pixo_b200_ctx *ctx);
/* number of host threads the host-side entropy coder may use (default: hardware threads) */
int pixo_b200_ctx_set_host_threads(pixo_b200_ctx *ctx, int n);
/* Observability of the one place host code can finish device work.  The GPU entropy stage writes
 * each frame's scan into a device buffer sized by a heuristic (half the raw frame + 64 KiB, x 9/8);
 * a frame that needs more is coded again on the GPU with the exact size (the kernel reports it),
 * and only if that is switched off, or the device stage reports a fault, does the host coder
 * finish the frame from the same GPU coefficient arrays.  host_fallbacks counts those frames
 * since the context was created (0 in normal operation).
 * set_scan_capacity: bytes_per_frame 0 restores the heuristic; gpu_retry 0 disables the second
 * GPU pass (test hook: a tiny capacity with gpu_retry 0 forces the host coder). */
uint64_t pixo_b200_ctx_host_fallbacks(const pixo_b200_ctx *ctx);
int pixo_b200_ctx_set_scan_capacity(pixo_b200_ctx *ctx, size_t bytes_per_frame, int gpu_retry);

/* device / pinned memory helpers so a Rust caller need not link the CUDA runtime */
int pixo_b200_dev_alloc(pixo_b200_ctx *ctx, size_t bytes, void **dptr);
int pixo_b200_dev_free(pixo_b200_ctx *ctx, void *dptr);
int pixo_b200_host_alloc_pinned(pixo_b200_ctx *ctx, size_t bytes, void **hptr);
int pixo_b200_host_free_pinned(pixo_b200_ctx *ctx, void *hptr);
int pixo_b200_upload(pixo_b200_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int pixo_b200_download(pixo_b200_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

/* ---- JPEG ------------------------------------------------------------------------------- */

/* QuantizationTables::with_quality — src/jpeg/quantize.rs:42-89.
 * lum_zz/chr_zz: zig-zag order u8 (what DQT carries); lum/chr: natural order f32 (arithmetic).
 * Any output pointer may be NULL.  Host-only, no device needed. */
void pixo_b200_quant_tables(int quality, uint8_t lum_zz[64], uint8_t chr_zz[64], float lum[64],
                            float chr[64]);

/* Block counts of the coefficient arrays: ny Y blocks, nc blocks per chroma component
 * (src/jpeg/mod.rs:1055-1125). */
int pixo_b200_jpeg_block_counts(uint32_t width, uint32_t height, uint32_t color_type,
                                uint32_t subsampling, size_t *ny, size_t *nc);

/* Replaces compute_all_coefficients — src/jpeg/mod.rs:932-966 (and the per-MCU transform
 * work inlined in encode_scan :1408-1563 and build_optimized_huffman_tables :684-824):
 * extract_block/extract_mcu_420 (:1565-1656) -> color::rgb_to_ycbcr (src/color.rs:60-77) ->
 * dct::dct_2d (src/jpeg/dct.rs:614-700) -> quantize_block (src/jpeg/quantize.rs:99-105).
 * y: ny*64, cb/cr: nc*64 int16 (cb/cr ignored for Gray), natural order unless
 * PIXO_B200_COEF_ZIGZAG, blocks in the reference's MCU order (4:2:0: Y TL,TR,BL,BR per MCU).
 * hist (optional, may be NULL): 536 u64 = dc_lum[12] dc_chrom[12] ac_lum[256] ac_chrom[256],
 * the count_block statistics of src/jpeg/mod.rs:826-860 with no restart interval. */
int pixo_b200_jpeg_coefficients(pixo_b200_ctx *ctx, const uint8_t *pixels, uint32_t width,
                                uint32_t height, uint32_t color_type, uint32_t subsampling,
                                const float lum_q[64], const float chr_q[64], int16_t *y,
                                int16_t *cb, int16_t *cr, uint32_t flags, uint64_t *hist);

/* Same, device pointers, asynchronous on the context's stream.  `n_images` frames of identical
 * geometry: frame i reads d_pixels + i*pixel_stride and writes d_y + i*y_stride (int16
 * elements), d_cb/d_cr + i*c_stride.  d_hist (optional): n_images*536 u64, zeroed by the call. */
int pixo_b200_jpeg_coefficients_dev(pixo_b200_ctx *ctx, const uint8_t *d_pixels,
                                    size_t pixel_stride, uint32_t n_images, uint32_t width,
                                    uint32_t height, uint32_t color_type, uint32_t subsampling,
                                    const float lum_q[64], const float chr_q[64], int16_t *d_y,
                                    size_t y_stride, int16_t *d_cb, int16_t *d_cr,
                                    size_t c_stride, uint32_t flags, uint64_t *d_hist);

/* Replaces pixo::jpeg::encode_into — src/jpeg/mod.rs:328-447 (baseline: encode_scan :1408).
 * GPU: colour/subsample/DCT/quantise, symbol statistics when optimize_huffman, Huffman bit
 * packing with 0xFF stuffing and restart markers (huffman.rs:423-481, src/bits.rs:195-290,
 * mod.rs:1423-1445); host: headers (:449-648) and Huffman table construction
 * (src/jpeg/huffman.rs:100-391).  Byte-identical to the reference.  Only the scan bytes come
 * back over PCIe.  restart_interval 0 = None.  progressive is outside this path
 * (PIXO_B200_ERR_UNSUPPORTED when non-zero); trellis_quant is accepted and ignored, exactly as
 * the reference's baseline encode_scan ignores use_trellis (src/jpeg/mod.rs:1408-1563 always
 * calls quantize_block). */
int pixo_b200_jpeg_encode(pixo_b200_ctx *ctx, const uint8_t *pixels, size_t pixels_len,
                          uint32_t width, uint32_t height, uint32_t color_type, uint32_t quality,
                          uint32_t subsampling, uint32_t restart_interval,
                          uint32_t optimize_huffman, uint32_t progressive, uint32_t trellis_quant,
                          uint8_t *out, size_t out_cap, size_t *out_len);

/* Batch of n_images frames of identical geometry and options (frame i at pixels + i*len).
 * Transfers and kernels of different frames overlap.  out: n_images
 * slots of out_cap_each bytes; out_lens[i] receives each JPEG's length. */
int pixo_b200_jpeg_encode_batch(pixo_b200_ctx *ctx, const uint8_t *pixels, size_t pixels_len_each,
                                uint32_t n_images, uint32_t width, uint32_t height,
                                uint32_t color_type, uint32_t quality, uint32_t subsampling,
                                uint32_t restart_interval, uint32_t optimize_huffman,
                                uint8_t *out, size_t out_cap_each, size_t *out_lens);

/* Device-resident variant of the whole hot path (asynchronous on the context's stream):
 * frame i at d_pixels + i*pixel_stride -> entropy-coded scan bytes (what encode_scan appends
 * between the SOS header and EOI, src/jpeg/mod.rs:1408-1563) at d_scan + i*scan_cap_each, byte
 * count in d_scan_len[i] (the size needed, also when it did not fit); d_overflow[i] != 0 when the
 * frame was not finished: bit 0 scan_cap_each was too small, bit 1 a device fault (spin limit), bit 2 a
 * segment of a frame that is coded in segments (few large frames) outgrew its internal buffer - does
 * not happen for JPEGs smaller than their raw pixels; pixo_b200_jpeg_encode* handle all three.  Baseline, standard Huffman tables, no restart interval.
 * Headers/EOI are the caller's (pixo_b200_jpeg_encode* add them). */
int pixo_b200_jpeg_encode_dev(pixo_b200_ctx *ctx, const uint8_t *d_pixels, size_t pixel_stride,
                              uint32_t n_images, uint32_t width, uint32_t height,
                              uint32_t color_type, uint32_t quality, uint32_t subsampling,
                              uint8_t *d_scan, size_t scan_cap_each, uint64_t *d_scan_len,
                              uint32_t *d_overflow);

/* Entropy-code caller-provided coefficient arrays (host) into a baseline JPEG: the host half of
 * pixo_b200_jpeg_encode on its own (src/jpeg/mod.rs:395-447,1408-1563 consuming arrays shaped
 * like compute_all_coefficients' result).  Host-only, no device needed. */
int pixo_b200_jpeg_entropy_encode(pixo_b200_ctx *ctx, const int16_t *y, const int16_t *cb,
                                  const int16_t *cr, uint32_t width, uint32_t height,
                                  uint32_t color_type, uint32_t quality, uint32_t subsampling,
                                  uint32_t restart_interval, uint32_t optimize_huffman,
                                  uint8_t *out, size_t out_cap, size_t *out_len);

/* The same for coefficient arrays that live on the DEVICE (natural order, the layout
 * pixo_b200_jpeg_coefficients_dev writes): optimised-table statistics (K3) and the Huffman /
 * stuffing / restart stage (k_huff) run on the GPU, only the scan bytes come back; out receives
 * the complete JPEG.  This is what a frame tiled over several GPUs uses once its bands'
 * coefficients have been gathered on one of them. */
int pixo_b200_jpeg_entropy_encode_dev(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb,
                                      const int16_t *d_cr, uint32_t width, uint32_t height,
                                      uint32_t color_type, uint32_t quality, uint32_t subsampling,
                                      uint32_t restart_interval, uint32_t optimize_huffman,
                                      uint8_t *out, size_t out_cap, size_t *out_len);

/* ---- one frame tiled over several GPUs (BASELINE config C4) --------------
 * The reference's analogue is compute_all_coefficients_parallel (src/jpeg/mod.rs:1137-1215, rayon
 * over MCU rows) followed by the sequential encode_scan.  Here every GPU owns a contiguous band of
 * MCU rows and runs the WHOLE path on it; only finished scan bytes are gathered:
 *   1. transform: pixo_b200_jpeg_coefficients_dev on the band's pixel rows (a band starts on an
 *      MCU row, so it is an image of its own: no halo; the frame's bottom clamp falls in the last band);
 *   2. the DC predictors cross bands: band r starts from band r-1's last DC per component
 *      (pixo_b200_jpeg_band_last_dc; one tiny all-gather);
 *   3. [optimize_huffman] pixo_b200_jpeg_band_histogram_dev, statistics summed over the bands
 *      (all-reduce of 536 u64), identical tables on every rank;
 *   4. pixo_b200_jpeg_band_entropy_dev: the band's Huffman code as a raw bit string (no 0xFF
 *      stuffing, no padding: both depend on the bit offset, which is not known yet) plus its bit
 *      count and last 7 bits (second tiny all-gather -> every band's bit offset in the frame's stream
 *      and the bits it inherits in its first byte);
 *   5. pixo_b200_jpeg_band_splice_dev: shift to the bit offset, complete the shared first byte, stuff
 *      0xFF -> 0xFF00 (src/bits.rs:245-259), 1-pad the frame's last byte (:261-272).  A band owns the
 *      stream bytes whose LAST bit it wrote;
 *   6. rank 0 writes pixo_b200_jpeg_write_headers, the bands' bytes in band order, EOI.
 * No restart interval on this path (every interval would need its own offset exchange).
 * The `_dev` calls return after their host outputs are valid.  Host twins (no device needed) take
 * host arrays: they serve the CPU-only tests and a host that merely assembles. */
int pixo_b200_jpeg_band_last_dc(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb,
                                const int16_t *d_cr, size_t ny, size_t nc, int32_t last_dc[3]);
int pixo_b200_jpeg_band_histogram_dev(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb,
                                      const int16_t *d_cr, uint32_t width, uint32_t band_height,
                                      uint32_t color_type, uint32_t subsampling,
                                      const int32_t dc_seed[3], uint64_t *d_hist /* 536, device */);
/* hist (host, optional): the frame's summed statistics -> optimised tables (standard when NULL or
 * when they cannot be built, as the reference's unwrap_or_default does).  d_raw: 16-byte aligned,
 * raw_cap a multiple of 4; it must stay untouched until the band has been spliced.  A raw_cap of at
 * least the band's pixel bytes + 1 MiB lets a long band be coded as several independent segments
 * (shorter look-back chains); with less room the band is coded as one string. */
int pixo_b200_jpeg_band_entropy_dev(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb,
                                    const int16_t *d_cr, uint32_t width, uint32_t band_height,
                                    uint32_t color_type, uint32_t subsampling,
                                    const int32_t dc_seed[3], const uint64_t *hist, uint8_t *d_raw,
                                    size_t raw_cap, uint64_t *nbits, uint32_t *tail7);
/* start_bit: bits of the frame's stream before this band; tail_in: the last (start_bit % 8) of them. */
int pixo_b200_jpeg_band_splice_dev(pixo_b200_ctx *ctx, const uint8_t *d_raw, uint64_t nbits,
                                   uint64_t start_bit, uint32_t tail_in, uint32_t is_last_band,
                                   uint8_t *d_out, size_t out_cap, uint64_t *out_len);
/* Stream-ordered variants: everything that crosses a band boundary stays in DEVICE memory, so a
 * rank can queue transform -> (NCCL all-gather of predictors) -> coding -> (all-gather of bits) ->
 * splice without a single host synchronisation; only the final byte count is read back.
 *   d_dc_seed:   int32[3] in device memory (written by an earlier operation on the stream)
 *   d_bits_tail: uint64[2] out, {bit count, last 7 bits}
 *   d_offset:    uint64[3] in, {start_bit, the last 7 bits of the stream before this band, is_last_band}
 *   d_flags:     uint32, OR-ed: bit 0 a capacity was too small, bit 1 device fault (caller zeroes it)
 * raw_cap must be at least the band's pixel bytes + 1 MiB. */
int pixo_b200_jpeg_band_entropy_dev_async(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb,
                                          const int16_t *d_cr, uint32_t width, uint32_t band_height,
                                          uint32_t color_type, uint32_t subsampling,
                                          const int32_t *d_dc_seed, const uint64_t *hist, uint8_t *d_raw,
                                          size_t raw_cap, uint64_t *d_bits_tail, uint32_t *d_flags);
int pixo_b200_jpeg_band_splice_dev_async(pixo_b200_ctx *ctx, const uint8_t *d_raw, const uint64_t *d_offset,
                                         uint8_t *d_out, size_t out_cap, uint64_t *d_out_len,
                                         uint32_t *d_flags);
int pixo_b200_jpeg_band_entropy(const int16_t *y, const int16_t *cb, const int16_t *cr, uint32_t width,
                                uint32_t band_height, uint32_t color_type, uint32_t subsampling,
                                const int32_t dc_seed[3], const uint64_t *hist, uint8_t *raw,
                                size_t raw_cap, uint64_t *nbits, uint32_t *tail7);
int pixo_b200_jpeg_band_histogram(const int16_t *y, const int16_t *cb, const int16_t *cr, uint32_t width,
                                  uint32_t band_height, uint32_t color_type, uint32_t subsampling,
                                  const int32_t dc_seed[3], uint64_t hist[536]);
int pixo_b200_jpeg_band_splice(const uint8_t *raw, uint64_t nbits, uint64_t start_bit, uint32_t tail_in,
                               uint32_t is_last_band, uint8_t *out, size_t out_cap, size_t *out_len);
/* SOI .. SOS of the frame (src/jpeg/mod.rs:395-430,449-648); out_cap >= 1024.  Host-only. */
int pixo_b200_jpeg_write_headers(uint32_t width, uint32_t height, uint32_t color_type, uint32_t quality,
                                 uint32_t subsampling, uint32_t restart_interval, const uint64_t *hist,
                                 uint8_t *out, size_t out_cap, size_t *out_len);

/* ---- PNG -------------------------------------------------------------------------------- */

/* Replaces filter::apply_filters_with_row_bytes — src/png/filter.rs:64-206 (+ the rayon path
 * apply_filters_parallel :574-608), i.e. filter_{sub,up,average,paeth} (src/simd/mod.rs:159-236,
 * normative scalar src/simd/fallback.rs:100-159), score_filter (:93-98), adaptive_filter
 * (filter.rs:302-393), adaptive_filter_fast (:474-527), bigrams_filter (:410-471).
 * Semantics follow the default-feature build (`parallel` on): area <= 4096 forces Sub; for
 * height <= 32 AdaptiveFast is sticky on row 0's winner (:147-166).
 * out: height*(row_bytes+1).  adler32_out (optional): Adler-32 of `out`
 * (src/compress/adler32.rs:11-47), computed on the device in the same pass.
 * strategy may be OR-ed with PIXO_B200_PNG_OPTIMIZE_ALPHA: the pre-pass encode() runs before
 * filtering when PngOptions::optimize_alpha is set (maybe_optimize_alpha, src/png/mod.rs:633-671:
 * colour bytes of fully transparent pixels become 0; bytes_per_pixel 4 = Rgba, 2 = GrayAlpha,
 * other pixel sizes are left alone) is applied on the fly while the rows are read. */
#define PIXO_B200_PNG_OPTIMIZE_ALPHA 0x100u
int pixo_b200_png_filter(pixo_b200_ctx *ctx, const uint8_t *data, uint32_t width,
                         uint32_t height, size_t row_bytes, uint32_t bytes_per_pixel,
                         uint32_t strategy, uint8_t *out, uint32_t *adler32_out);

/* Device-pointer, batched variant (asynchronous).  Frame i: d_data + i*in_stride ->
 * d_out + i*out_stride; d_adler (optional): n_images u32. */
int pixo_b200_png_filter_dev(pixo_b200_ctx *ctx, const uint8_t *d_data, size_t in_stride,
                             uint32_t n_images, uint32_t width, uint32_t height,
                             size_t row_bytes, uint32_t bytes_per_pixel, uint32_t strategy,
                             uint8_t *d_out, size_t out_stride, uint32_t *d_adler);

/* One image's rows in bands (filters read the RAW previous row, so a band
 * only needs the one raw row above it - an overlapping read, not an exchange).  d_rows: band_rows
 * rows of the image starting at some row r0; d_row_above: the raw row r0-1 (NULL for r0 == 0 =
 * zeros, src/png/filter.rs:112-117); image_height: rows of the WHOLE image (the strategy pre-rules
 * of apply_filters_with_row_bytes look at the whole image).  d_out: band_rows*(row_bytes+1);
 * d_adler (optional): Adler-32 of this band's slice of the filtered stream, started from the
 * initial state; combine the bands' values in order (s1 = s1A + s1B - 1, s2 = s2A + s2B +
 * lenB*(s1A - 1) mod 65521 - pixo_b200_adler32_combine). */
int pixo_b200_png_filter_rows_dev(pixo_b200_ctx *ctx, const uint8_t *d_rows, const uint8_t *d_row_above,
                                  uint32_t width, uint32_t image_height, uint32_t band_rows,
                                  size_t row_bytes, uint32_t bytes_per_pixel, uint32_t strategy,
                                  uint8_t *d_out, uint32_t *d_adler);
/* Adler-32 of A ++ B from adler32(A), adler32(B) and len(B).  Host-only. */
uint32_t pixo_b200_adler32_combine(uint32_t adler_a, uint32_t adler_b, uint64_t len_b);

/* Replaces compress::adler32::adler32 — src/compress/adler32.rs:11-47 (dispatch
 * src/simd/mod.rs:72-90).  Host buffer in, checksum out. */
int pixo_b200_adler32(pixo_b200_ctx *ctx, const uint8_t *data, size_t len, uint32_t *out);
/* device buffer; result written to d_out (device u32), asynchronous */
int pixo_b200_adler32_dev(pixo_b200_ctx *ctx, const uint8_t *d_data, size_t len, uint32_t *d_out);

#ifdef __cplusplus
}
#endif
#endif /* PIXO_B200_H */
