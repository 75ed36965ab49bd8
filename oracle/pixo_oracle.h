/*
 * pixo_oracle.h — CPU restatement ("oracle") of leerob/pixo's encode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or the CPU baseline.  The product (libpixo_b200.so) never
 * links, loads or calls it.
 *
 * Every function cites the reference file:line (relative to the pixo tree @ 437bf63, v0.4.1)
 * whose arithmetic it restates.  Built with
 *   gcc -O2 -ffp-contract=off -fno-fast-math -msse2 -mfpmath=sse
 * so every float op is one rounded IEEE-754 binary32 operation, as in the Rust reference.
 *
 * Parity: PINNED (oracle/README.md).  The restatement is checked against (a) every known-answer
 * test the reference holds for this path (tests/test_oracle_kat.py) and (b) complete JPEG / PNG
 * files produced by the reference's own compiled artefact (pixo_bg.wasm) executed in the build
 * container by oracle/wasm_ref: 71 JPEG + 63 PNG fixtures under tests/golden, reproduced byte
 * for byte (tests/test_golden_reference.py).
 */
#ifndef PIXO_ORACLE_H
#define PIXO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ColorType repr(u8), src/color.rs:8-18 */
enum { PO_GRAY = 0, PO_GRAY_ALPHA = 1, PO_RGB = 2, PO_RGBA = 3 };
/* Subsampling, src/jpeg/mod.rs:96-102 */
enum { PO_S444 = 0, PO_S420 = 1 };
/* FilterStrategy declaration order, src/png/mod.rs:345-364 */
enum {
    PO_F_NONE = 0, PO_F_SUB = 1, PO_F_UP = 2, PO_F_AVERAGE = 3, PO_F_PAETH = 4,
    PO_F_MINSUM = 5, PO_F_ADAPTIVE = 6, PO_F_ADAPTIVE_FAST = 7, PO_F_BIGRAMS = 8
};

/* ---- colour (src/color.rs:60-77) ---- */
void po_rgb_to_ycbcr(uint8_t r, uint8_t g, uint8_t b, uint8_t out[3]);

/* ---- quantisation (src/jpeg/quantize.rs) ---- */
extern const uint8_t PO_ZIGZAG[64];
void po_quant_tables(int quality, uint8_t lum_zz[64], uint8_t chr_zz[64],
                     float lum_nat[64], float chr_nat[64]);
void po_quantize_block(const float dct[64], const float q[64], int16_t out[64]);
void po_zigzag_reorder(const int16_t in[64], int16_t out[64]);

/* ---- DCT (src/jpeg/dct.rs:591-700) ---- */
void po_aan_dct_1d(float d[8]);
void po_dct_2d(const float in[64], float out[64]);

/* ---- block extraction (src/jpeg/mod.rs:1565-1656) ---- */
void po_extract_block(const uint8_t *data, size_t w, size_t h, size_t bx, size_t by,
                      int color_type, float yb[64], float cb[64], float cr[64]);
void po_extract_mcu_420(const uint8_t *data, size_t w, size_t h, size_t mx, size_t my,
                        float yb[4][64], float cb[64], float cr[64]);

/* number of blocks per component for a frame (src/jpeg/mod.rs:1055-1125) */
void po_jpeg_block_counts(uint32_t w, uint32_t h, int color_type, int subsampling,
                          size_t *ny, size_t *nc);

/* compute_all_coefficients (src/jpeg/mod.rs:932-1125): natural order, MCU order.
 * y: ny*64, cb/cr: nc*64 (unused for Gray).  Rows of MCUs [mcu_row0, mcu_row1) only when
 * mcu_row1 != 0 (used by the threaded CPU baseline); pass 0,0 for the whole frame. */
void po_jpeg_coefficients(const uint8_t *data, uint32_t w, uint32_t h, int color_type,
                          int subsampling, const float lum_q[64], const float chr_q[64],
                          int16_t *y, int16_t *cb, int16_t *cr,
                          uint32_t mcu_row0, uint32_t mcu_row1);

/* count_block histograms over a whole frame in scan order (src/jpeg/mod.rs:684-860).
 * hist layout: dc_lum[12], dc_chrom[12], ac_lum[256], ac_chrom[256]  (536 u64). */
void po_jpeg_histograms(const int16_t *y, const int16_t *cb, const int16_t *cr,
                        uint32_t w, uint32_t h, int color_type, int subsampling,
                        uint32_t restart_interval, uint64_t hist[536]);

/* symbol pre-scan of one block (src/jpeg/huffman.rs:431-478, src/jpeg/mod.rs:826-860):
 * emits (rs, amplitude bits, nbits) triples; first triple is the DC (cat, amp, cat).
 * Returns the number of triples; *dc_out receives the block's DC for chaining. */
int po_block_symbols(const int16_t nat[64], int16_t prev_dc, uint8_t rs[65], uint16_t amp[65],
                     uint8_t nbits[65], int16_t *dc_out);

/* Huffman table set (src/jpeg/huffman.rs:72-262) */
typedef struct {
    uint8_t bits[4][16];   /* dc_lum, dc_chrom, ac_lum, ac_chrom */
    uint8_t vals[4][256];
    int nvals[4];
    uint16_t code[4][256];
    uint8_t len[4][256];
} po_huff_tables;
void po_huff_standard(po_huff_tables *t);
/* optimized_from_counts (huffman.rs:167-205); returns 0 if it yields None (caller falls
 * back to the standard tables, jpeg/mod.rs:379-390), 1 otherwise. */
int po_huff_optimized(const uint64_t hist[536], int has_chroma, po_huff_tables *t);

/* Full encoder: pixo::jpeg::encode_into (src/jpeg/mod.rs:328-447), baseline only.
 * restart_interval 0 = None.  Returns bytes written, or a negative error:
 *  -1 invalid quality, -2 invalid dims, -3 too large, -4 unsupported colour, -5 bad length,
 *  -6 output capacity too small, -7 unsupported option (progressive/trellis). */
long po_jpeg_encode(const uint8_t *data, size_t data_len, uint32_t w, uint32_t h,
                    int color_type, int quality, int subsampling, uint32_t restart_interval,
                    int optimize_huffman, uint8_t *out, size_t cap);

/* Entropy-code precomputed coefficient arrays into a full baseline JPEG (same layout as
 * po_jpeg_encode).  Used to check the product's host entropy coder in isolation. */
long po_jpeg_encode_from_coefficients(const int16_t *y, const int16_t *cb, const int16_t *cr,
                                      uint32_t w, uint32_t h, int color_type, int quality,
                                      int subsampling, uint32_t restart_interval,
                                      int optimize_huffman, uint8_t *out, size_t cap);

/* ---- PNG (src/simd/fallback.rs:93-159, src/png/filter.rs) ---- */
void po_filter_sub(const uint8_t *row, size_t n, size_t bpp, uint8_t *out);
void po_filter_up(const uint8_t *row, const uint8_t *prev, size_t n, uint8_t *out);
void po_filter_average(const uint8_t *row, const uint8_t *prev, size_t n, size_t bpp, uint8_t *out);
void po_filter_paeth(const uint8_t *row, const uint8_t *prev, size_t n, size_t bpp, uint8_t *out);
uint8_t po_paeth_predictor(uint8_t a, uint8_t b, uint8_t c);
uint64_t po_score_filter(const uint8_t *f, size_t n);
size_t po_score_bigrams(const uint8_t *f, size_t n);

/* apply_filters_with_row_bytes (src/png/filter.rs:64-206, 574-608).
 * parallel_feature: 1 = default cargo features (rayon path for height > 32), 0 = built
 * without `parallel` (always the sequential loop, where AdaptiveFast is sticky).
 * out: height*(row_bytes+1).  Rows [row0,row1) only when row1 != 0 (threaded baseline; only
 * valid for row-independent strategies). */
void po_apply_filters(const uint8_t *data, uint32_t width, uint32_t height, size_t row_bytes,
                      size_t bpp, int strategy, int parallel_feature, uint8_t *out,
                      uint32_t row0, uint32_t row1);

/* maybe_optimize_alpha (src/png/mod.rs:633-671), in place: color_type 3 (Rgba) / 1 (GrayAlpha)
 * zero the colour bytes of pixels whose alpha is 0; other colour types are left alone. */
void po_optimize_alpha(uint8_t *data, size_t n_bytes, int color_type);

/* adler32 (src/compress/adler32.rs:26-47 == src/simd/fallback.rs:8-25) */
uint32_t po_adler32(const uint8_t *data, size_t n);
/* crc32 (src/simd/fallback.rs:27-58), adjacent known-answer only */
uint32_t po_crc32(const uint8_t *data, size_t n);

/* ---- synthetic generators restated from tests/support/synthetic.rs:74-85,183-197 and
 * benches/comparison.rs:32-59 (inputs only; used by tests and bench) ---- */
void po_gen_gradient_rgb(uint32_t w, uint32_t h, uint8_t *out);
void po_gen_noise(uint32_t w, uint32_t h, uint32_t channels, uint32_t seed, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
