/*
 * pixo_oracle.c — CPU restatement of pixo's JPEG / PNG encode hot path (see pixo_oracle.h).
 * TEST INFRASTRUCTURE ONLY — never linked into, loaded by, or shipped with the product.
 *
 * Compile: gcc -O2 -ffp-contract=off -fno-fast-math -msse2 -mfpmath=sse -fPIC -shared
 * All citations are file:line in the pixo tree @ 437bf63 (v0.4.1).
 */
#include "pixo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Colour — src/color.rs:60-77.  i32 math, arithmetic >>8, +128 after the shift, clamp.
 * ---------------------------------------------------------------------------------------- */
static inline int clamp_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

void po_rgb_to_ycbcr(uint8_t r8, uint8_t g8, uint8_t b8, uint8_t out[3])
{
    int32_t r = r8, g = g8, b = b8;
    int32_t y = (77 * r + 150 * g + 29 * b + 128) >> 8;
    int32_t cb = ((-43 * r - 85 * g + 128 * b + 128) >> 8) + 128;
    int32_t cr = ((128 * r - 107 * g - 21 * b + 128) >> 8) + 128;
    out[0] = (uint8_t)clamp_u8(y);
    out[1] = (uint8_t)clamp_u8(cb);
    out[2] = (uint8_t)clamp_u8(cr);
}

/* ------------------------------------------------------------------------------------------
 * Quantisation — src/jpeg/quantize.rs:4-113
 * ---------------------------------------------------------------------------------------- */
static const uint8_t STD_LUM[64] = {
    16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
    14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
    18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99 };
static const uint8_t STD_CHR[64] = {
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99 };

const uint8_t PO_ZIGZAG[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };

/* QuantizationTables::with_quality, quantize.rs:42-89 */
void po_quant_tables(int quality, uint8_t lum_zz[64], uint8_t chr_zz[64],
                     float lum_nat[64], float chr_nat[64])
{
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    uint32_t scale = quality < 50 ? 5000u / (uint32_t)quality : 200u - 2u * (uint32_t)quality;
    for (int i = 0; i < 64; i++) {
        uint32_t l = (STD_LUM[PO_ZIGZAG[i]] * scale + 50) / 100;
        uint32_t c = (STD_CHR[PO_ZIGZAG[i]] * scale + 50) / 100;
        l = l < 1 ? 1 : (l > 255 ? 255 : l);
        c = c < 1 ? 1 : (c > 255 ? 255 : c);
        if (lum_zz) lum_zz[i] = (uint8_t)l;
        if (chr_zz) chr_zz[i] = (uint8_t)c;
    }
    for (int i = 0; i < 64; i++) {
        uint32_t l = (STD_LUM[i] * scale + 50) / 100;
        uint32_t c = (STD_CHR[i] * scale + 50) / 100;
        l = l < 1 ? 1 : (l > 255 ? 255 : l);
        c = c < 1 ? 1 : (c > 255 ? 255 : c);
        if (lum_nat) lum_nat[i] = (float)l;
        if (chr_nat) chr_nat[i] = (float)c;
    }
}

/* Rust `f32 as i16`: saturating, NaN -> 0 */
static inline int16_t sat_i16(float v)
{
    if (v != v) return 0;
    if (v >= 32767.0f) return 32767;
    if (v <= -32768.0f) return -32768;
    return (int16_t)v;
}

/* quantize_block, quantize.rs:99-105: (dct/q).round() as i16 — IEEE divide, half away */
void po_quantize_block(const float dct[64], const float q[64], int16_t out[64])
{
    for (int i = 0; i < 64; i++) {
        volatile float quo = dct[i] / q[i];
        out[i] = sat_i16(roundf(quo));
    }
}

/* zigzag_reorder, quantize.rs:107-113 */
void po_zigzag_reorder(const int16_t in[64], int16_t out[64])
{
    for (int i = 0; i < 64; i++) out[i] = in[PO_ZIGZAG[i]];
}

/* ------------------------------------------------------------------------------------------
 * Forward DCT — src/jpeg/dct.rs:591-700 (f32 AAN, the one the encoder calls)
 * ---------------------------------------------------------------------------------------- */
static const float A1 = 0.70710678118654752440f; /* FRAC_1_SQRT_2, dct.rs:591 */
static const float A2 = 0.5411961f;              /* dct.rs:592 */
static const float A3 = 0.70710678118654752440f; /* dct.rs:593 */
static const float A4 = 1.3065629f;              /* dct.rs:594 */
static const float A5 = 0.38268343f;             /* dct.rs:595 */
static const float S[8] = { 0.3535534f, 0.2548978f, 0.2705981f, 0.3006724f,
                            0.3535534f, 0.4499881f, 0.6532815f, 1.2814578f }; /* :599-608 */

/* aan_dct_1d, dct.rs:648-700.  One rounded op per line, same order. */
void po_aan_dct_1d(float d[8])
{
    float tmp0 = d[0] + d[7];
    float tmp7 = d[0] - d[7];
    float tmp1 = d[1] + d[6];
    float tmp6 = d[1] - d[6];
    float tmp2 = d[2] + d[5];
    float tmp5 = d[2] - d[5];
    float tmp3 = d[3] + d[4];
    float tmp4 = d[3] - d[4];

    float tmp10 = tmp0 + tmp3;
    float tmp13 = tmp0 - tmp3;
    float tmp11 = tmp1 + tmp2;
    float tmp12 = tmp1 - tmp2;

    d[0] = tmp10 + tmp11;
    d[4] = tmp10 - tmp11;

    float z1 = (tmp12 + tmp13) * A1;
    d[2] = tmp13 + z1;
    d[6] = tmp13 - z1;

    tmp10 = tmp4 + tmp5;
    tmp11 = tmp5 + tmp6;
    tmp12 = tmp6 + tmp7;

    float z5 = (tmp10 - tmp12) * A5;
    float z2 = tmp10 * A2 + z5; /* mul then add: two roundings (-ffp-contract=off) */
    float z4 = tmp12 * A4 + z5;
    float z3 = tmp11 * A3;

    float z11 = tmp7 + z3;
    float z13 = tmp7 - z3;

    d[5] = z13 + z2;
    d[3] = z13 - z2;
    d[1] = z11 + z4;
    d[7] = z11 - z4;

    for (int i = 0; i < 8; i++) d[i] *= S[i];
}

/* dct_2d, dct.rs:614-646: rows first, then columns */
void po_dct_2d(const float in[64], float out[64])
{
    float tmp[64];
    for (int r = 0; r < 8; r++) {
        float row[8];
        memcpy(row, in + r * 8, sizeof row);
        po_aan_dct_1d(row);
        memcpy(tmp + r * 8, row, sizeof row);
    }
    for (int c = 0; c < 8; c++) {
        float col[8];
        for (int r = 0; r < 8; r++) col[r] = tmp[r * 8 + c];
        po_aan_dct_1d(col);
        for (int r = 0; r < 8; r++) out[r * 8 + c] = col[r];
    }
}

/* ------------------------------------------------------------------------------------------
 * Block extraction — src/jpeg/mod.rs:1565-1656
 * ---------------------------------------------------------------------------------------- */
static inline size_t min_sz(size_t a, size_t b) { return a < b ? a : b; }

/* extract_block, mod.rs:1565-1606 */
void po_extract_block(const uint8_t *data, size_t w, size_t h, size_t bx, size_t by,
                      int color_type, float yb[64], float cb[64], float cr[64])
{
    for (size_t dy = 0; dy < 8; dy++) {
        for (size_t dx = 0; dx < 8; dx++) {
            size_t x = min_sz(bx + dx, w - 1);
            size_t y = min_sz(by + dy, h - 1);
            size_t idx = dy * 8 + dx;
            if (color_type == PO_GRAY) {
                yb[idx] = (float)data[y * w + x] - 128.0f;
                cb[idx] = 0.0f;
                cr[idx] = 0.0f;
            } else {
                size_t p = (y * w + x) * 3;
                uint8_t c[3];
                po_rgb_to_ycbcr(data[p], data[p + 1], data[p + 2], c);
                yb[idx] = (float)c[0] - 128.0f;
                cb[idx] = (float)c[1] - 128.0f;
                cr[idx] = (float)c[2] - 128.0f;
            }
        }
    }
}

/* extract_mcu_420, mod.rs:1608-1656: chroma accumulated as f32, then *0.25 - 128.0 */
void po_extract_mcu_420(const uint8_t *data, size_t w, size_t h, size_t mx, size_t my,
                        float yb[4][64], float cb[64], float cr[64])
{
    for (int i = 0; i < 64; i++) { cb[i] = 0.0f; cr[i] = 0.0f; }
    for (size_t by = 0; by < 2; by++) {
        for (size_t bx = 0; bx < 2; bx++) {
            size_t bi = by * 2 + bx;
            for (size_t dy = 0; dy < 8; dy++) {
                for (size_t dx = 0; dx < 8; dx++) {
                    size_t x = min_sz(mx + bx * 8 + dx, w - 1);
                    size_t y = min_sz(my + by * 8 + dy, h - 1);
                    size_t p = (y * w + x) * 3;
                    uint8_t c[3];
                    po_rgb_to_ycbcr(data[p], data[p + 1], data[p + 2], c);
                    yb[bi][dy * 8 + dx] = (float)c[0] - 128.0f;
                    size_t gx = bx * 8 + dx, gy = by * 8 + dy;
                    size_t ci = (gy / 2) * 8 + gx / 2;
                    cb[ci] += (float)c[1];
                    cr[ci] += (float)c[2];
                }
            }
        }
    }
    for (int i = 0; i < 64; i++) {
        cb[i] = cb[i] * 0.25f - 128.0f;
        cr[i] = cr[i] * 0.25f - 128.0f;
    }
}

void po_jpeg_block_counts(uint32_t w, uint32_t h, int color_type, int subsampling,
                          size_t *ny, size_t *nc)
{
    if (color_type == PO_GRAY || subsampling == PO_S444) {
        size_t n = (((size_t)w + 7) / 8) * (((size_t)h + 7) / 8);
        *ny = n;
        *nc = color_type == PO_GRAY ? 0 : n;
    } else {
        size_t m = (((size_t)w + 15) / 16) * (((size_t)h + 15) / 16);
        *ny = 4 * m;
        *nc = m;
    }
}

/* compute_all_coefficients_sequential, mod.rs:1046-1125 (the parallel variant, :1128-1230,
 * yields the same arrays in the same order).  quantize_dct with use_trellis = false. */
void po_jpeg_coefficients(const uint8_t *data, uint32_t w32, uint32_t h32, int color_type,
                          int subsampling, const float lum_q[64], const float chr_q[64],
                          int16_t *y, int16_t *cb, int16_t *cr,
                          uint32_t mcu_row0, uint32_t mcu_row1)
{
    size_t w = w32, h = h32;
    float yb[4][64], cbb[64], crb[64], d[64];
    if (color_type == PO_GRAY || subsampling == PO_S444) {
        size_t bw = (w + 7) / 8, bh = (h + 7) / 8;
        size_t r0 = 0, r1 = bh;
        if (mcu_row1) { r0 = mcu_row0; r1 = mcu_row1 < bh ? mcu_row1 : bh; }
        for (size_t by = r0; by < r1; by++) {
            for (size_t bx = 0; bx < bw; bx++) {
                size_t bi = by * bw + bx;
                po_extract_block(data, w, h, bx * 8, by * 8, color_type, yb[0], cbb, crb);
                po_dct_2d(yb[0], d);
                po_quantize_block(d, lum_q, y + bi * 64);
                if (color_type != PO_GRAY) {
                    po_dct_2d(cbb, d);
                    po_quantize_block(d, chr_q, cb + bi * 64);
                    po_dct_2d(crb, d);
                    po_quantize_block(d, chr_q, cr + bi * 64);
                }
            }
        }
    } else {
        size_t mw = (w + 15) / 16, mh = (h + 15) / 16;
        size_t r0 = 0, r1 = mh;
        if (mcu_row1) { r0 = mcu_row0; r1 = mcu_row1 < mh ? mcu_row1 : mh; }
        for (size_t my = r0; my < r1; my++) {
            for (size_t mx = 0; mx < mw; mx++) {
                size_t mi = my * mw + mx;
                po_extract_mcu_420(data, w, h, mx * 16, my * 16, yb, cbb, crb);
                for (int k = 0; k < 4; k++) {
                    po_dct_2d(yb[k], d);
                    po_quantize_block(d, lum_q, y + (mi * 4 + k) * 64);
                }
                po_dct_2d(cbb, d);
                po_quantize_block(d, chr_q, cb + mi * 64);
                po_dct_2d(crb, d);
                po_quantize_block(d, chr_q, cr + mi * 64);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Symbol pre-scan — src/jpeg/huffman.rs:394-481, src/jpeg/mod.rs:826-870
 * ---------------------------------------------------------------------------------------- */
/* category / category_i16, huffman.rs:394-401, mod.rs:862-870 */
static inline uint8_t category(int16_t v)
{
    uint16_t a = v < 0 ? (uint16_t)(-(int32_t)v) : (uint16_t)v;
    uint8_t c = 0;
    while (a) { c++; a >>= 1; }
    return c;
}

/* encode_value, huffman.rs:404-418 */
static inline void encode_value(int16_t v, uint16_t *bits, uint8_t *n)
{
    uint8_t cat = category(v);
    if (cat == 0) { *bits = 0; *n = 0; return; }
    uint16_t b = v < 0 ? (uint16_t)(int16_t)(v - 1) : (uint16_t)v;
    *bits = (uint16_t)(b & ((1u << cat) - 1u));
    *n = cat;
}

int po_block_symbols(const int16_t nat[64], int16_t prev_dc, uint8_t rs[65], uint16_t amp[65],
                     uint8_t nbits[65], int16_t *dc_out)
{
    int16_t zz[64];
    po_zigzag_reorder(nat, zz);
    int n = 0;
    int16_t dc = zz[0];
    int16_t diff = (int16_t)(dc - prev_dc);
    uint8_t cat = category(diff);
    rs[n] = cat;
    encode_value(diff, &amp[n], &nbits[n]);
    n++;
    int run = 0;
    for (int i = 1; i < 64; i++) {
        int16_t ac = zz[i];
        if (ac == 0) { run++; continue; }
        while (run >= 16) { rs[n] = 0xF0; amp[n] = 0; nbits[n] = 0; n++; run -= 16; }
        rs[n] = (uint8_t)((run << 4) | category(ac));
        encode_value(ac, &amp[n], &nbits[n]);
        n++;
        run = 0;
    }
    if (run > 0) { rs[n] = 0x00; amp[n] = 0; nbits[n] = 0; n++; }
    if (dc_out) *dc_out = dc;
    return n;
}

/* count_block, mod.rs:826-860 */
static int16_t count_block(const int16_t nat[64], int16_t prev_dc, uint64_t *dc_counts,
                           uint64_t *ac_counts)
{
    uint8_t rs[65], nb[65];
    uint16_t amp[65];
    int16_t dc;
    int n = po_block_symbols(nat, prev_dc, rs, amp, nb, &dc);
    dc_counts[rs[0]]++;
    for (int i = 1; i < n; i++) ac_counts[rs[i]]++;
    return dc;
}

/* Frame iteration helper: calls fn for every block in scan order with its component.
 * Scan order: mod.rs:1449-1556 (Gray: Y; 444: Y,Cb,Cr per 8x8; 420: Y*4,Cb,Cr per 16x16). */
typedef struct {
    const int16_t *y, *cb, *cr;
    size_t total_mcus;
    int blocks_y_per_mcu;
    int has_chroma;
} frame_iter;

static void frame_iter_init(frame_iter *it, const int16_t *y, const int16_t *cb,
                            const int16_t *cr, uint32_t w, uint32_t h, int color_type,
                            int subsampling)
{
    size_t ny, nc;
    po_jpeg_block_counts(w, h, color_type, subsampling, &ny, &nc);
    it->y = y; it->cb = cb; it->cr = cr;
    it->has_chroma = color_type != PO_GRAY;
    if (color_type == PO_GRAY || subsampling == PO_S444) {
        it->total_mcus = ny; it->blocks_y_per_mcu = 1;
    } else {
        it->total_mcus = nc; it->blocks_y_per_mcu = 4;
    }
}

/* build_optimized_huffman_tables' counting loops, mod.rs:684-824 */
void po_jpeg_histograms(const int16_t *y, const int16_t *cb, const int16_t *cr,
                        uint32_t w, uint32_t h, int color_type, int subsampling,
                        uint32_t restart_interval, uint64_t hist[536])
{
    memset(hist, 0, 536 * sizeof(uint64_t));
    uint64_t *dc_lum = hist, *dc_chr = hist + 12, *ac_lum = hist + 24, *ac_chr = hist + 280;
    frame_iter it;
    frame_iter_init(&it, y, cb, cr, w, h, color_type, subsampling);
    int16_t py = 0, pcb = 0, pcr = 0;
    uint32_t mcu_count = 0;
    for (size_t m = 0; m < it.total_mcus; m++) {
        for (int k = 0; k < it.blocks_y_per_mcu; k++)
            py = count_block(y + (m * it.blocks_y_per_mcu + k) * 64, py, dc_lum, ac_lum);
        if (it.has_chroma) {
            pcb = count_block(cb + m * 64, pcb, dc_chr, ac_chr);
            pcr = count_block(cr + m * 64, pcr, dc_chr, ac_chr);
        }
        mcu_count++;
        if (restart_interval > 0 && mcu_count % restart_interval == 0 &&
            mcu_count < (uint32_t)it.total_mcus) {
            py = 0; pcb = 0; pcr = 0;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Huffman tables — src/jpeg/huffman.rs:17-391
 * ---------------------------------------------------------------------------------------- */
static const uint8_t DC_LUM_BITS[16] = { 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0 };
static const uint8_t DC_CHR_BITS[16] = { 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0 };
static const uint8_t DC_VALS[12] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 };
static const uint8_t AC_LUM_BITS[16] = { 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125 };
static const uint8_t AC_LUM_VALS[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61,
    0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52,
    0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25,
    0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45,
    0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64,
    0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
    0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3,
    0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8,
    0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };
static const uint8_t AC_CHR_BITS[16] = { 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119 };
static const uint8_t AC_CHR_VALS[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61,
    0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33,
    0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18,
    0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44,
    0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63,
    0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97,
    0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca,
    0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7,
    0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };

/* build_codes / build_codes_256 / build_code_table, huffman.rs:213-286.
 * table_len: 12 for DC, 256 for AC.  strict: build_code_table semantics (None on overflow). */
static int build_codes(const uint8_t bits[16], const uint8_t *vals, int nvals, int table_len,
                       int strict, uint16_t code_out[256], uint8_t len_out[256])
{
    memset(code_out, 0, 256 * sizeof(uint16_t));
    memset(len_out, 0, 256);
    uint16_t code = 0;
    int vi = 0;
    for (int length = 0; length < 16; length++) {
        for (int k = 0; k < bits[length]; k++) {
            if (vi >= nvals) {
                if (strict) return 0;
            } else {
                int sym = vals[vi];
                if (sym >= table_len) {
                    if (strict) return 0;
                } else {
                    code_out[sym] = code;
                    len_out[sym] = (uint8_t)(length + 1);
                }
                vi++;
            }
            code++;
        }
        code <<= 1;
    }
    return 1;
}

static void set_spec(po_huff_tables *t, int k, const uint8_t bits[16], const uint8_t *vals,
                     int nvals)
{
    memcpy(t->bits[k], bits, 16);
    memset(t->vals[k], 0, 256);
    memcpy(t->vals[k], vals, (size_t)nvals);
    t->nvals[k] = nvals;
}

/* HuffmanTables::new, huffman.rs:100-120 */
void po_huff_standard(po_huff_tables *t)
{
    set_spec(t, 0, DC_LUM_BITS, DC_VALS, 12);
    set_spec(t, 1, DC_CHR_BITS, DC_VALS, 12);
    set_spec(t, 2, AC_LUM_BITS, AC_LUM_VALS, 162);
    set_spec(t, 3, AC_CHR_BITS, AC_CHR_VALS, 162);
    for (int k = 0; k < 4; k++)
        build_codes(t->bits[k], t->vals[k], t->nvals[k], k < 2 ? 12 : 256, 0, t->code[k],
                    t->len[k]);
}

/* build_code_lengths, huffman.rs:313-391.  BinaryHeap<Reverse<(freq, idx)>> pops the smallest
 * (freq, node index) pair; leaf length = depth + 1 (sic), None if any length > 16. */
static int build_code_lengths(const uint64_t *counts, int n, uint8_t *lengths)
{
    enum { MAXN = 512 };
    uint64_t freq[MAXN];
    int left[MAXN], right[MAXN], symbol[MAXN], alive[MAXN];
    int nn = 0;
    memset(lengths, 0, (size_t)n);
    for (int s = 0; s < n; s++) {
        if (counts[s] == 0) continue;
        freq[nn] = counts[s]; left[nn] = right[nn] = -1; symbol[nn] = s; alive[nn] = 1; nn++;
    }
    if (nn == 0) return 0;
    if (nn == 1) { lengths[symbol[0]] = 1; return 1; }
    int live = nn;
    while (live > 1) {
        int i1 = -1, i2 = -1;
        for (int i = 0; i < nn; i++) {
            if (!alive[i]) continue;
            if (i1 < 0 || freq[i] < freq[i1]) i1 = i; /* ties: lowest index first */
        }
        alive[i1] = 0;
        for (int i = 0; i < nn; i++) {
            if (!alive[i]) continue;
            if (i2 < 0 || freq[i] < freq[i2]) i2 = i;
        }
        alive[i2] = 0;
        freq[nn] = freq[i1] + freq[i2]; left[nn] = i1; right[nn] = i2; symbol[nn] = -1;
        alive[nn] = 1; nn++;
        live--;
    }
    int root = nn - 1;
    int stack_n[MAXN], stack_d[MAXN], sp = 0;
    stack_n[sp] = root; stack_d[sp] = 0; sp++;
    while (sp) {
        sp--;
        int idx = stack_n[sp], depth = stack_d[sp];
        if (symbol[idx] >= 0) {
            int len = depth + 1;
            if (len > 16) return 0;
            lengths[symbol[idx]] = (uint8_t)len;
        } else {
            stack_n[sp] = left[idx]; stack_d[sp] = depth + 1; sp++;
            stack_n[sp] = right[idx]; stack_d[sp] = depth + 1; sp++;
        }
    }
    return 1;
}

/* build_bits_vals, huffman.rs:288-311: vals by (length, symbol) */
static int build_bits_vals(const uint64_t *counts, int n, uint8_t bits[16], uint8_t *vals,
                           int *nvals)
{
    uint8_t lengths[256];
    if (!build_code_lengths(counts, n, lengths)) return 0;
    memset(bits, 0, 16);
    for (int i = 0; i < n; i++) {
        if (!lengths[i]) continue;
        if (lengths[i] > 16) return 0;
        bits[lengths[i] - 1]++;
    }
    int k = 0;
    for (int len = 1; len <= 16; len++)
        for (int s = 0; s < n; s++)
            if (lengths[s] == len) vals[k++] = (uint8_t)s;
    *nvals = k;
    return 1;
}

/* optimized_from_counts + from_specs, huffman.rs:144-205 */
int po_huff_optimized(const uint64_t hist[536], int has_chroma, po_huff_tables *t)
{
    uint8_t bits[16], vals[256];
    int nv;
    if (!build_bits_vals(hist, 12, bits, vals, &nv)) return 0;          /* dc_lum ? */
    set_spec(t, 0, bits, vals, nv);
    if (!build_bits_vals(hist + 24, 256, bits, vals, &nv)) return 0;    /* ac_lum ? */
    set_spec(t, 2, bits, vals, nv);
    if (has_chroma && build_bits_vals(hist + 12, 12, bits, vals, &nv)) set_spec(t, 1, bits, vals, nv);
    else set_spec(t, 1, DC_CHR_BITS, DC_VALS, 12);
    if (has_chroma && build_bits_vals(hist + 280, 256, bits, vals, &nv)) set_spec(t, 3, bits, vals, nv);
    else set_spec(t, 3, AC_CHR_BITS, AC_CHR_VALS, 162);
    for (int k = 0; k < 4; k++)
        if (!build_codes(t->bits[k], t->vals[k], t->nvals[k], k < 2 ? 12 : 256, 1, t->code[k],
                         t->len[k]))
            return 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * BitWriterMsb — src/bits.rs:195-290
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t *buf;
    size_t len, cap;
    uint8_t cur;
    uint8_t pos; /* counts from 8 down to 0 */
    int overflow;
} bitw;

static inline void bw_push(bitw *w, uint8_t b)
{
    if (w->len < w->cap) w->buf[w->len++] = b;
    else w->overflow = 1;
}

/* write_bits, bits.rs:216-242 */
static void bw_write(bitw *w, uint32_t value, uint8_t nbits)
{
    uint8_t remaining = nbits;
    while (remaining > 0) {
        uint8_t space = w->pos;
        uint8_t to_write = remaining < space ? remaining : space;
        uint8_t shift = (uint8_t)(remaining - to_write);
        uint32_t mask = (1u << to_write) - 1u;
        uint8_t bits = (uint8_t)((value >> shift) & mask);
        w->pos = (uint8_t)(w->pos - to_write);
        w->cur |= (uint8_t)(bits << w->pos);
        remaining = (uint8_t)(remaining - to_write);
        if (w->pos == 0) { /* flush_byte_with_stuffing, bits.rs:245-254 */
            bw_push(w, w->cur);
            if (w->cur == 0xFF) bw_push(w, 0x00);
            w->cur = 0;
            w->pos = 8;
        }
    }
}

/* flush, bits.rs:261-272: pad with 1s, stuff if it became 0xFF */
static void bw_flush(bitw *w)
{
    if (w->pos < 8) {
        w->cur |= (uint8_t)((1u << w->pos) - 1u);
        bw_push(w, w->cur);
        if (w->cur == 0xFF) bw_push(w, 0x00);
        w->cur = 0;
        w->pos = 8;
    }
}

/* encode_block, huffman.rs:423-481 */
static int16_t encode_block(bitw *w, const int16_t nat[64], int16_t prev_dc, int is_lum,
                            const po_huff_tables *t)
{
    uint8_t rs[65], nb[65];
    uint16_t amp[65];
    int16_t dc;
    int n = po_block_symbols(nat, prev_dc, rs, amp, nb, &dc);
    int dct = is_lum ? 0 : 1, act = is_lum ? 2 : 3;
    bw_write(w, t->code[dct][rs[0]], t->len[dct][rs[0]]);
    if (rs[0] > 0) bw_write(w, amp[0], nb[0]);
    for (int i = 1; i < n; i++) {
        bw_write(w, t->code[act][rs[i]], t->len[act][rs[i]]);
        if (rs[i] != 0xF0 && rs[i] != 0x00) bw_write(w, amp[i], nb[i]);
    }
    return dc;
}

/* ------------------------------------------------------------------------------------------
 * Headers — src/jpeg/mod.rs:449-648
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint8_t *p; size_t len, cap; int overflow; } outbuf;
static void ob_push(outbuf *o, uint8_t b)
{
    if (o->len < o->cap) o->p[o->len++] = b; else o->overflow = 1;
}
static void ob_u16(outbuf *o, unsigned v) { ob_push(o, (uint8_t)(v >> 8)); ob_push(o, (uint8_t)v); }
static void ob_bytes(outbuf *o, const uint8_t *b, size_t n) { for (size_t i = 0; i < n; i++) ob_push(o, b[i]); }

static void write_headers(outbuf *o, uint32_t w, uint32_t h, int color_type, int subsampling,
                          const uint8_t lum_zz[64], const uint8_t chr_zz[64],
                          const po_huff_tables *t, uint32_t restart_interval)
{
    ob_u16(o, 0xFFD8);                                   /* write_soi  :449 */
    ob_u16(o, 0xFFE0); ob_u16(o, 16);                    /* write_app0 :457-482 */
    ob_bytes(o, (const uint8_t *)"JFIF\0", 5);
    ob_push(o, 1); ob_push(o, 1); ob_push(o, 0);
    ob_u16(o, 1); ob_u16(o, 1); ob_push(o, 0); ob_push(o, 0);
    ob_u16(o, 0xFFDB); ob_u16(o, 67); ob_push(o, 0); ob_bytes(o, lum_zz, 64); /* write_dqt :484-496 */
    ob_u16(o, 0xFFDB); ob_u16(o, 67); ob_push(o, 1); ob_bytes(o, chr_zz, 64);
    int ncomp = color_type == PO_GRAY ? 1 : 3;           /* write_sof_marker :518-573 */
    ob_u16(o, 0xFFC0); ob_u16(o, (unsigned)(8 + 3 * ncomp)); ob_push(o, 8);
    ob_u16(o, h & 0xFFFF); ob_u16(o, w & 0xFFFF); ob_push(o, (uint8_t)ncomp);
    if (ncomp == 1) {
        ob_push(o, 1); ob_push(o, 0x11); ob_push(o, 0);
    } else {
        ob_push(o, 1); ob_push(o, subsampling == PO_S420 ? 0x22 : 0x11); ob_push(o, 0);
        ob_push(o, 2); ob_push(o, 0x11); ob_push(o, 1);
        ob_push(o, 3); ob_push(o, 0x11); ob_push(o, 1);
    }
    static const uint8_t ids[4] = { 0x00, 0x01, 0x10, 0x11 };   /* write_dht :575-610 */
    for (int k = 0; k < 4; k++) {
        ob_u16(o, 0xFFC4); ob_u16(o, (unsigned)(2 + 1 + 16 + t->nvals[k]));
        ob_push(o, ids[k]); ob_bytes(o, t->bits[k], 16); ob_bytes(o, t->vals[k], (size_t)t->nvals[k]);
    }
    if (restart_interval) {                               /* write_dri :592-596 */
        ob_u16(o, 0xFFDD); ob_u16(o, 4); ob_u16(o, restart_interval & 0xFFFF);
    }
    ob_u16(o, 0xFFDA); ob_u16(o, (unsigned)(6 + 2 * ncomp)); ob_push(o, (uint8_t)ncomp); /* write_sos :612-648 */
    if (ncomp == 1) {
        ob_push(o, 1); ob_push(o, 0x00);
    } else {
        ob_push(o, 1); ob_push(o, 0x00); ob_push(o, 2); ob_push(o, 0x11); ob_push(o, 3); ob_push(o, 0x11);
    }
    ob_push(o, 0); ob_push(o, 63); ob_push(o, 0);
}

/* encode_scan, mod.rs:1408-1563, consuming precomputed coefficient arrays */
static void encode_scan(bitw *bw, const frame_iter *it, const po_huff_tables *t,
                        uint32_t restart_interval)
{
    int16_t py = 0, pcb = 0, pcr = 0;
    uint8_t rst_idx = 0;
    uint32_t mcu_count = 0;
    for (size_t m = 0; m < it->total_mcus; m++) {
        for (int k = 0; k < it->blocks_y_per_mcu; k++)
            py = encode_block(bw, it->y + (m * it->blocks_y_per_mcu + k) * 64, py, 1, t);
        if (it->has_chroma) {
            pcb = encode_block(bw, it->cb + m * 64, pcb, 0, t);
            pcr = encode_block(bw, it->cr + m * 64, pcr, 0, t);
        }
        mcu_count++;
        if (restart_interval > 0 && mcu_count % restart_interval == 0 &&
            mcu_count < (uint32_t)it->total_mcus) { /* handle_restart, mod.rs:1423-1445 */
            bw_flush(bw);
            bw_push(bw, 0xFF);
            bw_push(bw, (uint8_t)(0xD0 + (rst_idx & 7)));
            rst_idx = (uint8_t)((rst_idx + 1) & 7);
            py = 0; pcb = 0; pcr = 0;
        }
    }
    bw_flush(bw);
}

long po_jpeg_encode_from_coefficients(const int16_t *y, const int16_t *cb, const int16_t *cr,
                                      uint32_t w, uint32_t h, int color_type, int quality,
                                      int subsampling, uint32_t restart_interval,
                                      int optimize_huffman, uint8_t *out, size_t cap)
{
    uint8_t lum_zz[64], chr_zz[64];
    po_quant_tables(quality, lum_zz, chr_zz, NULL, NULL);
    po_huff_tables t;
    int have = 0;
    if (optimize_huffman) {
        uint64_t hist[536];
        po_jpeg_histograms(y, cb, cr, w, h, color_type, subsampling, restart_interval, hist);
        have = po_huff_optimized(hist, color_type != PO_GRAY, &t);
    }
    if (!have) po_huff_standard(&t);
    outbuf o = { out, 0, cap, 0 };
    write_headers(&o, w, h, color_type, subsampling, lum_zz, chr_zz, &t, restart_interval);
    if (o.overflow) return -6;
    frame_iter it;
    frame_iter_init(&it, y, cb, cr, w, h, color_type, subsampling);
    bitw bw = { out + o.len, 0, cap - o.len, 0, 8, 0 };
    encode_scan(&bw, &it, &t, restart_interval);
    if (bw.overflow) return -6;
    o.len += bw.len;
    ob_u16(&o, 0xFFD9);                                   /* write_eoi :444 */
    if (o.overflow) return -6;
    return (long)o.len;
}

/* encode_into, mod.rs:328-447 */
long po_jpeg_encode(const uint8_t *data, size_t data_len, uint32_t w, uint32_t h,
                    int color_type, int quality, int subsampling, uint32_t restart_interval,
                    int optimize_huffman, uint8_t *out, size_t cap)
{
    if (quality == 0 || quality > 100 || quality < 0) return -1;
    if (w == 0 || h == 0) return -2;
    if (w > 65535 || h > 65535) return -3;
    size_t bpp;
    if (color_type == PO_RGB) bpp = 3;
    else if (color_type == PO_GRAY) bpp = 1;
    else return -4;
    if (data_len != (size_t)w * h * bpp) return -5;
    float lum[64], chr[64];
    po_quant_tables(quality, NULL, NULL, lum, chr);
    size_t ny, nc;
    po_jpeg_block_counts(w, h, color_type, subsampling, &ny, &nc);
    int16_t *y = (int16_t *)malloc(ny * 64 * sizeof(int16_t));
    int16_t *cb = nc ? (int16_t *)malloc(nc * 64 * sizeof(int16_t)) : NULL;
    int16_t *cr = nc ? (int16_t *)malloc(nc * 64 * sizeof(int16_t)) : NULL;
    po_jpeg_coefficients(data, w, h, color_type, subsampling, lum, chr, y, cb, cr, 0, 0);
    long r = po_jpeg_encode_from_coefficients(y, cb, cr, w, h, color_type, quality, subsampling,
                                              restart_interval, optimize_huffman, out, cap);
    free(y); free(cb); free(cr);
    return r;
}

/* ------------------------------------------------------------------------------------------
 * PNG filters — src/simd/fallback.rs:93-159 (normative scalar semantics)
 * ---------------------------------------------------------------------------------------- */
void po_filter_sub(const uint8_t *row, size_t n, size_t bpp, uint8_t *out)
{
    for (size_t i = 0; i < n; i++) {
        uint8_t left = i >= bpp ? row[i - bpp] : 0;
        out[i] = (uint8_t)(row[i] - left);
    }
}

void po_filter_up(const uint8_t *row, const uint8_t *prev, size_t n, uint8_t *out)
{
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)(row[i] - prev[i]);
}

void po_filter_average(const uint8_t *row, const uint8_t *prev, size_t n, size_t bpp, uint8_t *out)
{
    for (size_t i = 0; i < n; i++) {
        uint16_t left = i >= bpp ? row[i - bpp] : 0;
        uint16_t above = prev[i];
        uint8_t avg = (uint8_t)((left + above) / 2);
        out[i] = (uint8_t)(row[i] - avg);
    }
}

/* fallback_paeth_predictor, fallback.rs:143-159 */
uint8_t po_paeth_predictor(uint8_t a8, uint8_t b8, uint8_t c8)
{
    int16_t a = a8, b = b8, c = c8;
    int16_t p = (int16_t)(a + b - c);
    int16_t pa = (int16_t)abs(p - a), pb = (int16_t)abs(p - b), pc = (int16_t)abs(p - c);
    if (pa <= pb && pa <= pc) return a8;
    if (pb <= pc) return b8;
    return c8;
}

void po_filter_paeth(const uint8_t *row, const uint8_t *prev, size_t n, size_t bpp, uint8_t *out)
{
    for (size_t i = 0; i < n; i++) {
        uint8_t left = i >= bpp ? row[i - bpp] : 0;
        uint8_t above = prev[i];
        uint8_t ul = i >= bpp ? prev[i - bpp] : 0;
        out[i] = (uint8_t)(row[i] - po_paeth_predictor(left, above, ul));
    }
}

/* score_filter, fallback.rs:93-98: sum |i8| as u64 */
uint64_t po_score_filter(const uint8_t *f, size_t n)
{
    uint64_t s = 0;
    for (size_t i = 0; i < n; i++) {
        int8_t v = (int8_t)f[i];
        s += (uint64_t)(v < 0 ? -(int)v : (int)v);
    }
    return s;
}

/* score_bigrams, png/filter.rs:635-649: distinct adjacent byte pairs */
size_t po_score_bigrams(const uint8_t *f, size_t n)
{
    static __thread uint8_t seen[65536];
    memset(seen, 0, sizeof seen);
    size_t cnt = 0;
    for (size_t i = 0; i + 1 < n; i++) {
        unsigned key = ((unsigned)f[i] << 8) | f[i + 1];
        if (!seen[key]) { seen[key] = 1; cnt++; }
    }
    return cnt;
}

typedef struct { uint8_t *none, *sub, *up, *avg, *paeth; } scratch5;

/* adaptive_filter, png/filter.rs:302-393 (MinSum aliases it, :396-404). */
static void adaptive_filter(const uint8_t *row, const uint8_t *prev, size_t n, size_t bpp,
                            uint8_t *out, scratch5 *s)
{
    uint8_t best = PO_F_NONE;
    uint64_t best_score = UINT64_MAX;
    uint64_t early = (uint64_t)n / 4 + 1;
    const uint8_t *win = NULL;

    memcpy(s->none, row, n);
    uint64_t sc = po_score_filter(s->none, n);
    if (sc < best_score) {
        best_score = sc; best = PO_F_NONE;
        if (best_score <= early) { win = s->none; goto emit; }
    }
    if (best_score == 0) { win = s->none; goto emit; }

    po_filter_sub(row, n, bpp, s->sub);
    sc = po_score_filter(s->sub, n);
    if (sc < best_score) {
        best_score = sc; best = PO_F_SUB;
        if (best_score == 0 || best_score <= early) { win = s->sub; goto emit; }
    }
    po_filter_up(row, prev, n, s->up);
    sc = po_score_filter(s->up, n);
    if (sc < best_score) {
        best_score = sc; best = PO_F_UP;
        if (best_score == 0 || best_score <= early) { win = s->up; goto emit; }
    }
    po_filter_average(row, prev, n, bpp, s->avg);
    sc = po_score_filter(s->avg, n);
    if (sc < best_score) {
        best_score = sc; best = PO_F_AVERAGE;
        if (best_score == 0 || best_score <= early) { win = s->avg; goto emit; }
    }
    po_filter_paeth(row, prev, n, bpp, s->paeth);
    sc = po_score_filter(s->paeth, n);
    if (sc < best_score) best = PO_F_PAETH;
    switch (best) {
    case PO_F_NONE: win = s->none; break;
    case PO_F_SUB: win = s->sub; break;
    case PO_F_UP: win = s->up; break;
    case PO_F_AVERAGE: win = s->avg; break;
    default: win = s->paeth; break;
    }
emit:
    out[0] = best;
    memcpy(out + 1, win, n);
}

/* bigrams_filter, png/filter.rs:410-471 */
static void bigrams_filter(const uint8_t *row, const uint8_t *prev, size_t n, size_t bpp,
                           uint8_t *out, scratch5 *s)
{
    uint8_t best = PO_F_NONE;
    size_t best_score = SIZE_MAX, sc;
    memcpy(s->none, row, n);
    sc = po_score_bigrams(s->none, n);
    if (sc < best_score) { best_score = sc; best = PO_F_NONE; }
    po_filter_sub(row, n, bpp, s->sub);
    sc = po_score_bigrams(s->sub, n);
    if (sc < best_score) { best_score = sc; best = PO_F_SUB; }
    po_filter_up(row, prev, n, s->up);
    sc = po_score_bigrams(s->up, n);
    if (sc < best_score) { best_score = sc; best = PO_F_UP; }
    po_filter_average(row, prev, n, bpp, s->avg);
    sc = po_score_bigrams(s->avg, n);
    if (sc < best_score) { best_score = sc; best = PO_F_AVERAGE; }
    po_filter_paeth(row, prev, n, bpp, s->paeth);
    sc = po_score_bigrams(s->paeth, n);
    if (sc < best_score) best = PO_F_PAETH;
    const uint8_t *win = best == PO_F_NONE ? s->none : best == PO_F_SUB ? s->sub
                       : best == PO_F_UP ? s->up : best == PO_F_AVERAGE ? s->avg : s->paeth;
    out[0] = best;
    memcpy(out + 1, win, n);
}

/* adaptive_filter_fast, png/filter.rs:474-527 */
static void adaptive_filter_fast(const uint8_t *row, const uint8_t *prev, size_t n, size_t bpp,
                                 uint8_t *out, scratch5 *s)
{
    po_filter_sub(row, n, bpp, s->sub);
    uint8_t best = PO_F_SUB;
    uint64_t best_score = po_score_filter(s->sub, n);
    uint64_t early = (uint64_t)n / 8 + 1;
    if (best_score <= early) { out[0] = best; memcpy(out + 1, s->sub, n); return; }

    po_filter_up(row, prev, n, s->up);
    uint64_t up_score = po_score_filter(s->up, n);
    if (up_score < best_score) { best_score = up_score; best = PO_F_UP; }
    if (best_score <= early) {
        out[0] = best;
        memcpy(out + 1, best == PO_F_SUB ? s->sub : s->up, n);
        return;
    }
    po_filter_paeth(row, prev, n, bpp, s->paeth);
    uint64_t pscore = po_score_filter(s->paeth, n);
    if (pscore < best_score) best = PO_F_PAETH;
    out[0] = best;
    memcpy(out + 1, best == PO_F_SUB ? s->sub : best == PO_F_UP ? s->up : s->paeth, n);
}

/* filter_row, png/filter.rs:529-571 */
static void filter_row(const uint8_t *row, const uint8_t *prev, size_t n, size_t bpp,
                       int strategy, uint8_t *out, scratch5 *s)
{
    switch (strategy) {
    case PO_F_NONE: out[0] = 0; memcpy(out + 1, row, n); break;
    case PO_F_SUB: out[0] = 1; po_filter_sub(row, n, bpp, out + 1); break;
    case PO_F_UP: out[0] = 2; po_filter_up(row, prev, n, out + 1); break;
    case PO_F_AVERAGE: out[0] = 3; po_filter_average(row, prev, n, bpp, out + 1); break;
    case PO_F_PAETH: out[0] = 4; po_filter_paeth(row, prev, n, bpp, out + 1); break;
    case PO_F_MINSUM:
    case PO_F_ADAPTIVE: adaptive_filter(row, prev, n, bpp, out, s); break;
    case PO_F_ADAPTIVE_FAST: adaptive_filter_fast(row, prev, n, bpp, out, s); break;
    default: bigrams_filter(row, prev, n, bpp, out, s); break;
    }
}

/* apply_filters_with_row_bytes, png/filter.rs:64-206 + apply_filters_parallel :574-608 */
void po_apply_filters(const uint8_t *data, uint32_t width, uint32_t height, size_t row_bytes,
                      size_t bpp, int strategy, int parallel_feature, uint8_t *out,
                      uint32_t row0, uint32_t row1)
{
    size_t frs = row_bytes + 1;
    uint8_t *zero_row = (uint8_t *)calloc(row_bytes ? row_bytes : 1, 1);
    uint8_t *sb = (uint8_t *)malloc(5 * (row_bytes ? row_bytes : 1));
    scratch5 s = { sb, sb + row_bytes, sb + 2 * row_bytes, sb + 3 * row_bytes, sb + 4 * row_bytes };

    size_t area = (size_t)width * (size_t)height;
    int adaptive_like = strategy == PO_F_ADAPTIVE || strategy == PO_F_ADAPTIVE_FAST ||
                        strategy == PO_F_BIGRAMS;
    if (area <= 4096 && adaptive_like) { strategy = PO_F_SUB; adaptive_like = 0; } /* :77-86 */

    uint32_t r0 = 0, r1 = height;
    if (row1) { r0 = row0; r1 = row1 < height ? row1 : height; }

    if (parallel_feature && height > 32 && adaptive_like) {
        /* apply_filters_parallel: every row independent, prev = raw previous row */
        for (uint32_t y = r0; y < r1; y++) {
            const uint8_t *row = data + (size_t)y * row_bytes;
            const uint8_t *prev = y == 0 ? zero_row : data + (size_t)(y - 1) * row_bytes;
            filter_row(row, prev, row_bytes, bpp, strategy, out + (size_t)y * frs, &s);
        }
    } else {
        /* sequential loop :113-184; AdaptiveFast is biased to the previous winner :147-166 */
        int last_adaptive = -1;
        for (uint32_t y = 0; y < r1; y++) {
            const uint8_t *row = data + (size_t)y * row_bytes;
            const uint8_t *prev = y == 0 ? zero_row : data + (size_t)(y - 1) * row_bytes;
            uint8_t *o = out + (size_t)y * frs;
            if (strategy == PO_F_ADAPTIVE_FAST) {
                int st = PO_F_ADAPTIVE_FAST;
                if (last_adaptive == PO_F_SUB) st = PO_F_SUB;
                else if (last_adaptive == PO_F_UP) st = PO_F_UP;
                else if (last_adaptive == PO_F_PAETH) st = PO_F_PAETH;
                filter_row(row, prev, row_bytes, bpp, st, o, &s);
                last_adaptive = o[0];
            } else {
                if (y < r0) continue;
                filter_row(row, prev, row_bytes, bpp, strategy, o, &s);
            }
        }
    }
    free(zero_row);
    free(sb);
}

/* ------------------------------------------------------------------------------------------
 * optimize_alpha pre-pass — src/png/mod.rs:633-671
 * ---------------------------------------------------------------------------------------- */
void po_optimize_alpha(uint8_t *data, size_t n_bytes, int color_type)
{
    if (color_type == PO_RGBA) {
        for (size_t i = 0; i + 4 <= n_bytes; i += 4)
            if (data[i + 3] == 0) data[i] = data[i + 1] = data[i + 2] = 0;
    } else if (color_type == PO_GRAY_ALPHA) {
        for (size_t i = 0; i + 2 <= n_bytes; i += 2)
            if (data[i + 1] == 0) data[i] = 0;
    }
}

/* ------------------------------------------------------------------------------------------
 * Checksums — src/compress/adler32.rs:26-47, src/simd/fallback.rs:8-58
 * ---------------------------------------------------------------------------------------- */
uint32_t po_adler32(const uint8_t *data, size_t n)
{
    const uint32_t MOD = 65521u;
    const size_t NMAX = 5552;
    uint32_t s1 = 1, s2 = 0;
    size_t i = 0;
    while (i < n) {
        size_t end = i + NMAX < n ? i + NMAX : n;
        for (; i < end; i++) { s1 += data[i]; s2 += s1; }
        s1 %= MOD;
        s2 %= MOD;
    }
    return (s2 << 16) | s1;
}

uint32_t po_crc32(const uint8_t *data, size_t n)
{
    static uint32_t table[256];
    static int init = 0;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int j = 0; j < 8; j++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
            table[i] = c;
        }
        init = 1;
    }
    uint32_t crc = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) crc = (crc >> 8) ^ table[(crc ^ data[i]) & 0xFF];
    return crc ^ 0xFFFFFFFFu;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic inputs — tests/support/synthetic.rs:74-85,183-197
 * ---------------------------------------------------------------------------------------- */
void po_gen_gradient_rgb(uint32_t w, uint32_t h, uint8_t *out)
{
    uint32_t wd = w ? w : 1, hd = h ? h : 1, sd = (w + h) ? (w + h) : 1;
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            *out++ = (uint8_t)((x * 255) / wd);
            *out++ = (uint8_t)((y * 255) / hd);
            *out++ = (uint8_t)(((x + y) * 127) / sd);
        }
}

void po_gen_noise(uint32_t w, uint32_t h, uint32_t channels, uint32_t seed, uint8_t *out)
{
    uint32_t state = seed;
    size_t n = (size_t)w * h * channels;
    for (size_t i = 0; i < n; i++) {
        state = state * 1103515245u + 12345u;
        out[i] = (uint8_t)(state >> 16);
    }
}
