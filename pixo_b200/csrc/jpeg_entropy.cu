// jpeg_entropy.cu — baseline Huffman entropy coding of the quantised coefficient arrays ON the
// GPU, so that only finished scan bytes cross PCIe.
//
// Restates the bit stream of pixo's sequential coder byte for byte:
//   encode_block            src/jpeg/huffman.rs:423-481 (DC diff, (run,size) symbols, ZRL, EOB)
//   category / encode_value src/jpeg/huffman.rs:394-418
//   BitWriterMsb            src/bits.rs:195-290 (MSB-first, 0xFF -> 0xFF00 stuffing, 1-padding)
//   encode_scan             src/jpeg/mod.rs:1408-1563 (scan order: Y..,Cb,Cr per MCU)
//
// The DC predictor of a block is the previous block of the same component in the coefficient
// array, so every block's code length is independent:
//   1. k_huff_len    one thread per block: bits of its code                        (u32/block)
//   2. k_scan_local / k_scan_top   exclusive scan of those lengths in scan order   (bit offsets)
//   3. k_huff_emit   one thread per block: writes its bits at its offset into a zeroed raw
//                    buffer (big-endian words; the two boundary words by atomicOr)
//   4. k_ff_count / k_ff_top / k_stuff   pad the last byte with 1s, count 0xFF bytes per chunk,
//                    scan, and copy with 0x00 inserted after every 0xFF
// Restart intervals stay on the host coder (jpeg_host.cpp): they need per-interval padding.
#include "common.cuh"
#include "jpeg_host.hpp"

namespace pixo {
namespace {

__host__ __device__ constexpr int zz2(int i)
{
    constexpr int t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                           12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                           35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                           58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return t[i];
}

struct HuffDev {
    uint16_t dc_code[2][12];
    uint16_t ac_code[2][256];
    uint8_t dc_len[2][12];
    uint8_t ac_len[2][256];
};

struct EntParams {
    const int16_t *y, *cb, *cr;
    size_t y_stride, c_stride;     // int16 elements between images
    uint32_t bpm;                  // blocks per MCU in scan order: 6 (4:2:0), 3 (4:4:4), 1 (gray)
    uint32_t y_per_mcu;            // 4, 1, 1
    uint64_t nblocks;              // per image, scan order
    uint32_t *blk_bits;            // [n][nblocks_padded] lengths, then in-chunk exclusive offsets
    uint64_t blk_pitch;
    uint32_t *chunk_tot;           // [n][nchunks]
    uint64_t *chunk_base;          // [n][nchunks]
    uint32_t nchunks;
    uint64_t *total_bits;          // [n]
    uint8_t *raw;                  // [n][raw_cap]
    uint64_t raw_cap;
    uint32_t *ff_tot;              // [n][ff_chunks]
    uint64_t *ff_base;             // [n][ff_chunks]
    uint32_t ff_chunks;
    uint8_t *out;                  // [n][out_cap]
    uint64_t out_cap;
    uint64_t *out_len;             // [n] final byte count
    uint32_t *overflow;            // [n] set when a capacity was exceeded
};

constexpr int SCAN_CH = 2048;      // blocks per scan chunk (256 threads x 8)
constexpr int FF_CH = 4096;        // raw bytes per stuffing chunk (128 threads x 32)

__device__ __forceinline__ int cat16(int v)
{
    const int a = v < 0 ? -v : v;
    return 32 - __clz(a);
}

// locate block `s` (scan order) : pointer to its 64 coefficients, its table (0 lum / 1 chroma)
// and the DC of its predecessor in the same component (0 at the start of the scan)
__device__ __forceinline__ const int16_t *locate(const EntParams &P, uint32_t img, uint64_t s,
                                                 int &tbl, int &prev_dc)
{
    const uint64_t m = s / P.bpm;
    const uint32_t k = (uint32_t)(s - m * P.bpm);
    const int16_t *arr;
    uint64_t idx;
    if (k < P.y_per_mcu) { arr = P.y + (size_t)img * P.y_stride; idx = m * P.y_per_mcu + k; tbl = 0; }
    else if (k == P.y_per_mcu) { arr = P.cb + (size_t)img * P.c_stride; idx = m; tbl = 1; }
    else { arr = P.cr + (size_t)img * P.c_stride; idx = m; tbl = 1; }
    prev_dc = idx ? arr[(idx - 1) * 64] : 0;
    return arr + idx * 64;
}

struct BitSink {
    uint32_t *words;
    uint64_t widx, wcap;
    uint64_t acc;
    int filled;
    bool shared_first;
    __device__ __forceinline__ void put(uint32_t code, int len)
    {
        if (len == 0) return;
        acc |= (uint64_t)code << (64 - filled - len);
        filled += len;
        if (filled >= 32) {
            const uint32_t w = __byte_perm((uint32_t)(acc >> 32), 0, 0x0123);  // big-endian bytes
            if (widx < wcap) {
                if (shared_first) atomicOr(&words[widx], w); else words[widx] = w;
            }
            shared_first = false;
            ++widx;
            acc <<= 32;
            filled -= 32;
        }
    }
    __device__ __forceinline__ void finish()
    {
        if (filled > 0 && widx < wcap)
            atomicOr(&words[widx], __byte_perm((uint32_t)(acc >> 32), 0, 0x0123));
    }
};

// Symbolise one block held in 32 packed words (natural order) — encode_block's walk.
template <bool EMIT>
__device__ __forceinline__ uint32_t walk_block(const uint32_t (&w)[32], int prev_dc, int tbl,
                                               const HuffDev &T, BitSink *sink)
{
    uint32_t bits = 0;
    {
        const int dc = (int)(int16_t)(w[0] & 0xFFFF);
        const int diff = (int)(int16_t)(dc - prev_dc);
        const int cat = cat16(diff);
        const int len = T.dc_len[tbl][cat];
        bits += len + cat;
        if (EMIT) {
            const uint32_t amp = (uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << cat) - 1u);
            sink->put(((uint32_t)T.dc_code[tbl][cat] << cat) | amp, len + cat);
        }
    }
    int run = 0;
#pragma unroll
    for (int i = 1; i < 64; ++i) {
        const int nat = zz2(i);
        const uint32_t word = w[nat >> 1];
        const int c = (int)(int16_t)((nat & 1) ? (word >> 16) : (word & 0xFFFF));
        if (c == 0) {
            ++run;
        } else {
            while (run >= 16) {
                bits += T.ac_len[tbl][0xF0];
                if (EMIT) sink->put(T.ac_code[tbl][0xF0], T.ac_len[tbl][0xF0]);
                run -= 16;
            }
            const int cat = cat16(c);
            const int rs = (run << 4) | cat;
            const int len = T.ac_len[tbl][rs];
            bits += len + cat;
            if (EMIT) {
                const uint32_t amp = (uint32_t)(c < 0 ? c - 1 : c) & ((1u << cat) - 1u);
                sink->put(((uint32_t)T.ac_code[tbl][rs] << cat) | amp, len + cat);
            }
            run = 0;
        }
    }
    if (run > 0) {
        bits += T.ac_len[tbl][0];
        if (EMIT) sink->put(T.ac_code[tbl][0], T.ac_len[tbl][0]);
    }
    return bits;
}

__device__ __forceinline__ void load_block(const int16_t *p, uint32_t (&w)[32])
{
    const uint4 *src = reinterpret_cast<const uint4 *>(p);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint4 t = __ldg(src + k);
        w[k * 4] = t.x; w[k * 4 + 1] = t.y; w[k * 4 + 2] = t.z; w[k * 4 + 3] = t.w;
    }
}

__global__ void __launch_bounds__(128)
k_huff_len(const __grid_constant__ EntParams P, const __grid_constant__ HuffDev Tp)
{
    __shared__ HuffDev T;
    for (int i = threadIdx.x; i < (int)(sizeof(HuffDev) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t *>(&T)[i] = reinterpret_cast<const uint32_t *>(&Tp)[i];
    __syncthreads();
    const uint32_t img = blockIdx.y;
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.nblocks) return;
    int tbl, prev;
    const int16_t *blk = locate(P, img, s, tbl, prev);
    uint32_t w[32];
    load_block(blk, w);
    P.blk_bits[(size_t)img * P.blk_pitch + s] = walk_block<false>(w, prev, tbl, T, nullptr);
}

// exclusive scan of SCAN_CH values per CTA (in place) + chunk totals
__global__ void __launch_bounds__(256) k_scan_local(const __grid_constant__ EntParams P)
{
    __shared__ uint32_t wsum[8];
    const uint32_t img = blockIdx.y, chunk = blockIdx.x;
    uint32_t *v = P.blk_bits + (size_t)img * P.blk_pitch + (size_t)chunk * SCAN_CH;
    const uint64_t base = (uint64_t)chunk * SCAN_CH;
    const int t = threadIdx.x;
    uint32_t x[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint64_t i = base + (uint64_t)t * 8 + k;
        x[k] = i < P.nblocks ? v[t * 8 + k] : 0u;
        sum += x[k];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if ((t & 31) >= o) inc += n;
    }
    if ((t & 31) == 31) wsum[t >> 5] = inc;
    __syncthreads();
    uint32_t wbase = 0;
    for (int k = 0; k < (t >> 5); ++k) wbase += wsum[k];
    uint32_t run = wbase + inc - sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint64_t i = base + (uint64_t)t * 8 + k;
        if (i < P.nblocks) v[t * 8 + k] = run;
        run += x[k];
    }
    if (t == 255) P.chunk_tot[(size_t)img * P.nchunks + chunk] = wbase + inc;
}

// one CTA per image: exclusive scan of a u32 array of `n` totals into u64 bases + grand total
__device__ void scan_top(const uint32_t *tot, uint64_t *base, uint32_t n, unsigned long long *grand)
{
    __shared__ unsigned long long wsum[8];
    __shared__ unsigned long long carry;
    const int t = threadIdx.x;
    if (t == 0) carry = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += 256) {
        const uint32_t i = i0 + t;
        const unsigned long long x = i < n ? tot[i] : 0ull;
        unsigned long long inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long m = __shfl_up_sync(0xffffffffu, inc, o);
            if ((t & 31) >= o) inc += m;
        }
        if ((t & 31) == 31) wsum[t >> 5] = inc;
        __syncthreads();
        unsigned long long wb = 0;
        for (int k = 0; k < (t >> 5); ++k) wb += wsum[k];
        if (i < n) base[i] = carry + wb + inc - x;
        __syncthreads();
        if (t == 255) carry += wb + inc;
        __syncthreads();
    }
    if (t == 0) *grand = carry;
}

__global__ void __launch_bounds__(256) k_scan_top(const __grid_constant__ EntParams P)
{
    const uint32_t img = blockIdx.x;
    scan_top(P.chunk_tot + (size_t)img * P.nchunks, P.chunk_base + (size_t)img * P.nchunks, P.nchunks,
             reinterpret_cast<unsigned long long *>(P.total_bits + img));
    __syncthreads();
    if (threadIdx.x == 0 && (P.total_bits[img] + 7) / 8 > P.raw_cap) P.overflow[img] = 1;
}

__global__ void __launch_bounds__(128)
k_huff_emit(const __grid_constant__ EntParams P, const __grid_constant__ HuffDev Tp)
{
    __shared__ HuffDev T;
    for (int i = threadIdx.x; i < (int)(sizeof(HuffDev) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t *>(&T)[i] = reinterpret_cast<const uint32_t *>(&Tp)[i];
    __syncthreads();
    const uint32_t img = blockIdx.y;
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.nblocks) return;
    int tbl, prev;
    const int16_t *blk = locate(P, img, s, tbl, prev);
    uint32_t w[32];
    load_block(blk, w);
    const uint64_t off = P.chunk_base[(size_t)img * P.nchunks + s / SCAN_CH] +
                         P.blk_bits[(size_t)img * P.blk_pitch + s];
    BitSink sink;
    sink.words = reinterpret_cast<uint32_t *>(P.raw + (size_t)img * P.raw_cap);
    sink.wcap = P.raw_cap / 4;
    sink.widx = off >> 5;
    sink.filled = (int)(off & 31);
    sink.acc = 0;
    sink.shared_first = sink.filled != 0;
    walk_block<true>(w, prev, tbl, T, &sink);
    sink.finish();
}

// ---- 0xFF stuffing ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_ff_count(const __grid_constant__ EntParams P)
{
    __shared__ uint32_t wsum[4];
    const uint32_t img = blockIdx.y, chunk = blockIdx.x;
    const uint64_t bits = P.total_bits[img];
    const uint64_t nbytes = (bits + 7) / 8;
    uint8_t *raw = P.raw + (size_t)img * P.raw_cap;
    const uint64_t b0 = (uint64_t)chunk * FF_CH + (uint64_t)threadIdx.x * 32;
    uint32_t cnt = 0;
    if (b0 < nbytes && nbytes <= P.raw_cap) {
        // BitWriterMsb::flush: pad the final partial byte with 1s (src/bits.rs:261-272)
        if (nbytes - 1 >= b0 && nbytes - 1 < b0 + 32 && (bits & 7))
            raw[nbytes - 1] |= (uint8_t)((1u << (8 - (bits & 7))) - 1u);
        const uint4 a = *reinterpret_cast<const uint4 *>(raw + b0);
        const uint4 b = *reinterpret_cast<const uint4 *>(raw + b0 + 16);
        const uint32_t wv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint64_t wb = b0 + 4 * k;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (wb + j < nbytes && ((wv[k] >> (8 * j)) & 0xFF) == 0xFF) ++cnt;
        }
    }
    uint32_t v = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) P.ff_tot[(size_t)img * P.ff_chunks + chunk] = wsum[0] + wsum[1] + wsum[2] + wsum[3];
}

__global__ void __launch_bounds__(256) k_ff_top(const __grid_constant__ EntParams P)
{
    const uint32_t img = blockIdx.x;
    __shared__ unsigned long long total_ff;
    scan_top(P.ff_tot + (size_t)img * P.ff_chunks, P.ff_base + (size_t)img * P.ff_chunks, P.ff_chunks, &total_ff);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t nbytes = (P.total_bits[img] + 7) / 8;
        const uint64_t len = nbytes + total_ff;
        P.out_len[img] = len;
        if (len > P.out_cap) P.overflow[img] = 1;
    }
}

__global__ void __launch_bounds__(128) k_stuff(const __grid_constant__ EntParams P)
{
    __shared__ uint32_t wsum[4];
    const uint32_t img = blockIdx.y, chunk = blockIdx.x;
    const uint64_t nbytes = (P.total_bits[img] + 7) / 8;
    if (P.overflow[img]) return;
    const uint8_t *raw = P.raw + (size_t)img * P.raw_cap;
    uint8_t *out = P.out + (size_t)img * P.out_cap;
    const int t = threadIdx.x;
    const uint64_t b0 = (uint64_t)chunk * FF_CH + (uint64_t)t * 32;
    uint32_t wv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t cnt = 0;
    int nvalid = 0;
    if (b0 < nbytes) {
        nvalid = (int)min((uint64_t)32, nbytes - b0);
        const uint4 a = *reinterpret_cast<const uint4 *>(raw + b0);
        const uint4 b = *reinterpret_cast<const uint4 *>(raw + b0 + 16);
        wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w; wv[4] = b.x; wv[5] = b.y; wv[6] = b.z; wv[7] = b.w;
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (4 * k + j < nvalid && ((wv[k] >> (8 * j)) & 0xFF) == 0xFF) ++cnt;
    }
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if ((t & 31) >= o) inc += n;
    }
    if ((t & 31) == 31) wsum[t >> 5] = inc;
    __syncthreads();
    uint32_t wb = 0;
    for (int k = 0; k < (t >> 5); ++k) wb += wsum[k];
    uint64_t o = b0 + P.ff_base[(size_t)img * P.ff_chunks + chunk] + wb + inc - cnt;
    if (cnt == 0 && nvalid == 32 && (o & 3) == 0) {
        uint32_t *o32 = reinterpret_cast<uint32_t *>(out + o);
#pragma unroll
        for (int k = 0; k < 8; ++k) o32[k] = wv[k];
        return;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (4 * k + j < nvalid) {
                const uint8_t byte = (uint8_t)(wv[k] >> (8 * j));
                out[o++] = byte;
                if (byte == 0xFF) out[o++] = 0x00;
            }
}

}  // namespace

// Device scratch layout for n images (all sizes in bytes, 256-aligned)
struct EntropyPlan {
    size_t blk_pitch, nchunks, ff_chunks;
    size_t off_blk, off_ctot, off_cbase, off_tbits, off_fftot, off_ffbase, off_outlen, off_ovf, off_raw, total;
};

static size_t a256(size_t v) { return (v + 255) / 256 * 256; }

static EntropyPlan plan_entropy(uint32_t n, uint64_t nblocks, uint64_t raw_cap)
{
    EntropyPlan p;
    p.nchunks = (size_t)((nblocks + SCAN_CH - 1) / SCAN_CH);
    p.blk_pitch = p.nchunks * SCAN_CH;
    p.ff_chunks = (size_t)((raw_cap + FF_CH - 1) / FF_CH);
    size_t o = 0;
    p.off_blk = o; o += a256((size_t)n * p.blk_pitch * 4);
    p.off_ctot = o; o += a256((size_t)n * p.nchunks * 4);
    p.off_cbase = o; o += a256((size_t)n * p.nchunks * 8);
    p.off_tbits = o; o += a256((size_t)n * 8);
    p.off_fftot = o; o += a256((size_t)n * p.ff_chunks * 4);
    p.off_ffbase = o; o += a256((size_t)n * p.ff_chunks * 8);
    p.off_outlen = o; o += a256((size_t)n * 8);
    p.off_ovf = o; o += a256((size_t)n * 4);
    p.off_raw = o; o += (size_t)n * raw_cap;
    p.total = o;
    return p;
}

size_t entropy_scratch_bytes(uint32_t n, const FrameGeometry &g, uint64_t raw_cap)
{
    return plan_entropy(n, g.ny + 2 * g.nc, raw_cap).total;
}

// Enqueue the whole entropy stage for n images on ctx->stream.  d_scratch: entropy_scratch_bytes.
// d_out: n * out_cap bytes of scan data; *d_out_len / *d_overflow point into the scratch.
int launch_jpeg_entropy(pixo_b200_ctx *ctx, const int16_t *d_y, size_t y_stride, const int16_t *d_cb,
                        const int16_t *d_cr, size_t c_stride, uint32_t n, const FrameGeometry &g,
                        const HuffTables &t, uint8_t *d_scratch, uint64_t raw_cap, uint8_t *d_out,
                        uint64_t out_cap, uint64_t **d_out_len, uint32_t **d_overflow)
{
    const uint64_t nblocks = g.ny + 2 * g.nc;
    const EntropyPlan pl = plan_entropy(n, nblocks, raw_cap);
    EntParams P;
    P.y = d_y; P.cb = d_cb; P.cr = d_cr; P.y_stride = y_stride; P.c_stride = c_stride;
    P.bpm = g.y_per_mcu + (g.has_chroma ? 2 : 0);
    P.y_per_mcu = g.y_per_mcu;
    P.nblocks = nblocks;
    P.blk_bits = reinterpret_cast<uint32_t *>(d_scratch + pl.off_blk);
    P.blk_pitch = pl.blk_pitch;
    P.chunk_tot = reinterpret_cast<uint32_t *>(d_scratch + pl.off_ctot);
    P.chunk_base = reinterpret_cast<uint64_t *>(d_scratch + pl.off_cbase);
    P.nchunks = (uint32_t)pl.nchunks;
    P.total_bits = reinterpret_cast<uint64_t *>(d_scratch + pl.off_tbits);
    P.raw = d_scratch + pl.off_raw;
    P.raw_cap = raw_cap;
    P.ff_tot = reinterpret_cast<uint32_t *>(d_scratch + pl.off_fftot);
    P.ff_base = reinterpret_cast<uint64_t *>(d_scratch + pl.off_ffbase);
    P.ff_chunks = (uint32_t)pl.ff_chunks;
    P.out = d_out; P.out_cap = out_cap;
    P.out_len = reinterpret_cast<uint64_t *>(d_scratch + pl.off_outlen);
    P.overflow = reinterpret_cast<uint32_t *>(d_scratch + pl.off_ovf);
    *d_out_len = P.out_len;
    *d_overflow = P.overflow;

    HuffDev T;
    memset(&T, 0, sizeof T);
    for (int k = 0; k < 2; ++k) {
        for (int i = 0; i < 12; ++i) { T.dc_code[k][i] = t.code[k][i]; T.dc_len[k][i] = t.len[k][i]; }
        for (int i = 0; i < 256; ++i) { T.ac_code[k][i] = t.code[2 + k][i]; T.ac_len[k][i] = t.len[2 + k][i]; }
    }
    if (n > 65535) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "entropy stage: too many images per call");
    cudaStream_t st = ctx->stream;
    PIXO_CUDA(ctx, cudaMemsetAsync(P.overflow, 0, (size_t)n * 4, st));
    PIXO_CUDA(ctx, cudaMemsetAsync(P.raw, 0, (size_t)n * raw_cap, st));
    const unsigned gb = (unsigned)((nblocks + 127) / 128);
    k_huff_len<<<dim3(gb, n), 128, 0, st>>>(P, T);
    k_scan_local<<<dim3(P.nchunks, n), 256, 0, st>>>(P);
    k_scan_top<<<n, 256, 0, st>>>(P);
    k_huff_emit<<<dim3(gb, n), 128, 0, st>>>(P, T);
    k_ff_count<<<dim3(P.ff_chunks, n), 128, 0, st>>>(P);
    k_ff_top<<<n, 256, 0, st>>>(P);
    k_stuff<<<dim3(P.ff_chunks, n), 128, 0, st>>>(P);
    ctx->launches += 7;
    PIXO_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace pixo
