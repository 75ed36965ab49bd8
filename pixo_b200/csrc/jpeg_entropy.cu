// jpeg_entropy.cu — baseline Huffman entropy coding of the quantised coefficient arrays ON the
// GPU, so that only finished scan bytes cross PCIe.
//
// Restates the bit stream of pixo's sequential coder byte for byte:
//   encode_block            src/jpeg/huffman.rs:423-481 (DC diff, (run,size) symbols, ZRL, EOB)
//   category / encode_value src/jpeg/huffman.rs:394-418
//   BitWriterMsb            src/bits.rs:195-290 (MSB-first, 0xFF -> 0xFF00 stuffing, 1-padding)
//   encode_scan             src/jpeg/mod.rs:1408-1563 (scan order: Y..,Cb,Cr per MCU)
//
// One kernel, one pass over the coefficients (k_huff).  The DC predictor of a block is the
// previous block of the same component in the coefficient array, so every block's code is
// independent of the others; only its POSITION in the stream is not.  The kernel is persistent
// and every warp works alone (warp-level synchronisation only): it draws a chunk of 32
// consecutive blocks (scan order) of one image from a ticket counter, then
//   1. each lane codes its block once into a private shared-memory slot (the zig-zag reorder
//      happens in registers on the way in; a 64-bit non-zero mask drives the symbol loop, so the loop runs once per
//      non-zero coefficient and there is a single, small copy of the symbol code);
//   2. the block bit lengths are scanned in the warp; the chunk total enters a decoupled
//      look-back chain (one status word per chunk: bit count + the chunk's last 7 bits), which
//      yields the chunk's bit offset in the image's stream and the partial byte it inherits;
//   3. the slots are funnel-shifted into a shared window aligned to the stream's 32-bit words;
//   4. the chunk owns every byte whose last bit it wrote.  It counts its 0xFF bytes, a second
//      look-back chain turns those counts into the number of stuffed zeros before the chunk, and
//      the window is copied out with the 0x00s inserted, 16 bytes per store.
// Tickets are dispensed chunk-major across the images of a batch, so their chains advance side
// by side.  Nothing but the final scan bytes is written to global memory.  Restart intervals
// (handle_restart, src/jpeg/mod.rs:1423-1445) run here too: every interval is a bit stream of its
// own (chunks never straddle one, chain 1 restarts with it, its last chunk pads and appends the
// RSTn marker), while chain 2 - every byte written so far - runs across the whole image.
#include "common.cuh"
#include "jpeg_host.hpp"

namespace pixo {
namespace {

// natural index of zig-zag position i (src/jpeg/quantize.rs:18-22)
__host__ __device__ constexpr int zz_nat(int i)
{
    constexpr int t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                           12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                           35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                           58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return t[i];
}

// Huffman tables as the symbol loop wants them: one word per symbol,
//   entry = (code << (32 - len)) | (len + cat),   cat = symbol & 15 (AC) or the DC category
// i.e. the code left-aligned in the upper half-word and the total field width (code + amplitude
// bits) in the low five bits.
// AC entries sit at word  run * 12 + (cat - 1)  (cat 1..10): the 32 entries a warp asks for most
// (run < 8, cat <= 4) fall into 32 different shared-memory banks, so the per-symbol lookup is
// conflict-free for typical blocks (the plain run * 16 + cat layout put runs 0/2/4/6 on the same
// banks: 8.8 M excess wavefronts per 32 4K frames, profiles/r01).  ZRL and EOB follow the grid.
constexpr int AC_STRIDE = 12, AC_ZRL = 190, AC_EOB = 191, AC_WORDS = 192;
struct HuffDev {
    uint32_t dc[2][12];
    uint32_t ac[2][AC_WORDS];
};

struct EntParams {
    const int16_t *y, *cb, *cr;    // blocks in natural order, as compute_all_coefficients returns them
    size_t y_stride, c_stride;     // int16 elements between images
    uint32_t bpm;                  // blocks per MCU in scan order: 6 (4:2:0), 3 (4:4:4), 1 (gray)
    uint32_t y_per_mcu;            // 4, 1, 1
    uint32_t nblocks;              // per image, scan order
    uint32_t nchunks;              // per image
    uint32_t rst_mcus;             // restart interval in MCUs, 0 = none (a chunk never straddles an interval)
    uint32_t rst_blocks;           // ... in blocks
    uint32_t cpi;                  // chunks per full interval
    uint32_t nimages;
    unsigned long long *st_bits;   // [n][nchunks] look-back chain 1: stream bits
    unsigned long long *st_ff;     // [n][nchunks] look-back chain 2: 0xFF bytes
    uint32_t *ticket;              // chunk dispenser (launch order == dependency order)
    uint8_t *out;                  // [n][out_cap]
    uint64_t out_cap;
    uint64_t *out_len;             // [n] final byte count
    // RAW + segments: every image is cut into seg_per_img runs of whole MCUs, each coded as a bit
    // string of its own ("pseudo image" q = image * seg_per_img + segment: status words, raw buffer,
    // bit count and tail are all indexed by q); k_seg_* splice them afterwards.  0/1 = not segmented.
    uint32_t seg_per_img;
    uint32_t nblocks_last;         // blocks of an image's last segment (the others have nblocks)
    size_t seg_y_stride, seg_c_stride;   // int16 elements between the segments of an image
    const int *dc_seed_dev;        // the same three predictors in device memory (stream-ordered callers), or null
    int dc_seed[3];                // DC predictors (Y, Cb, Cr) before block 0: 0 for a whole image, the previous
                                   // band's last DCs when the arrays are one band of a frame tiled over several GPUs
    unsigned long long *out_tail;  // RAW only: [n] the stream's last 7 bits
    uint32_t *overflow;            // [n] bit 0: out_cap was exceeded (out_len = the size needed); bit 1: a chain timed out
};

#ifndef HUFF_EMIT_V
#define HUFF_EMIT_V 1      // 0: byte stores (round 1); 1: aligned word stores
#endif
#ifndef HUFF_ZRL_SPLIT
#define HUFF_ZRL_SPLIT 0   // 1: a second copy of the symbol loop without the ZRL test for warps that need none
                           // (measured: 706 -> 727 us per 32 4K frames - the mask test and the larger code cost more than the
                           // four instructions per symbol save)
#endif
constexpr int CB = 32;             // blocks per chunk == one warp
static_assert(CB * 4 == 128, "the slot word stride is spelled out in code_block's PTX");
constexpr int HUFF_WARPS = 4;      // warps per CTA (they only share the tables)
#ifndef HUFF_CTAS_N
#define HUFF_CTAS_N 6
#endif
constexpr int HUFF_CTAS_PER_SM = HUFF_CTAS_N;
#ifndef HUFF_SLOT_W
#define HUFF_SLOT_W 16
#endif
constexpr int SLOT_W = HUFF_SLOT_W;  // words of a block's code kept in shared memory (16: 512 bits)
constexpr int MAX_W = 54;          // worst case: 27 + 63 * 26 = 1665 bits
constexpr int WIN_W = 256;         // stream words assembled per round (32 bytes per lane)
constexpr int WIN_B = WIN_W * 4;
static_assert(WIN_B <= SLOT_W * CB * 4, "phase A parks the assembled window in the chunk's slot buffer");
constexpr int SBUF_B = 2 * WIN_B + 48;    // stuffed bytes of a window + alignment slack + the head pad
constexpr uint32_t SPIN_LIMIT = 1u << 22;
constexpr int LB_GROUPS = 4;       // look-back window: 32 * LB_GROUPS predecessors per step (8 measured slower)

constexpr unsigned long long ST_AGG = 1ull << 62, ST_PFX = 2ull << 62;
constexpr unsigned long long ST_VAL = (1ull << 55) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long pack_status(unsigned long long flag, uint32_t tail,
                                                          unsigned long long value)
{
    return flag | ((unsigned long long)(tail & 0x7F) << 55) | (value & ST_VAL);
}

// Exclusive prefix over chunks [0, chunk) of one image (whole warp; decoupled look-back, 128
// predecessors per step, four per lane).  tail_in = the 7-bit tail published by chunk-1.
// Returns the prefix in the low 55 bits, tail_in in bits 55..61, bit 62 = the chain timed out.
// (Inlined at both call sites: an out-of-line copy measured 2-15 % slower.)
__device__ __forceinline__ unsigned long long look_back(const unsigned long long *st, int chunk, int lane)
{
    unsigned long long excl = 0;
    uint32_t tl = 0, spins = 0;
    bool first = true, fault = false;
    int base = chunk - 1;
    // Chunks finish roughly in ticket order: wait (one lane, sleeping) until the nearest
    // predecessor has published, then the wide scan below almost never has to retry.
    if (lane == 0) {
        while ((ld_status(st + base) >> 62) == 0) {
            if (++spins > SPIN_LIMIT) break;
            __nanosleep(100);
        }
    }
    __syncwarp();
    while (base >= 0) {
        unsigned long long v[LB_GROUPS];
        const unsigned long long *p = st + (base - lane);
#pragma unroll
        for (int k = 0; k < LB_GROUPS; ++k) v[k] = base - lane - 32 * k >= 0 ? ld_status(p - 32 * k) : ST_PFX;
        unsigned long long step = 0;
        bool retry = false, done = false;
#pragma unroll
        for (int k = 0; k < LB_GROUPS; ++k) {
            const uint32_t flag = (uint32_t)(v[k] >> 62);
            const uint32_t inv = __ballot_sync(0xffffffffu, flag == 0);
            const uint32_t pm = __ballot_sync(0xffffffffu, flag == 2);
            const int stop = pm ? __ffs(pm) - 1 : 32;           // nearest inclusive prefix, if any
            if (inv & ((2u << min(stop, 31)) - 1u)) { retry = true; break; }
            // aggregates are small (a chunk's bits / bytes): one 32-bit warp reduction
            step += __reduce_add_sync(0xffffffffu, lane < stop ? (uint32_t)v[k] : 0u);
            if (pm) { step += __shfl_sync(0xffffffffu, v[k], stop) & ST_VAL; done = true; break; }
        }
        if (retry) {
            if (++spins > SPIN_LIMIT) { fault = true; break; }
            __nanosleep(20);
            continue;
        }
        excl += step;
        if (first) { tl = __shfl_sync(0xffffffffu, (uint32_t)(v[0] >> 55) & 0x7Fu, 0); first = false; }
        if (done) break;
        base -= 32 * LB_GROUPS;
    }
    return (excl & ST_VAL) | ((unsigned long long)tl << 55) | (fault ? 1ull << 62 : 0ull);
}

// bits 0..15 -> even positions, bits 16..31 -> odd positions (outer perfect shuffle)
__device__ __forceinline__ uint32_t interleave16(uint32_t x)
{
    uint32_t t;
    t = (x ^ (x >> 8)) & 0x0000FF00u; x ^= t ^ (t << 8);
    t = (x ^ (x >> 4)) & 0x00F000F0u; x ^= t ^ (t << 4);
    t = (x ^ (x >> 2)) & 0x0C0C0C0Cu; x ^= t ^ (t << 2);
    t = (x ^ (x >> 1)) & 0x22222222u; x ^= t ^ (t << 1);
    return x;
}

// 0x80 in every byte of w that equals 0xFF
__device__ __forceinline__ uint32_t ff_bytes(uint32_t w)
{
    return ((w & 0x7F7F7F7Fu) + 0x01010101u) & w & 0x80808080u;
}

// ---- the symbol loop -------------------------------------------------------------------------
__device__ __forceinline__ int lds_s16(uint32_t a)
{
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v));
}
__device__ __forceinline__ uint32_t msb_index(uint32_t v)  // FLO: 31 - clz, v != 0
{
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// Codes one block (encode_block, src/jpeg/huffman.rs:423-481) into 32-bit words
//   word k -> shared [sa_slot + k * CB * 4] for k < SLOT_W; SPILL: later words -> spill[k - SLOT_W],
//   !SPILL: later words are dropped (the caller sees the length and runs the SPILL variant).
// Pending bits are kept left-aligned in `acc`; a symbol arrives left-aligned too (vl, n bits).
// Returns the block's length in bits; *acc_out = the last, partial word (left-aligned).
// ZRLS == false is the variant for warps none of whose blocks holds a zero run of 16 or more (the
// caller checks the masks): the symbol loop then carries no ZRL test and no reconvergence point.
template <bool SPILL, bool ZRLS>
__device__ __forceinline__ uint32_t code_block(uint32_t M0, uint32_t M1, int diff, const uint32_t *dctab,
                                               uint32_t sa_ac, uint32_t sa_stage, uint32_t sa_slot,
                                               uint32_t *spill, uint32_t *acc_out)
{
    uint32_t acc = 0, filled = 0;
    uint32_t sp = sa_slot;
    const uint32_t sp_end = sa_slot + SLOT_W * CB * 4;
    auto put = [&](uint32_t vl, uint32_t n) {
        if (!SPILL) {  // branch-free: a full word is stored under a predicate
            asm volatile(
                "{\n\t"
                ".reg .pred p, q;\n\t"
                ".reg .b32 t, hi, lo, tot;\n\t"
                "shr.b32 t, %4, %1;\n\t"
                "or.b32 hi, %0, t;\n\t"
                "shf.r.wrap.b32 lo, 0, %4, %1;\n\t"   // vl << (32 - filled); 0 when filled == 0
                "add.u32 tot, %1, %5;\n\t"
                "setp.ge.u32 p, tot, 32;\n\t"
                "setp.lt.and.u32 q, %2, %3, p;\n\t"
                "@q st.shared.u32 [%2], hi;\n\t"
                "@p add.u32 %2, %2, 128;\n\t"
                "selp.b32 %0, lo, hi, p;\n\t"
                "and.b32 %1, tot, 31;\n\t"
                "}"
                : "+r"(acc), "+r"(filled), "+r"(sp)
                : "r"(sp_end), "r"(vl), "r"(n));
            return;
        }
        const uint32_t hi = acc | (vl >> filled);
        const uint32_t lo = __funnelshift_r(0u, vl, filled);  // vl << (32 - filled); 0 when filled == 0
        const uint32_t total = filled + n;
        if (total >= 32u) {
            if (sp < sp_end) sts_u32(sp, hi);
            else if (SPILL) spill[(sp - sp_end) / (CB * 4)] = hi;
            sp += CB * 4;
            acc = lo;
        } else {
            acc = hi;
        }
        filled = total & 31u;
    };
    {   // DC difference
        const uint32_t a = (uint32_t)abs(diff);
        const uint32_t cat = 32u - (uint32_t)__clz(a);
        const uint32_t e = dctab[cat];
        const uint32_t amp = a ^ (((1u << cat) - 1u) & (uint32_t)(diff >> 31));
        const uint32_t n = e & 31u;
        put((e & 0xFFFF0000u) | (n ? amp << (32u - n) : 0u), n);
    }
    const uint32_t zrl = lds_u32(sa_ac + AC_ZRL * 4), eob = lds_u32(sa_ac + AC_EOB * 4);
    uint32_t nprev = ~0u;  // -(previous position) - 1
    // (no constants live across the loop: at 80 registers ptxas re-materialises them every
    // iteration - the single-bit mask comes from BMSK, the amplitude mask from a shifted sign)
    const uint32_t sa_ac_top = sa_ac + 31u * 4u;   // entry of (run, cat) = sa_ac_top + run*48 - clz(|c|)*4
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        uint32_t mb = __brev(half ? M1 : (M0 & ~1u));  // scan order == descending bit index
        const uint32_t top = half * 32 + 31;
        while (mb) {
            const uint32_t f = msb_index(mb);
            uint32_t bit;
            asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(bit) : "r"(f));   // 1 << f
            mb ^= bit;
            const uint32_t pos = top - f;
            uint32_t run = pos + nprev;
            nprev = ~pos;
            // coefficient pos sits at stage word pos >> 1, half-word pos & 1:
            // address = stage + pos * 64 - (pos & 1) * 62, spelled as two multiply-adds
            uint32_t caddr;
            asm("{\n\t.reg .b32 h, x;\n\tand.b32 h, %1, 1;\n\tmad.lo.u32 x, %1, 64, %2;\n\tmad.lo.u32 %0, h, -62, x;\n\t}"
                : "=r"(caddr) : "r"(pos), "r"(sa_stage));
            const int c = lds_s16(caddr);
            if (ZRLS) {
#pragma unroll 1
                while (run >= 16u) { put(zrl & 0xFFFF0000u, zrl & 31u); run -= 16u; }  // rare: keep it small
            }
            const uint32_t a = (uint32_t)abs(c);
            const uint32_t lz = (uint32_t)__clz((int)a);   // 32 - cat
            const uint32_t e = lds_u32(sa_ac_top + run * (AC_STRIDE * 4u) - lz * 4u);
            const uint32_t amp = a ^ ((uint32_t)(c >> 31) >> lz);     // c >= 0: c; c < 0: (c - 1) masked to cat bits
            const uint32_t n = e & 31u;
            put((e & 0xFFFF0000u) | __funnelshift_r(0u, amp, n), n);   // amp << (32 - n), 2 <= n <= 26
        }
    }
    if (nprev != ~63u) put(eob & 0xFFFF0000u, eob & 31u);
    *acc_out = acc;
    return ((sp - sa_slot) / (CB * 4)) * 32u + filled;
}

// warp-wide exclusive scan; *total = sum over the warp
__device__ __forceinline__ uint32_t warp_scan(uint32_t x, int lane, uint32_t *total)
{
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    *total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - x;
}

// Shared memory of one warp.  Two chunks are in flight (see k_huff), each with its own slots.
// The coefficient stage is dead once a chunk's blocks are coded; the stream window and the
// stuffed bytes reuse it.
struct WarpMem {
    uint32_t slot[2][SLOT_W * CB];
    union {
        uint32_t stage[32 * CB];
        struct {
            uint32_t obuf[WIN_W];
            uint8_t sbuf[SBUF_B];
        } w;
    };
    uint32_t tl[CB];
};

// What a warp remembers about a chunk between its phases (lane-private unless noted).
struct ChunkState {
    uint32_t chunk, img;   // uniform
    uint32_t L, o_t;       // this lane's block: code length, bit offset inside the chunk
    int nwt;               // words the block occupies
    uint32_t Lc, ctail;    // uniform: chunk bits, its last 7 bits
    int buf;               // slot / spill buffer
    unsigned long long Pc; // uniform: bits before the chunk (after phase A)
    uint32_t tailin;       // uniform: the 7 bits before the chunk
    uint32_t Ftot;         // uniform: 0xFF bytes the chunk owns
    uint32_t own;          // uniform: output bytes the chunk writes (owned + stuffed zeros + RSTn marker)
    uint32_t marker;       // uniform: 0xD0..0xD7 when the chunk closes a restart interval, else 0
    bool first, last, final_; // uniform: first / last chunk of its bit stream (image or interval); last of the image
    uint32_t ffb;          // 0xFF bytes before this lane's piece of the kept window
    bool kept;             // uniform: phase A left the assembled (single) window in the slot buffer
    bool fault;
    bool skip;             // uniform: the chunk lies past the end of a short last segment: nothing to do
};

// Persistent kernel; every WARP works on its own: it draws chunks of 32 blocks from the ticket
// counter and takes each through three phases with warp-level synchronisation only,
//   W  code the blocks into slots, publish the chunk's bit count              (chain 1)
//   A  look back for the bit offset, assemble + count 0xFF, publish the count  (chain 2)
//   B  look back for the stuffed-byte offset, assemble again, emit
// software-pipelined as  A(j) W(j+1) B(j):  a look-back runs a phase after the value it depends
// on was published by this warp's neighbours in the chain, so it seldom has to wait for them.
//
// RAW (a band of a frame that is tiled over several GPUs, pixo_b200_jpeg_band_entropy_dev): the
// band's bit string starts at an unknown bit of the frame's stream, so nothing that depends on
// byte alignment can happen here.  Phase A writes the UNSTUFFED bytes of the band-local string
// (bit 0 = the band's first bit, the last partial byte zero-filled, no 1-padding) straight from the
// assembled windows, phase B and chain 2 do not exist, and the final chunk reports the string's
// bit count and its last 7 bits.  k_splice_* below turn such a string into scan bytes once the
// bit offset is known.
template <bool RAW>
__global__ void __launch_bounds__(32 * HUFF_WARPS, HUFF_CTAS_PER_SM)
k_huff(const __grid_constant__ EntParams P, const __grid_constant__ HuffDev Tp)
{
    __shared__ HuffDev T;
    __shared__ __align__(16) WarpMem wmem[HUFF_WARPS];
    static_assert(sizeof(((WarpMem *)0)->w) <= sizeof(((WarpMem *)0)->stage), "window + stuffed bytes must fit the stage");

    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (int)(sizeof(HuffDev) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t *>(&T)[i] = reinterpret_cast<const uint32_t *>(&Tp)[i];
    __syncthreads();
    WarpMem &M = wmem[threadIdx.x >> 5];
    uint32_t *const obuf = M.w.obuf;
    uint8_t *const sbuf = M.w.sbuf;
    const uint32_t sa_stage = (uint32_t)__cvta_generic_to_shared(M.stage) + 4u * lane;
    const uint32_t total_chunks = P.nchunks * P.nimages;
    uint32_t spill[2][MAX_W - SLOT_W];  // words beyond SLOT_W (long blocks): local memory

    // ---- W: code the lane's block of chunk `id` into slot buffer `buf`, publish the bit count ----
    auto phase_w = [&](uint32_t id, int buf, ChunkState &C) {
        // chunk-major dispensing: the n images' chains advance side by side
        C.chunk = id / P.nimages;
        C.img = id - C.chunk * P.nimages;
        C.buf = buf;
        C.fault = false;
        // the chunk's place: with a restart interval every interval is its own bit stream
        // (handle_restart, src/jpeg/mod.rs:1423-1445) and is cut into chunks separately
        uint32_t s0, iend, interval = 0;
        uint32_t nblk = P.nblocks;
        size_t y_off = (size_t)C.img * P.y_stride, c_off = (size_t)C.img * P.c_stride;
        bool seg_prev = false;     // block 0 continues the previous segment's DC chain
        if (RAW && P.seg_per_img > 1) {
            const uint32_t ii = C.img / P.seg_per_img, seg = C.img - ii * P.seg_per_img;
            y_off = (size_t)ii * P.y_stride + (size_t)seg * P.seg_y_stride;
            c_off = (size_t)ii * P.c_stride + (size_t)seg * P.seg_c_stride;
            if (seg == P.seg_per_img - 1) nblk = P.nblocks_last;
            seg_prev = seg != 0;
        }
        C.skip = false;
        if (P.rst_blocks) {
            interval = C.chunk / P.cpi;
            const uint32_t sub = C.chunk - interval * P.cpi;
            s0 = interval * P.rst_blocks + sub * CB;
            iend = (uint32_t)min((unsigned long long)(interval + 1) * P.rst_blocks, (unsigned long long)nblk);
            C.first = sub == 0;
        } else {
            s0 = C.chunk * CB;
            iend = nblk;
            C.first = C.chunk == 0;
        }
        if (s0 >= iend) { C.skip = true; return; }   // short last segment: no such chunk
        C.last = s0 + CB >= iend;
        C.final_ = C.last && iend == nblk;
        C.marker = (C.last && !C.final_) ? 0xD0u + (interval & 7u) : 0u;
        const uint32_t s = s0 + lane;
        const int nv = (int)min((uint32_t)CB, iend - s0);
        uint32_t *const slot = M.slot[buf];
        uint32_t L = 0, tail7 = 0;
        int nwt = 0;
        uint32_t M0 = 0, M1 = 0;
        int diff = 0, tbl = 0;
        bool zrl_here = false;
        if (s < iend) {
            const uint32_t m = s / P.bpm;
            const uint32_t k = s - m * P.bpm;
            const int16_t *arr;
            size_t idx;
            int seed;
            if (k < P.y_per_mcu) { arr = P.y + y_off; idx = (size_t)m * P.y_per_mcu + k; tbl = 0; seed = P.dc_seed[0]; }
            else if (k == P.y_per_mcu) { arr = P.cb + c_off; idx = m; tbl = 1; seed = P.dc_seed[1]; }
            else { arr = P.cr + c_off; idx = m; tbl = 1; seed = P.dc_seed[2]; }
            // DC predictors restart with the interval (src/jpeg/mod.rs:1433-1443)
            const bool dc_reset = P.rst_mcus && m % P.rst_mcus == 0 && (k == 0 || k >= P.y_per_mcu);
            // (a later segment's first block follows the previous segment's last one in the same array)
            if (P.dc_seed_dev && idx == 0 && !seg_prev) seed = P.dc_seed_dev[k < P.y_per_mcu ? 0 : (k == P.y_per_mcu ? 1 : 2)];
            const int prev_dc = dc_reset ? 0 : ((idx || seg_prev) ? arr[((long long)idx - 1) * 64] : seed);
            const uint4 *src = reinterpret_cast<const uint4 *>(arr + idx * 64);
            uint32_t e0 = 0, e1 = 0;
            int dc;
            {
                uint32_t n[32];  // the block as K1 wrote it: natural order, two coefficients per word
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 v = __ldg(src + q);
                    n[q * 4] = v.x; n[q * 4 + 1] = v.y; n[q * 4 + 2] = v.z; n[q * 4 + 3] = v.w;
                }
                dc = (int)(int16_t)(n[0] & 0xFFFF);
                // zig-zag reorder (zigzag_reorder, src/jpeg/quantize.rs:107-113) on the way into the
                // stage: word j = coefficients zz(2j), zz(2j+1); all indices are compile-time
                uint32_t w[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int i0 = zz_nat(2 * j), i1 = zz_nat(2 * j + 1);
                    w[j] = __byte_perm(n[i0 >> 1], n[i1 >> 1],
                                       ((i0 & 1) ? 0x0032 : 0x0010) | ((i1 & 1) ? 0x7600 : 0x5400));
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) M.stage[j * CB + lane] = w[j];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    e0 += __vminu2(w[j], 0x00010001u) * (1u << j);       // disjoint bits: + is |
                    e1 += __vminu2(w[16 + j], 0x00010001u) * (1u << j);
                }
            }
            asm volatile("" ::: "memory");  // the stage is read back through ld.shared below
            M0 = interleave16(e0); M1 = interleave16(e1);  // bit i = coefficient i != 0
            diff = (int)(int16_t)(dc - prev_dc);
            // does a non-zero coefficient follow 16 or more zeros (a ZRL symbol)?  position 0 (DC)
            // bounds the first run; a8 bit i = positions i..i+15 all zero
#if HUFF_ZRL_SPLIT
            const unsigned long long N = ((unsigned long long)M1 << 32) | M0 | 1ull;
            unsigned long long a = ~N & (~N >> 1);
            a &= a >> 2; a &= a >> 4; a &= a >> 8;
            zrl_here = ((a << 16) & N) != 0;
#endif
        }
#if HUFF_ZRL_SPLIT
        const bool any_zrl = __any_sync(0xffffffffu, zrl_here);   // uniform: which symbol loop the warp runs
#endif
        (void)zrl_here;
        if (s < iend) {
            const uint32_t sa_ac = (uint32_t)__cvta_generic_to_shared(&T.ac[tbl][0]);
            const uint32_t sa_slot = (uint32_t)__cvta_generic_to_shared(slot) + 4u * lane;
            uint32_t acc;
#if HUFF_ZRL_SPLIT
            if (any_zrl) L = code_block<false, true>(M0, M1, diff, T.dc[tbl], sa_ac, sa_stage, sa_slot, spill[buf], &acc);
            else L = code_block<false, false>(M0, M1, diff, T.dc[tbl], sa_ac, sa_stage, sa_slot, spill[buf], &acc);
#else
            L = code_block<false, true>(M0, M1, diff, T.dc[tbl], sa_ac, sa_stage, sa_slot, spill[buf], &acc);
#endif
            if (L > SLOT_W * 32u)  // long block: run again, keeping the words past the slot in local memory
                L = code_block<true, true>(M0, M1, diff, T.dc[tbl], sa_ac, sa_stage, sa_slot, spill[buf], &acc);
            asm volatile("" ::: "memory");  // slot words were written through st.shared
            const int nw = (int)(L >> 5), filled = (int)(L & 31u);
            nwt = nw;
            if (filled) {
                if (nw < SLOT_W) slot[nw * CB + lane] = acc; else spill[buf][nw - SLOT_W] = acc;
                nwt = nw + 1;
            }
            const uint32_t lastw = nw == 0 ? 0u : (nw - 1 < SLOT_W ? slot[(nw - 1) * CB + lane] : spill[buf][nw - 1 - SLOT_W]);
            tail7 = __funnelshift_rc(acc, lastw, 32 - filled) & 0x7Fu;
        }
        M.tl[lane] = (L << 7) | tail7;
        __syncwarp();  // every lane is done with the stage; tl[] visible
        C.L = L;
        C.nwt = nwt;
        C.o_t = warp_scan(L, lane, &C.Lc);
        uint32_t ctail = 0;
        if (lane == 0) {  // the chunk's last 7 bits (a block has >= 2 bits: at most 4 steps)
            int got = 0;
            for (int k = nv - 1; k >= 0 && got < 7; --k) {
                const uint32_t x = M.tl[k];
                const int take = min((int)(x >> 7), 7 - got);
                ctail |= (x & ((1u << take) - 1u)) << got;
                got += take;
            }
            st_status(P.st_bits + (size_t)C.img * P.nchunks + C.chunk,
                      pack_status(C.first ? ST_PFX : ST_AGG, ctail, C.Lc));
        }
        C.ctail = __shfl_sync(0xffffffffu, ctail, 0);
        C.Pc = 0;
        C.tailin = 0;
        C.Ftot = 0;
        C.own = 0;
        C.ffb = 0;
        C.kept = false;
        __syncwarp();
    };

    // ---- stuffed bytes of one window (this lane's 32 bytes in wv) -> sbuf -> global ----------------
    // a, b: the chunk's owned byte range inside the window; ffb / Fr: 0xFF bytes before this
    // lane's piece / in the whole window; G: output index of the window's first owned byte.
    // mk: RSTn marker byte to append after the window's bytes (0 = none).
    auto emit_window = [&](const ChunkState &C, const uint32_t (&wv)[8], uint32_t ffb, uint32_t Fr, int a, int b,
                           unsigned long long G, uint32_t mk) {
        uint8_t *outp = P.out + (size_t)C.img * P.out_cap;
        const uint32_t nr0 = (uint32_t)max(b - a, 0) + Fr;
        const uint32_t nr = nr0 + (mk ? 2u : 0u);
        const uint32_t shb = (uint32_t)((reinterpret_cast<uintptr_t>(outp) + G) & 15u);
        // Every lane with owned bytes emits its whole 32-byte piece (bytes outside the owned
        // range land outside the part of sbuf that is copied out); sbuf index 16 + shb is
        // the chunk's first owned byte of this window.
        if (32 * lane < b) {
            uint32_t dst = 16u + shb + (uint32_t)(32 * lane) - (uint32_t)a + ffb;
#if HUFF_EMIT_V == 0
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t w = wv[j];
                if (ff_bytes(w) == 0) {
                    if ((dst & 3u) == 0) {
                        *reinterpret_cast<uint32_t *>(sbuf + dst) = __byte_perm(w, 0, 0x0123);
                    } else {
                        sbuf[dst] = (uint8_t)(w >> 24); sbuf[dst + 1] = (uint8_t)(w >> 16);
                        sbuf[dst + 2] = (uint8_t)(w >> 8); sbuf[dst + 3] = (uint8_t)w;
                    }
                    dst += 4;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t byte = (w >> (24 - 8 * i)) & 0xFFu;
                        sbuf[dst++] = (uint8_t)byte;
                        if (byte == 0xFFu) sbuf[dst++] = 0;
                    }
                }
            }
#else
            // A lane without a 0xFF among its 32 bytes (all of them on smooth content, ~7 of 8 on
            // noise) writes them as ALIGNED words whatever its byte offset: word k = the tail of
            // little-endian word k-1 and the head of word k (one funnel shift), plus at most three
            // single bytes at either end - all predicated, the same instruction sequence in every
            // lane.  7 word + <= 6 byte stores instead of 32 byte stores (which were 25 % of the
            // kernel's shared-memory wavefronts).  A lane that holds a 0xFF goes byte by byte.
            uint32_t anyff = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) anyff |= ff_bytes(wv[j]);
            if (anyff == 0) {
                uint32_t m[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = __byte_perm(wv[j], 0, 0x0123);   // memory order
                const uint32_t al = dst & 3u, sh8 = 8u * al;
                uint8_t *base = sbuf + (dst - al);                 // aligned
                uint32_t *wbase = reinterpret_cast<uint32_t *>(base);
#pragma unroll
                for (int k = 1; k < 8; ++k) wbase[k] = __funnelshift_l(m[k - 1], m[k], sh8);
                if (al == 0) {
                    wbase[0] = m[0];
                } else {
                    // head: bytes al..3 of word 0 <- the low 4 - al bytes of m[0]
                    base[3] = (uint8_t)(m[0] >> (24u - sh8));
                    if (al < 3) base[2] = (uint8_t)(m[0] >> (16u - sh8));
                    if (al < 2) base[1] = (uint8_t)m[0];
                    // tail: bytes 0..al-1 of word 8 <- the top al bytes of m[7]
                    base[32] = (uint8_t)(m[7] >> (32u - sh8));
                    if (al > 1) base[33] = (uint8_t)(m[7] >> (40u - sh8));
                    if (al > 2) base[34] = (uint8_t)(m[7] >> 24);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t w = wv[j];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t byte = (w >> (24 - 8 * i)) & 0xFFu;
                        sbuf[dst++] = (uint8_t)byte;
                        if (byte == 0xFFu) sbuf[dst++] = 0;
                    }
                }
            }
#endif
        }
        __syncwarp();
        if (mk) {  // 0xFF 0xDn, not subject to stuffing; after the barrier: the last lane's spare bytes land here too
            if (lane == 0) { sbuf[16u + shb + nr0] = 0xFF; sbuf[16u + shb + nr0 + 1] = (uint8_t)mk; }
            __syncwarp();
        }
        if (G + nr <= P.out_cap) {
            uint8_t *gdst = outp + G - shb;  // 16-byte aligned
            const uint32_t end = shb + nr;
            const uint32_t full_lo = (shb + 15u) >> 4, full_hi = end >> 4;  // whole 16-byte pieces
            const uint8_t *sb = sbuf + 16;
            for (uint32_t c16 = full_lo + lane; c16 < full_hi; c16 += 32)
                *reinterpret_cast<uint4 *>(gdst + c16 * 16) = *reinterpret_cast<const uint4 *>(sb + c16 * 16);
            // ragged head (lanes 0-15) and tail (lanes 16-31), one byte per lane
            const uint32_t hb = (uint32_t)lane < 16u ? shb + lane : max(full_hi, full_lo) * 16u + (lane - 16u);
            const bool in_head = (uint32_t)lane < 16u && hb < min(full_lo * 16u, end);
            const bool in_tail = lane >= 16 && full_hi >= full_lo && hb < end;
            if (in_head || in_tail) gdst[hb] = sb[hb];
        } else if (lane == 0) {
            atomicOr(&P.overflow[C.img], 1u);
        }
        __syncwarp();  // sbuf is rewritten by the next window, or by the next chunk's stage
    };

    // ---- assemble the chunk's stream window by window; count its 0xFF bytes (EMIT: and write) ----
    // gbase: output index of the chunk's first owned byte (EMIT only).  Returns the 0xFF count.
    auto sweep = [&](ChunkState &C, bool emit, unsigned long long gbase) -> uint32_t {
        const uint32_t *const slot = M.slot[C.buf];
        const uint32_t *const spl = spill[C.buf];
        const bool last_chunk = C.last;
        const uint32_t q0 = (uint32_t)C.Pc & 31u;         // bit offset of the chunk inside window word 0
        const uint32_t endbit = q0 + C.Lc;                // window bit index one past the chunk
        const uint32_t padc = (last_chunk && !RAW) ? ((8u - (endbit & 7u)) & 7u) : 0u;   // 1-padding (bits.rs:261-272)
        const uint32_t ob0 = q0 >> 3;                     // owned window bytes [ob0, ob1)
        const uint32_t ob1 = RAW ? ((last_chunk ? endbit + 7u : endbit) >> 3) : (endbit >> 3) + (padc ? 1u : 0u);
        const int nrounds = max(1, (int)((ob1 + WIN_B - 1) / WIN_B));
        // per-lane constants of the funnel-shifted copy
        const uint32_t D = q0 + C.o_t;
        const int d0 = (int)(D >> 5), sh = (int)(D & 31u);
        const int nd = C.L ? (int)((sh + C.L + 31u) >> 5) : 0;   // destination words
        const int nwt = C.nwt;
        uint32_t Fsum = 0;
#pragma unroll 1
        for (int r = 0; r < nrounds; ++r) {
            for (int i = lane; i < WIN_W / 4; i += 32) reinterpret_cast<uint4 *>(obuf)[i] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            {
                const int wlo = r * WIN_W;
                const int kb = max(0, wlo - d0), ke = min(nd, wlo + WIN_W - d0);
                uint32_t prev = 0;
                if (nwt <= SLOT_W) {  // the usual case: every word is in the slot
                    if (kb > 0 && kb <= nwt) prev = slot[(kb - 1) * CB + lane];
                    for (int k = kb; k < ke; ++k) {
                        const uint32_t cur = k < nwt ? slot[k * CB + lane] : 0u;
                        const uint32_t v = __funnelshift_r(cur, prev, sh);
                        uint32_t *dst = obuf + (d0 + k - wlo);
                        if (k == 0 || k == nd - 1) atomicOr(dst, v); else *dst = v;
                        prev = cur;
                    }
                } else {
                    if (kb > 0 && kb <= nwt) prev = (kb - 1) < SLOT_W ? slot[(kb - 1) * CB + lane] : spl[kb - 1 - SLOT_W];
                    for (int k = kb; k < ke; ++k) {
                        uint32_t cur = 0;
                        if (k < nwt) cur = k < SLOT_W ? slot[k * CB + lane] : spl[k - SLOT_W];
                        const uint32_t v = __funnelshift_r(cur, prev, sh);
                        uint32_t *dst = obuf + (d0 + k - wlo);
                        if (k == 0 || k == nd - 1) atomicOr(dst, v); else *dst = v;
                        prev = cur;
                    }
                }
                if (lane == 0) {
                    const uint32_t q = q0 & 7u;  // inherited bits of the straddling first byte
                    if (r == 0 && q) atomicOr(&obuf[0], (C.tailin & ((1u << q) - 1u)) << (32u - q0));
                    const int pw = (int)(endbit >> 5) - wlo;
                    if (padc && pw >= 0 && pw < WIN_W)
                        atomicOr(&obuf[pw], ((1u << padc) - 1u) << (32u - (endbit & 31u) - padc));
                }
            }
            __syncwarp();
            // Count the 0xFF bytes in this lane's 32 window bytes.  No bounds: a window byte the
            // chunk does not own is either untouched (0) or the unfinished last byte, whose low
            // bits are still 0 - never 0xFF.
            const int wb0 = r * WIN_B;
            const int a = max((int)ob0 - wb0, 0), b = min((int)ob1 - wb0, WIN_B);
            uint32_t wv[8];
            {
                const uint4 x = reinterpret_cast<const uint4 *>(obuf)[2 * lane];
                const uint4 y = reinterpret_cast<const uint4 *>(obuf)[2 * lane + 1];
                wv[0] = x.x; wv[1] = x.y; wv[2] = x.z; wv[3] = x.w; wv[4] = y.x; wv[5] = y.y; wv[6] = y.z; wv[7] = y.w;
            }
            if (RAW) {
                // window byte i is byte (Pc >> 5) * 4 + wb0 + i of the band's raw string
                uint8_t *outp = P.out + (size_t)C.img * P.out_cap;
                const unsigned long long wbase = (C.Pc >> 5) * 4ull + (unsigned long long)wb0;
                if (wbase + (unsigned long long)max(b, 0) <= P.out_cap) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int i0 = 32 * lane + 4 * j;
                        if (i0 >= a && i0 + 4 <= b) {
                            *reinterpret_cast<uint32_t *>(outp + wbase + i0) = __byte_perm(wv[j], 0, 0x0123);
                        } else if (i0 + 4 > a && i0 < b) {
#pragma unroll
                            for (int t = 0; t < 4; ++t)
                                if (i0 + t >= a && i0 + t < b) outp[wbase + i0 + t] = (uint8_t)(wv[j] >> (24 - 8 * t));
                        }
                    }
                } else if (lane == 0) {
                    atomicOr(&P.overflow[C.img], 1u);
                }
                __syncwarp();
                continue;
            }
            uint32_t cnt = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) cnt += __popc(ff_bytes(wv[j]));
            uint32_t Fr;
            const uint32_t ffb = warp_scan(cnt, lane, &Fr);
            if (!emit && nrounds == 1) {
                // the usual case: keep the assembled window (in the chunk's slot buffer, which is
                // not needed any more) so that phase B emits it without assembling again
                uint4 *keep = reinterpret_cast<uint4 *>(M.slot[C.buf]);
                keep[2 * lane] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                keep[2 * lane + 1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
                C.ffb = ffb;
                C.kept = true;
            }
            if (emit) {
                emit_window(C, wv, ffb, Fr, a, b, gbase + (r == 0 ? 0u : (uint32_t)(wb0 - (int)ob0)) + Fsum,
                            r == nrounds - 1 ? C.marker : 0u);
            }
            Fsum += Fr;
            __syncwarp();  // obuf / sbuf are rewritten by the next round, or by the next chunk's stage
        }
        if (emit && lane == 0 && C.final_) {
            const unsigned long long total = gbase + (ob1 - ob0) + Fsum;
            P.out_len[C.img] = total;
            if (total > P.out_cap) atomicOr(&P.overflow[C.img], 1u);
        }
        if (!emit) C.own = (ob1 - ob0) + Fsum + (C.marker ? 2u : 0u);
        return Fsum;
    };

    // ---- A: bit offset from chain 1, then the chunk's 0xFF count into chain 2 ------------------------
    auto phase_a = [&](ChunkState &C) {
        if (C.skip) return;
        unsigned long long *st1 = P.st_bits + (size_t)C.img * P.nchunks;
        unsigned long long *st2 = P.st_ff + (size_t)C.img * P.nchunks;
        if (!C.first) {
            const unsigned long long lb = look_back(st1, (int)C.chunk, lane);
            C.Pc = lb & ST_VAL;
            C.tailin = (uint32_t)(lb >> 55) & 0x7Fu;
            C.fault |= (lb >> 62) != 0;
            if (lane == 0) st_status(st1 + C.chunk, pack_status(ST_PFX, C.ctail, C.Pc + C.Lc));
        }
        C.Ftot = sweep(C, false, 0);
        if (RAW) {
            if (lane == 0 && C.final_) {
                P.out_len[C.img] = C.Pc + C.Lc;       // BITS
                P.out_tail[C.img] = C.ctail;
                if (((C.Pc + C.Lc + 7) >> 3) > P.out_cap) atomicOr(&P.overflow[C.img], 1u);
            }
            if (lane == 0 && C.fault) atomicOr(&P.overflow[C.img], 2u);
            return;
        }
        if (lane == 0) st_status(st2 + C.chunk, pack_status(C.chunk == 0 ? ST_PFX : ST_AGG, 0, C.own));
    };
    // ---- B: stuffed-byte offset from chain 2, then the bytes ------------------------------------------
    auto phase_b = [&](ChunkState &C) {
        if (RAW || C.skip) return;
        unsigned long long *st2 = P.st_ff + (size_t)C.img * P.nchunks;
        unsigned long long ffx = 0;
        if (C.chunk) {
            const unsigned long long lb = look_back(st2, (int)C.chunk, lane);
            ffx = lb & ST_VAL;
            C.fault |= (lb >> 62) != 0;
            if (lane == 0) st_status(st2 + C.chunk, pack_status(ST_PFX, 0, ffx + C.own));
        }
        const unsigned long long gbase = ffx;  // chain 2 counts every byte written before this chunk
        if (C.kept) {
            const uint32_t q0 = (uint32_t)C.Pc & 31u, endbit = q0 + C.Lc;
            const uint32_t padc = C.last ? ((8u - (endbit & 7u)) & 7u) : 0u;
            const uint32_t ob0 = q0 >> 3, ob1 = (endbit >> 3) + (padc ? 1u : 0u);
            const uint4 *keep = reinterpret_cast<const uint4 *>(M.slot[C.buf]);
            const uint4 x = keep[2 * lane], y = keep[2 * lane + 1];
            const uint32_t wv[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
            emit_window(C, wv, C.ffb, C.Ftot, (int)ob0, (int)ob1, gbase, C.marker);
            if (lane == 0 && C.final_) {
                const unsigned long long total = gbase + (ob1 - ob0) + C.Ftot;
                P.out_len[C.img] = total;
                if (total > P.out_cap) atomicOr(&P.overflow[C.img], 1u);
            }
        } else {
            sweep(C, true, gbase);
        }
        if (lane == 0 && C.fault) atomicOr(&P.overflow[C.img], 2u);
    };

    ChunkState cur, pend;
    bool have_pend = false;
    int buf = 0;
    for (;;) {
        uint32_t id = 0;
        if (lane == 0) id = atomicAdd(P.ticket, 1u);
        id = __shfl_sync(0xffffffffu, id, 0);
        const bool have = id < total_chunks;
        if (have_pend) phase_a(pend);
        if (have) phase_w(id, buf, cur);
        if (have_pend) phase_b(pend);
        if (!have) break;
        pend = cur;
        have_pend = true;
        buf ^= 1;
    }
}

}  // namespace

// Device scratch layout for n images (all sizes in bytes, 256-aligned)
struct EntropyPlan {
    size_t nchunks;
    size_t off_st1, off_st2, off_ticket, off_ovf, zero_bytes, off_outlen, off_tail, total;
};

static size_t a256(size_t v) { return (v + 255) / 256 * 256; }

static EntropyPlan plan_entropy(uint32_t n, uint64_t nblocks, uint64_t rst_blocks)
{
    EntropyPlan p;
    p.nchunks = (size_t)((nblocks + CB - 1) / CB);
    if (rst_blocks && rst_blocks < nblocks) {  // every interval is chunked on its own
        const uint64_t n_int = (nblocks + rst_blocks - 1) / rst_blocks, cpi = (rst_blocks + CB - 1) / CB;
        const uint64_t last = nblocks - (n_int - 1) * rst_blocks;
        p.nchunks = (size_t)((n_int - 1) * cpi + (last + CB - 1) / CB);
    }
    size_t o = 0;
    p.off_st1 = o; o += a256((size_t)n * p.nchunks * 8);
    p.off_st2 = o; o += a256((size_t)n * p.nchunks * 8);
    p.off_ticket = o; o += 256;
    p.off_ovf = o; o += a256((size_t)n * 4);
    p.zero_bytes = o;  // everything up to here is cleared per launch
    p.off_outlen = o; o += a256((size_t)n * 8);
    p.off_tail = o; o += a256((size_t)n * 8);
    p.total = o;
    return p;
}

size_t entropy_scratch_bytes(uint32_t n, const FrameGeometry &g, uint32_t restart_interval)
{
    const uint64_t bpm = g.y_per_mcu + (g.has_chroma ? 2 : 0);
    return plan_entropy(n, g.ny + 2 * g.nc, (uint64_t)restart_interval * bpm).total;
}

// ---- splicing a band's raw bit string into the frame's scan ------------------------------------
// T = tail_in (s bits, the frame's stream bits that precede the band inside its first byte) ++ B
// (the band's raw string, nbits).  The band owns T's whole bytes; the frame's last band also owns
// the final partial byte, padded with 1s (BitWriterMsb::flush, src/bits.rs:261-272).  Every 0xFF is
// followed by 0x00 (flush_byte_with_stuffing, :245-259).  Two small kernels: per-tile 0xFF counts,
// then every tile sums the counts before it and writes its stuffed bytes.
namespace {

constexpr int SPL_TILE = 4096;  // T bytes per CTA
constexpr int SPL_THREADS = 256;

struct SpliceParams {
    const uint8_t *raw;
    unsigned long long nbits;
    uint32_t s, tail_in, last;
    unsigned long long nbytes;   // T bytes this band emits before stuffing (incl. the padded one)
    uint32_t *cnt;               // [ntiles]
    uint8_t *out;
    unsigned long long out_cap;
    unsigned long long *out_len;
    uint32_t *overflow;
};

// byte m of T (with the final partial byte 1-padded when this is the frame's last band)
__device__ __forceinline__ uint32_t splice_byte(const SpliceParams &P, unsigned long long m)
{
    const unsigned long long rb = (P.nbits + 7) >> 3;   // raw bytes that exist
    const uint32_t cur = m < rb ? P.raw[m] : 0u;
    const uint32_t prev = m ? P.raw[m - 1] : P.tail_in;
    uint32_t v = (((prev << 8) | cur) >> P.s) & 0xFFu;
    const unsigned long long tbits = P.s + P.nbits;
    if (m == (tbits >> 3)) v |= 0xFFu >> (uint32_t)(tbits & 7);   // only reached when last && tbits % 8 != 0
    return v;
}

// The same sixteen bytes at once.  T is the raw string shifted right by `phase` bits behind the inherited
// bits, so four big-endian words of T are four funnel shifts over the raw words (and the raw byte before
// them).  Only for a thread whose piece lies wholly inside the raw string, short of its last byte (no
// padding, no missing bytes), at a 16-byte aligned address; the ragged ends take the byte-wise path.
__device__ __forceinline__ bool t_words16(const uint8_t *raw, unsigned long long rb, uint32_t phase, uint32_t tail_in,
                                          unsigned long long m0, uint32_t (&O)[4])
{
    if (m0 + 16 >= rb || (reinterpret_cast<uintptr_t>(raw + m0) & 15)) return false;
    const uint4 L = *reinterpret_cast<const uint4 *>(raw + m0);
    const uint32_t pb = m0 ? raw[m0 - 1] : tail_in;
    const uint32_t A0 = __byte_perm(L.x, 0, 0x0123), A1 = __byte_perm(L.y, 0, 0x0123);
    const uint32_t A2 = __byte_perm(L.z, 0, 0x0123), A3 = __byte_perm(L.w, 0, 0x0123);
    O[0] = __funnelshift_r(A0, pb, phase);
    O[1] = __funnelshift_r(A1, A0, phase);
    O[2] = __funnelshift_r(A2, A1, phase);
    O[3] = __funnelshift_r(A3, A2, phase);
    return true;
}
__device__ __forceinline__ uint32_t ff_count16(const uint32_t (&O)[4])
{
    return __popc(ff_bytes(O[0])) + __popc(ff_bytes(O[1])) + __popc(ff_bytes(O[2])) + __popc(ff_bytes(O[3]));
}
// sixteen stuffing-free bytes (four big-endian words) -> sb[dst .. dst + 16) at any byte alignment:
// three aligned word stores plus at most three single bytes at either end
__device__ __forceinline__ void put16(uint8_t *sb, uint32_t dst, const uint32_t (&O)[4])
{
    uint32_t m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = __byte_perm(O[k], 0, 0x0123);   // memory order
    const uint32_t al = dst & 3u, sh8 = 8u * al;
    uint8_t *base = sb + (dst - al);
    uint32_t *wb = reinterpret_cast<uint32_t *>(base);
#pragma unroll
    for (int k = 1; k < 4; ++k) wb[k] = __funnelshift_l(m[k - 1], m[k], sh8);
    if (al == 0) {
        wb[0] = m[0];
    } else {
        base[3] = (uint8_t)(m[0] >> (24u - sh8));
        if (al < 3) base[2] = (uint8_t)(m[0] >> (16u - sh8));
        if (al < 2) base[1] = (uint8_t)m[0];
        base[16] = (uint8_t)(m[3] >> (32u - sh8));
        if (al > 1) base[17] = (uint8_t)(m[3] >> (40u - sh8));
        if (al > 2) base[18] = (uint8_t)(m[3] >> 24);
    }
}

__global__ void __launch_bounds__(SPL_THREADS) k_splice_count(const __grid_constant__ SpliceParams P)
{
    __shared__ uint32_t red[SPL_THREADS / 32];
    const unsigned long long m0 = (unsigned long long)blockIdx.x * SPL_TILE + threadIdx.x * 16;
    uint32_t c = 0, O[4];
    if (t_words16(P.raw, (P.nbits + 7) >> 3, P.s, P.tail_in, m0, O)) {
        c = ff_count16(O);
    } else {
        for (int i = 0; i < 16; ++i)
            if (m0 + i < P.nbytes) c += splice_byte(P, m0 + i) == 0xFFu;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < SPL_THREADS / 32; ++i) t += red[i];
        P.cnt[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(SPL_THREADS) k_splice_emit(const __grid_constant__ SpliceParams P)
{
    __shared__ unsigned long long s_before;
    __shared__ uint32_t wsum[SPL_THREADS / 32];
    __shared__ __align__(16) uint8_t sb[2 * SPL_TILE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // 0xFF bytes in the tiles before this one
    unsigned long long before = 0;
    for (uint32_t i = tid; i < blockIdx.x; i += SPL_THREADS) before += P.cnt[i];
    for (int o = 16; o; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (tid == 0) s_before = 0;
    __syncthreads();
    if (lane == 0 && before) atomicAdd(&s_before, before);
    const unsigned long long m0 = (unsigned long long)blockIdx.x * SPL_TILE + tid * 16;
    uint32_t v[16], c = 0, O[4];
    const bool fast = t_words16(P.raw, (P.nbits + 7) >> 3, P.s, P.tail_in, m0, O);
    if (fast) {
        c = ff_count16(O);
    } else {
        for (int i = 0; i < 16; ++i) {
            v[i] = m0 + i < P.nbytes ? splice_byte(P, m0 + i) : 0x100u;
            c += v[i] == 0xFFu;
        }
    }
    // exclusive scan of the per-thread 0xFF counts over the CTA
    uint32_t inc = c;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
    for (int i = 0; i < SPL_THREADS / 32; ++i) { if (i < warp) woff += wsum[i]; total += wsum[i]; }
    uint32_t dst = tid * 16 + woff + inc - c;
    if (fast && c == 0) {
        put16(sb, dst, O);
    } else {
        if (fast)
            for (int i = 0; i < 16; ++i) v[i] = (O[i >> 2] >> (24 - 8 * (i & 3))) & 0xFFu;
        for (int i = 0; i < 16; ++i) {
            if (v[i] > 0xFFu) break;
            sb[dst++] = (uint8_t)v[i];
            if (v[i] == 0xFFu) sb[dst++] = 0;
        }
    }
    __syncthreads();
    const unsigned long long tile_first = (unsigned long long)blockIdx.x * SPL_TILE;
    const unsigned long long tile_n = P.nbytes > tile_first ? min((unsigned long long)SPL_TILE, P.nbytes - tile_first) : 0ull;
    const unsigned long long g0 = tile_first + s_before;
    const uint32_t nout = (uint32_t)tile_n + total;
    if (g0 + nout <= P.out_cap) {
        for (uint32_t i = tid; i < nout; i += SPL_THREADS) P.out[g0 + i] = sb[i];
    } else if (tid == 0) {
        atomicOr(P.overflow, 1u);
    }
    if (tid == 0 && blockIdx.x == gridDim.x - 1) *P.out_len = g0 + nout;
}

}  // namespace

// ---- segmented coding: many short chains instead of one long one ------------------------------------
// One image (or a handful) gives the single-pass kernel only one look-back chain per image: with
// ~3500 warps in flight on the same chain a chunk has to look back over thousands of predecessors
// (a 16 384^2 frame took 1.95 ms in k_huff against 0.39 ms for its transform).  Instead the image is
// cut into S runs of whole MCUs, each coded by k_huff<RAW> as a bit string of its own - S independent,
// short chains advancing side by side - and four small kernels splice the strings: per-segment bit
// offsets (k_seg_prefix), 0xFF counts per 4 KB tile of the shifted stream (k_seg_count), their prefix
// (k_seg_scan), and the stuffed bytes (k_seg_emit).  No host round trip in between.  The same
// machinery splices a band of a frame tiled over several GPUs (base bit offset and inherited bits
// come from the other ranks).
namespace {

struct SegRec {
    unsigned long long nbits, byte_off, nbytes;   // byte_off: T-bytes of the image's earlier segments
    uint32_t phase, tail_in, tile_off, last;
};

struct SegParams {
    const uint8_t *raw;             // [n * S][raw_cap]
    unsigned long long raw_cap;
    const unsigned long long *bits; // [n * S] from k_huff<RAW>
    const unsigned long long *tails;
    uint32_t S, max_tiles;
    unsigned long long base_bit;    // bits of the stream before segment 0 (a band of a tiled frame; 0 otherwise)
    uint32_t base_tail, last_band;  // the stream's last base_bit % 8 bits; 1: the last segment ends the stream (1-pad)
    const unsigned long long *base_dev;   // {base_bit, the previous band's last 7 bits, last_band} in device memory, or null
    SegRec *rec;                    // [n * S]
    uint32_t *ntiles;               // [n]
    uint32_t *cnt;                  // [n][max_tiles] 0xFF counts, then their exclusive prefix
    uint8_t *out;                   // [n][out_cap]
    unsigned long long out_cap;
    unsigned long long *out_len;    // [n]
    uint32_t *overflow;             // [n]
    const uint32_t *raw_overflow;   // [n * S] flags of the coding kernel
};

__device__ __forceinline__ uint32_t seg_byte(const uint8_t *raw, const SegRec &r, unsigned long long m)
{
    const unsigned long long rb = (r.nbits + 7) >> 3;
    const uint32_t cur = m < rb ? raw[m] : 0u;
    const uint32_t prev = m ? raw[m - 1] : r.tail_in;
    uint32_t v = (((prev << 8) | cur) >> r.phase) & 0xFFu;
    const unsigned long long tbits = r.phase + r.nbits;
    if (m == (tbits >> 3)) v |= 0xFFu >> (uint32_t)(tbits & 7);   // reached only by the padded last byte
    return v;
}

// a band's totals for the exchange with the other ranks: {bits, last 7 bits}, and its flags
__global__ void k_band_totals(const unsigned long long *bits, const unsigned long long *tails, const uint32_t *ovf,
                              uint32_t S, unsigned long long *out2, uint32_t *flags)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long total = 0, tail = 0;
    uint32_t f = 0;
    for (uint32_t q = 0; q < S; ++q) {
        total += bits[q];
        if (bits[q]) tail = tails[q];
        f |= ovf[q];
    }
    out2[0] = total;
    out2[1] = tail & 0x7Full;
    if (f) atomicOr(flags, f);
}

// exclusive prefix over the CTA's SPL_THREADS values (every thread calls; `sh` holds one word per warp)
template <typename T>
__device__ __forceinline__ T cta_exclusive_scan(T x, T *sh, T *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = x;
    for (int o = 1; o < 32; o <<= 1) {
        const T v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    __syncthreads();              // sh may still be read from a previous scan
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    T before = 0, all = 0;
    for (int w = 0; w < SPL_THREADS / 32; ++w) { const T v = sh[w]; if (w < warp) before += v; all += v; }
    *total = all;
    return before + inc - x;
}

// Per image: every segment's record (bit phase, inherited bits, byte and tile offsets) from the
// segments' bit counts.  One CTA per image, thread s = segment s (S <= SEG_MAX == SPL_THREADS): three
// prefix sums (bits, bytes, tiles) and a "nearest earlier non-empty segment" scan for the inherited bits.
__global__ void __launch_bounds__(SPL_THREADS) k_seg_prefix(const __grid_constant__ SegParams P)
{
    __shared__ unsigned long long sh64[SPL_THREADS / 32];
    __shared__ uint32_t sh32[SPL_THREADS / 32];
    __shared__ int shmax[SPL_THREADS / 32];
    const uint32_t i = blockIdx.x, s = threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long base = P.base_bit;
    uint32_t base_tail = P.base_tail, last_band = P.last_band;
    if (P.base_dev) { base = P.base_dev[0]; base_tail = (uint32_t)P.base_dev[1]; last_band = (uint32_t)P.base_dev[2]; }
    const bool live = s < P.S;
    const uint32_t q = i * P.S + s;
    const unsigned long long nbits = live ? P.bits[q] : 0ull;
    const uint32_t bad_here = (live && P.raw_overflow) ? P.raw_overflow[q] : 0u;
    unsigned long long total_bits;
    const unsigned long long start = base + cta_exclusive_scan<unsigned long long>(nbits, sh64, &total_bits);
    SegRec r;
    r.nbits = nbits;
    r.phase = (uint32_t)(start & 7);
    r.last = (live && s == P.S - 1 && last_band) ? 1u : 0u;
    const unsigned long long tbits = r.phase + r.nbits;
    r.nbytes = live ? (tbits >> 3) + ((r.last && (tbits & 7)) ? 1 : 0) : 0ull;
    unsigned long long total_bytes;
    r.byte_off = cta_exclusive_scan<unsigned long long>(r.nbytes, sh64, &total_bytes);
    uint32_t total_tiles;
    r.tile_off = cta_exclusive_scan<uint32_t>((uint32_t)((r.nbytes + SPL_TILE - 1) / SPL_TILE), sh32, &total_tiles);
    // the bits inherited in the first byte: the last 7 bits of the nearest earlier segment that has any
    int near = (live && nbits) ? (int)s : -1;      // inclusive running maximum, then shifted by one
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, near, o);
        if (lane >= o) near = max(near, v);
    }
    if (lane == 31) shmax[warp] = near;
    __syncthreads();
    int prev = __shfl_up_sync(0xffffffffu, near, 1);
    if (lane == 0) prev = -1;
    for (int w = 0; w < warp; ++w) prev = max(prev, shmax[w]);
    const uint32_t tail_prev = prev >= 0 ? (uint32_t)P.tails[i * P.S + prev] : base_tail;
    r.tail_in = tail_prev & ((1u << r.phase) - 1u);
    if (live) P.rec[q] = r;
    const uint32_t bad = (uint32_t)__syncthreads_or((int)bad_here);
    if (s == 0) {
        P.ntiles[i] = total_tiles;
        // a raw segment that did not fit (or a faulted chain): the caller codes the image again unsegmented
        if (bad || total_tiles > P.max_tiles) { P.overflow[i] = 4u | (bad & 2u); P.ntiles[i] = 0; P.out_len[i] = 0; }
        else if (total_tiles == 0) P.out_len[i] = 0;
    }
}

// A CTA of the splice kernels works through SPL_TPC consecutive tiles of one image (one tile per CTA
// left the kernels latency-bound: 17 000 CTAs for a 70 MB scan, each behind a chain of dependent loads).
// The image's segment records are read once into shared memory.
constexpr int SPL_TPC = 8;
constexpr int SEG_MAX = 256;
static_assert(SEG_MAX == SPL_THREADS, "thread s of a splice CTA looks at segment record s");

__device__ __forceinline__ void load_seg_table(const SegParams &P, uint32_t i, SegRec *tab)
{
    if (threadIdx.x < P.S) tab[threadIdx.x] = P.rec[i * P.S + threadIdx.x];
    __syncthreads();
}
// the segment tile t belongs to: the last one whose first tile is <= t (tile_off is non-decreasing and
// tab[0].tile_off == 0).  Whole CTA, one barrier: thread s looks at record s.
__device__ __forceinline__ uint32_t seg_of_tile(const SegRec *tab, uint32_t S, uint32_t t)
{
    return (uint32_t)__syncthreads_count(threadIdx.x < S && tab[threadIdx.x].tile_off <= t) - 1u;
}

// nout staged bytes -> outp[g0 ..).  The stage was filled from index shb = (address of outp + g0) & 15 on,
// so stage and destination agree modulo 16: whole 16-byte pieces, single bytes at the two ragged ends.
__device__ __forceinline__ void copy_out16(uint8_t *outp, unsigned long long g0, const uint8_t *sb, uint32_t shb,
                                           uint32_t nout, int tid)
{
    uint8_t *gdst = outp + g0 - shb;   // 16-byte aligned
    const uint32_t end = shb + nout;
    const uint32_t full_lo = (shb + 15u) >> 4, full_hi = end >> 4;
    for (uint32_t c16 = full_lo + tid; c16 < full_hi; c16 += SPL_THREADS)
        *reinterpret_cast<uint4 *>(gdst + c16 * 16) = *reinterpret_cast<const uint4 *>(sb + c16 * 16);
    if (tid < 16) {                       // ragged head
        const uint32_t hb = shb + tid;
        if (hb < min(full_lo * 16u, end)) gdst[hb] = sb[hb];
    } else if (tid < 32) {                // ragged tail
        const uint32_t tb = max(full_hi, full_lo) * 16u + (tid - 16);
        if (full_hi >= full_lo && tb < end) gdst[tb] = sb[tb];
    }
}

__global__ void __launch_bounds__(SPL_THREADS) k_seg_count(const __grid_constant__ SegParams P)
{
    __shared__ uint32_t red[SPL_TPC][SPL_THREADS / 32];
    __shared__ SegRec tab[SEG_MAX];
    const uint32_t i = blockIdx.y, t0 = blockIdx.x * SPL_TPC, nt = P.ntiles[i];
    if (t0 >= nt) return;
    load_seg_table(P, i, tab);
    const uint32_t t1 = min(nt, t0 + SPL_TPC);
    for (uint32_t t = t0; t < t1; ++t) {
        const uint32_t s = seg_of_tile(tab, P.S, t);
        const SegRec &r = tab[s];
        const uint8_t *raw = P.raw + (size_t)(i * P.S + s) * P.raw_cap;
        const unsigned long long m0 = (unsigned long long)(t - r.tile_off) * SPL_TILE + threadIdx.x * 16;
        uint32_t c = 0, O[4];
        if (t_words16(raw, (r.nbits + 7) >> 3, r.phase, r.tail_in, m0, O)) {
            c = ff_count16(O);
        } else {
            for (int k = 0; k < 16; ++k)
                if (m0 + k < r.nbytes) c += seg_byte(raw, r, m0 + k) == 0xFFu;
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if ((threadIdx.x & 31) == 0) red[t - t0][threadIdx.x >> 5] = c;
    }
    __syncthreads();
    if (threadIdx.x < t1 - t0) {
        uint32_t tot = 0;
        for (int k = 0; k < SPL_THREADS / 32; ++k) tot += red[threadIdx.x][k];
        P.cnt[(size_t)i * P.max_tiles + t0 + threadIdx.x] = tot;
    }
}

// exclusive prefix of an image's tile counts, in place (one CTA per image)
__global__ void __launch_bounds__(1024) k_seg_scan(const __grid_constant__ SegParams P)
{
    __shared__ uint32_t wtot[32];
    const uint32_t i = blockIdx.x, n = P.ntiles[i];
    uint32_t *c = P.cnt + (size_t)i * P.max_tiles;
    const uint32_t per = (n + 1023) / 1024, lo = threadIdx.x * per, hi = min(n, lo + per);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t sum = 0;
    for (uint32_t k = lo; k < hi; ++k) sum += c[k];
    uint32_t inc = sum;                        // inclusive scan of the 1024 partial sums: warp, then warp totals
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wtot[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += v;
        }
        wtot[lane] = w;
    }
    __syncthreads();
    uint32_t run = inc - sum + (warp ? wtot[warp - 1] : 0u);
    for (uint32_t k = lo; k < hi; ++k) { const uint32_t v = c[k]; c[k] = run; run += v; }
}

__global__ void __launch_bounds__(SPL_THREADS) k_seg_emit(const __grid_constant__ SegParams P)
{
    __shared__ uint32_t wsum[SPL_THREADS / 32];
    __shared__ __align__(16) uint8_t sb[2 * SPL_TILE + 32];
    __shared__ SegRec tab[SEG_MAX];
    const uint32_t i = blockIdx.y, t0 = blockIdx.x * SPL_TPC, nt = P.ntiles[i];
    if (t0 >= nt) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_seg_table(P, i, tab);
    uint8_t *outp = P.out + (size_t)i * P.out_cap;
    const uint32_t t1 = min(nt, t0 + SPL_TPC);
    for (uint32_t t = t0; t < t1; ++t) {
        const uint32_t s = seg_of_tile(tab, P.S, t);
        const SegRec &r = tab[s];
        const uint8_t *raw = P.raw + (size_t)(i * P.S + s) * P.raw_cap;
        const unsigned long long tile_first = (unsigned long long)(t - r.tile_off) * SPL_TILE;
        const unsigned long long m0 = tile_first + tid * 16;
        const unsigned long long g0 = r.byte_off + tile_first + P.cnt[(size_t)i * P.max_tiles + t];
        const uint32_t shb = (uint32_t)((reinterpret_cast<uintptr_t>(outp) + g0) & 15u);
        uint32_t v[16], c = 0, O[4];
        const bool fast = t_words16(raw, (r.nbits + 7) >> 3, r.phase, r.tail_in, m0, O);
        if (fast) {
            c = ff_count16(O);
        } else {
            for (int k = 0; k < 16; ++k) {
                v[k] = m0 + k < r.nbytes ? seg_byte(raw, r, m0 + k) : 0x100u;
                c += v[k] == 0xFFu;
            }
        }
        uint32_t inc = c;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t nb = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += nb;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t woff = 0, total = 0;
        for (int k = 0; k < SPL_THREADS / 32; ++k) { if (k < warp) woff += wsum[k]; total += wsum[k]; }
        uint32_t dst = shb + tid * 16 + woff + inc - c;
        if (fast && c == 0) {
            put16(sb, dst, O);
        } else {
            if (fast)
                for (int k = 0; k < 16; ++k) v[k] = (O[k >> 2] >> (24 - 8 * (k & 3))) & 0xFFu;
            for (int k = 0; k < 16; ++k) {
                if (v[k] > 0xFFu) break;
                sb[dst++] = (uint8_t)v[k];
                if (v[k] == 0xFFu) sb[dst++] = 0;
            }
        }
        __syncthreads();
        const unsigned long long tile_n = min((unsigned long long)SPL_TILE, r.nbytes - tile_first);
        const uint32_t nout = (uint32_t)tile_n + total;
        if (g0 + nout <= P.out_cap) copy_out16(outp, g0, sb, shb, nout, tid);
        else if (tid == 0) atomicOr(&P.overflow[i], 1u);
        if (tid == 0 && t == nt - 1) P.out_len[i] = g0 + nout;   // the size needed, also when it did not fit
        __syncthreads();   // the stage and wsum are rewritten by the next tile
    }
}

}  // namespace

size_t splice_scratch_bytes(uint64_t nbits) { return a256(((nbits + 16) / 8 / SPL_TILE + 2) * 4) + 256; }

// d_scratch: splice_scratch_bytes(nbits).  *d_out_len / *d_overflow point into it.
int launch_splice(pixo_b200_ctx *ctx, const uint8_t *d_raw, uint64_t nbits, uint32_t phase, uint32_t tail_in,
                  bool last, uint8_t *d_scratch, uint8_t *d_out, uint64_t out_cap, uint64_t **d_out_len,
                  uint32_t **d_overflow)
{
    SpliceParams P;
    P.raw = d_raw; P.nbits = nbits; P.s = phase & 7u; P.tail_in = tail_in & ((1u << (phase & 7u)) - 1u);
    P.last = last ? 1u : 0u;
    const uint64_t tbits = (uint64_t)P.s + nbits;
    P.nbytes = (tbits >> 3) + ((last && (tbits & 7)) ? 1 : 0);
    const unsigned ntiles = (unsigned)((P.nbytes + SPL_TILE - 1) / SPL_TILE);
    P.out_len = reinterpret_cast<unsigned long long *>(d_scratch);
    P.overflow = reinterpret_cast<uint32_t *>(d_scratch + 8);
    P.cnt = reinterpret_cast<uint32_t *>(d_scratch + 256);
    P.out = d_out; P.out_cap = out_cap;
    *d_out_len = reinterpret_cast<uint64_t *>(P.out_len);
    *d_overflow = P.overflow;
    PIXO_CUDA(ctx, cudaMemsetAsync(d_scratch, 0, 256, ctx->stream));
    if (ntiles == 0) return 0;   // nothing owned: length 0
    k_splice_count<<<ntiles, SPL_THREADS, 0, ctx->stream>>>(P);
    k_splice_emit<<<ntiles, SPL_THREADS, 0, ctx->stream>>>(P);
    ctx->launches += 2;
    PIXO_CUDA(ctx, cudaGetLastError());
    return 0;
}

static void make_huff_dev(const HuffTables &t, HuffDev *Tp)
{
    HuffDev &T = *Tp;
    memset(&T, 0, sizeof T);
    for (int k = 0; k < 2; ++k) {
        for (int cat = 0; cat < 12; ++cat)
            if (t.len[k][cat])
                T.dc[k][cat] = ((uint32_t)t.code[k][cat] << (32 - t.len[k][cat])) | (uint32_t)(t.len[k][cat] + cat);
        for (int rs = 0; rs < 256; ++rs) {
            const int cat = rs & 15, run = rs >> 4;
            if (!t.len[2 + k][rs] || cat > 10) continue;
            const uint32_t e = ((uint32_t)t.code[2 + k][rs] << (32 - t.len[2 + k][rs])) | (uint32_t)(t.len[2 + k][rs] + cat);
            if (rs == 0x00) T.ac[k][AC_EOB] = e;
            else if (rs == 0xF0) T.ac[k][AC_ZRL] = e;
            else if (cat >= 1) T.ac[k][run * AC_STRIDE + cat - 1] = e;
        }
    }
}

// How many segments per image: enough chains to keep a look-back short (~256 chains in flight), none
// shorter than 512 chunks.  1 = do not segment.
static uint32_t segments_for(uint32_t n, uint64_t total_mcus, uint64_t bpm)
{
    if (const char *e = getenv("PIXO_B200_SEGMENTS")) return (uint32_t)std::min(SEG_MAX, std::max(1, atoi(e)));   // test hook
    // Measured on B200: for 4K frames (6 075 chunks) the four extra launches cost more than the
    // shorter chains save (k_huff 72 -> 111 us for one frame, 715 -> 1110 us for 32); a 16 384^2 frame
    // (196 608 chunks on ONE chain) is where the look-back distance hurts.  So: few images, each long.
    const uint64_t chunks = total_mcus * bpm / CB;
    if (n > 8 || chunks < 16384) return 1;
    // 16 384^2 frame on one B200, whole device path: 16 segments 1.75 ms, 32: 1.46, 64: 1.27, 128: 1.19,
    // 256: 1.16 (k_huff<RAW> 729 -> 594 us from 64 to 256: with ~7000 chunks in flight a chain of 1/256
    // keeps the nearest inclusive prefix inside one 32-wide look-back step)
    uint32_t S = SEG_MAX / n;
    while (S > 1 && chunks / S < 512) S >>= 1;
    return S < 2 ? 1 : S;
}

struct SegPlan {
    uint32_t S;
    uint64_t seg_mcus, last_mcus;
    size_t raw_cap;                 // bytes per segment
    uint32_t max_tiles;             // per image
    EntropyPlan ent;                // status words etc. for n * S pseudo images
    size_t off_ent, off_rec, off_ntiles, off_cnt, total;   // offsets into the context's segment scratch
    size_t raw_bytes, off_bits, off_tails, raw_total;      // layout of the raw area: strings, then the segments' bit counts and tails
};

static SegPlan plan_segments(uint32_t n, uint32_t S, uint64_t total_mcus, uint64_t bpm, uint64_t mcu_raw_bytes)
{
    SegPlan p;
    p.seg_mcus = (total_mcus + S - 1) / S;
    S = (uint32_t)((total_mcus + p.seg_mcus - 1) / p.seg_mcus);   // no empty last segment
    p.S = S;
    p.last_mcus = total_mcus - p.seg_mcus * (S - 1);
    // room for a segment's raw string: as many bytes as its pixels take (q=100 noise stays below 0.7 of
    // that), at most what its blocks can possibly need (64 x 26 bits + DC < 216 bytes per block).  It
    // does not depend on the caller's output capacity, so an output that is too small is still measured.
    const uint64_t fair = p.seg_mcus * mcu_raw_bytes + 4096, worst = p.seg_mcus * bpm * 216 + 64;
    p.raw_cap = (size_t)a256(std::min(fair, worst));
    p.max_tiles = (uint32_t)(S * (p.raw_cap / SPL_TILE + 2));
    p.ent = plan_entropy(n * S, p.seg_mcus * bpm, 0);
    size_t o = 0;
    p.off_ent = o; o += a256(p.ent.total);
    p.raw_bytes = a256((size_t)n * S * p.raw_cap);
    p.off_bits = p.raw_bytes;
    p.off_tails = p.off_bits + a256((size_t)n * S * 8);
    p.raw_total = p.off_tails + a256((size_t)n * S * 8);
    p.off_rec = o; o += a256((size_t)n * S * sizeof(SegRec));
    p.off_ntiles = o; o += a256((size_t)n * 4);
    p.off_cnt = o; o += a256((size_t)n * p.max_tiles * 4);
    p.total = o;
    return p;
}

// k_huff<RAW> over n * S segments + the four splice kernels; final bytes in d_out, lengths / flags in
// the caller's scratch (same contract as the unsegmented launch).  base_bit / base_tail / last_stream:
// see SegParams (a band of a tiled frame passes its offset; a whole image passes 0, 0, true).
static int launch_segmented(pixo_b200_ctx *ctx, EntParams P, const HuffDev &T, uint32_t n, const FrameGeometry &g,
                            const SegPlan &sp, uint8_t *seg_scratch, uint8_t *raw_area, uint8_t *d_out, uint64_t out_cap,
                            uint64_t *d_out_len, uint32_t *d_overflow, uint64_t base_bit, uint32_t base_tail,
                            bool last_stream, bool code, bool splice, const uint64_t *base_dev = nullptr)
{
    cudaStream_t st = ctx->stream;
    const uint64_t bpm = g.y_per_mcu + (g.has_chroma ? 2 : 0);
    uint8_t *ent = seg_scratch + sp.off_ent;
    P.nimages = n * sp.S;
    P.seg_per_img = sp.S;
    P.nblocks = (uint32_t)(sp.seg_mcus * bpm);
    P.nblocks_last = (uint32_t)(sp.last_mcus * bpm);
    P.nchunks = (uint32_t)sp.ent.nchunks;
    P.seg_y_stride = (size_t)sp.seg_mcus * g.y_per_mcu * 64;
    P.seg_c_stride = (size_t)sp.seg_mcus * 64;
    P.rst_blocks = P.rst_mcus = P.cpi = 0;
    P.st_bits = reinterpret_cast<unsigned long long *>(ent + sp.ent.off_st1);
    P.st_ff = reinterpret_cast<unsigned long long *>(ent + sp.ent.off_st2);
    P.ticket = reinterpret_cast<uint32_t *>(ent + sp.ent.off_ticket);
    P.overflow = reinterpret_cast<uint32_t *>(ent + sp.ent.off_ovf);
    P.out_len = reinterpret_cast<uint64_t *>(ent + sp.ent.off_outlen);
    P.out_tail = reinterpret_cast<unsigned long long *>(ent + sp.ent.off_tail);
    P.out = raw_area;
    P.out_cap = sp.raw_cap;
    if (code) {
        PIXO_CUDA(ctx, cudaMemsetAsync(ent, 0, sp.ent.zero_bytes, st));
        const size_t want = ((size_t)P.nimages * P.nchunks + HUFF_WARPS - 1) / HUFF_WARPS;
        const unsigned grid = (unsigned)std::min<size_t>(want, (size_t)ctx->sm_count * HUFF_CTAS_PER_SM);
        k_huff<true><<<grid, 32 * HUFF_WARPS, 0, st>>>(P, T);
        ctx->launches += 1;
        PIXO_CUDA(ctx, cudaGetLastError());
        // the segments' bit counts and tails travel with the strings (a band is spliced by a later call)
        PIXO_CUDA(ctx, cudaMemcpyAsync(raw_area + sp.off_bits, P.out_len, (size_t)n * sp.S * 8, cudaMemcpyDeviceToDevice, st));
        PIXO_CUDA(ctx, cudaMemcpyAsync(raw_area + sp.off_tails, P.out_tail, (size_t)n * sp.S * 8, cudaMemcpyDeviceToDevice, st));
    }
    if (!splice) return 0;
    SegParams Q;
    Q.raw = P.out; Q.raw_cap = sp.raw_cap;
    Q.bits = reinterpret_cast<const unsigned long long *>(raw_area + sp.off_bits);
    Q.tails = reinterpret_cast<const unsigned long long *>(raw_area + sp.off_tails);
    Q.S = sp.S; Q.max_tiles = sp.max_tiles;
    Q.base_bit = base_bit; Q.base_tail = base_tail; Q.last_band = last_stream ? 1u : 0u;
    Q.base_dev = reinterpret_cast<const unsigned long long *>(base_dev);
    Q.rec = reinterpret_cast<SegRec *>(seg_scratch + sp.off_rec);
    Q.ntiles = reinterpret_cast<uint32_t *>(seg_scratch + sp.off_ntiles);
    Q.cnt = reinterpret_cast<uint32_t *>(seg_scratch + sp.off_cnt);
    Q.out = d_out; Q.out_cap = out_cap;
    Q.out_len = reinterpret_cast<unsigned long long *>(d_out_len);
    Q.overflow = d_overflow;
    Q.raw_overflow = code ? P.overflow : nullptr;   // a band's flags were checked by the host when it was coded
    k_seg_prefix<<<n, SPL_THREADS, 0, st>>>(Q);
    k_seg_count<<<dim3((sp.max_tiles + SPL_TPC - 1) / SPL_TPC, n), SPL_THREADS, 0, st>>>(Q);
    k_seg_scan<<<n, 1024, 0, st>>>(Q);
    k_seg_emit<<<dim3((sp.max_tiles + SPL_TPC - 1) / SPL_TPC, n), SPL_THREADS, 0, st>>>(Q);
    ctx->launches += 4;
    PIXO_CUDA(ctx, cudaGetLastError());
    return 0;
}

// Enqueue the entropy stage for n images (natural-order coefficient arrays) on ctx->stream.
// d_scratch: entropy_scratch_bytes.  d_out: n * out_cap bytes of scan data; *d_out_len /
// *d_overflow point into the scratch.
int launch_jpeg_entropy(pixo_b200_ctx *ctx, const int16_t *d_y, size_t y_stride, const int16_t *d_cb,
                        const int16_t *d_cr, size_t c_stride, uint32_t n, const FrameGeometry &g,
                        const HuffTables &t, uint32_t restart_interval, uint8_t *d_scratch, uint8_t *d_out,
                        uint64_t out_cap, uint64_t **d_out_len, uint32_t **d_overflow, const int *dc_seed,
                        uint64_t **d_raw_tail)
{
    const bool raw = d_raw_tail != nullptr;
    if (raw && restart_interval)
        return set_error(ctx, PIXO_B200_ERR_UNSUPPORTED, "band-local raw coding does not take a restart interval");
    if (raw && (out_cap & 3))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "raw buffer capacity must be a multiple of 4");
    const uint64_t nblocks = g.ny + 2 * g.nc;
    const uint64_t bpm_ = g.y_per_mcu + (g.has_chroma ? 2 : 0);
    uint64_t rst_blocks = (uint64_t)restart_interval * bpm_;
    if (rst_blocks >= nblocks) rst_blocks = 0;  // a single interval: no marker is ever written
    const EntropyPlan pl = plan_entropy(n, nblocks, rst_blocks);
    if (nblocks > 0xFFFFFFFFull || (uint64_t)n * pl.nchunks > 0x7FFFFFFFull)
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "entropy stage: too many blocks per call");
    EntParams P;
    P.y = d_y; P.cb = d_cb; P.cr = d_cr; P.y_stride = y_stride; P.c_stride = c_stride;
    P.bpm = g.y_per_mcu + (g.has_chroma ? 2 : 0);
    P.y_per_mcu = g.y_per_mcu;
    P.nblocks = (uint32_t)nblocks;
    P.nchunks = (uint32_t)pl.nchunks;
    P.rst_blocks = (uint32_t)rst_blocks;
    P.rst_mcus = rst_blocks ? restart_interval : 0u;
    P.cpi = rst_blocks ? (uint32_t)((rst_blocks + CB - 1) / CB) : 0u;
    P.nimages = n;
    P.st_bits = reinterpret_cast<unsigned long long *>(d_scratch + pl.off_st1);
    P.st_ff = reinterpret_cast<unsigned long long *>(d_scratch + pl.off_st2);
    P.ticket = reinterpret_cast<uint32_t *>(d_scratch + pl.off_ticket);
    P.overflow = reinterpret_cast<uint32_t *>(d_scratch + pl.off_ovf);
    P.out_len = reinterpret_cast<uint64_t *>(d_scratch + pl.off_outlen);
    P.out = d_out; P.out_cap = out_cap;
    P.out_tail = reinterpret_cast<unsigned long long *>(d_scratch + pl.off_tail);
    for (int k = 0; k < 3; ++k) P.dc_seed[k] = dc_seed ? dc_seed[k] : 0;
    P.dc_seed_dev = nullptr;
    *d_out_len = P.out_len;
    *d_overflow = P.overflow;
    if (raw) *d_raw_tail = reinterpret_cast<uint64_t *>(P.out_tail);

    HuffDev T;
    make_huff_dev(t, &T);
    cudaStream_t st = ctx->stream;
    PIXO_CUDA(ctx, cudaMemsetAsync(d_scratch, 0, pl.zero_bytes, st));
    P.seg_per_img = 1; P.nblocks_last = P.nblocks; P.seg_y_stride = P.seg_c_stride = 0;
    // few images: cut each into segments (short look-back chains) and splice - see k_seg_*
    const uint32_t S = (rst_blocks == 0 && !ctx->no_segments) ? segments_for(n, g.total_mcus(), bpm_) : 1;
    const uint64_t mcu_raw = (uint64_t)g.y_per_mcu * 64 * (g.has_chroma ? 3 : 1);
    SegPlan sp = plan_segments(n, S, g.total_mcus(), bpm_, mcu_raw);
    if (raw && sp.S > 1 && sp.raw_total > out_cap) sp.S = 1;   // the caller's band buffer has no room for segments
    ctx->last_band_segments = 1;
    if (sp.S > 1) {
        PIXO_TRY(ensure_dev(ctx, ctx->d_raw, sp.total + (raw ? 0 : sp.raw_total)));
        auto *seg_scratch = reinterpret_cast<uint8_t *>(ctx->d_raw.ptr);
        if (!raw)
            return launch_segmented(ctx, P, T, n, g, sp, seg_scratch, seg_scratch + sp.total, d_out, out_cap, P.out_len,
                                    P.overflow, 0, 0, true, true, true);
        // a band of a tiled frame: code now into the CALLER's raw buffer, splice when the bit offset is
        // known (launch_band_splice_segments); the per-segment bit counts and tails are summed up by the caller
        pixo_b200_ctx::BandInfo bi;
        bi.segments = sp.S; bi.bpm = (uint32_t)bpm_; bi.y_per_mcu = g.y_per_mcu; bi.has_chroma = g.has_chroma;
        bi.mcus = g.total_mcus();
        ctx->bands[d_out] = bi;
        ctx->last_band_segments = sp.S;
        PIXO_TRY(launch_segmented(ctx, P, T, n, g, sp, seg_scratch, d_out, nullptr, 0, nullptr, nullptr, 0, 0, false, true, false));
        *d_out_len = reinterpret_cast<uint64_t *>(seg_scratch + sp.off_ent + sp.ent.off_outlen);
        *d_overflow = reinterpret_cast<uint32_t *>(seg_scratch + sp.off_ent + sp.ent.off_ovf);
        *d_raw_tail = reinterpret_cast<uint64_t *>(seg_scratch + sp.off_ent + sp.ent.off_tail);
        return 0;
    }
    if (raw) ctx->bands.erase(d_out);
    const size_t want = ((size_t)n * pl.nchunks + HUFF_WARPS - 1) / HUFF_WARPS;
    const unsigned grid = (unsigned)std::min<size_t>(want, (size_t)ctx->sm_count * HUFF_CTAS_PER_SM);
    if (raw) k_huff<true><<<grid, 32 * HUFF_WARPS, 0, st>>>(P, T);
    else k_huff<false><<<grid, 32 * HUFF_WARPS, 0, st>>>(P, T);
    ctx->launches += 1;
    PIXO_CUDA(ctx, cudaGetLastError());
    return 0;
}


// Splice the raw segments pixo_b200_jpeg_band_entropy_dev left in the caller's buffer d_raw into the
// band's scan bytes.  (Call only when the context knows d_raw as a segmented band.)
int launch_band_splice_segments(pixo_b200_ctx *ctx, const uint8_t *d_raw, uint64_t base_bit, uint32_t base_tail,
                                bool last, uint8_t *d_scratch, uint8_t *d_out, uint64_t out_cap,
                                uint64_t **d_out_len, uint32_t **d_overflow)
{
    const pixo_b200_ctx::BandInfo bi = ctx->bands.at(d_raw);
    FrameGeometry g;
    g.y_per_mcu = bi.y_per_mcu; g.has_chroma = bi.has_chroma;
    const uint64_t mcu_raw = (uint64_t)g.y_per_mcu * 64 * (g.has_chroma ? 3 : 1);
    const SegPlan sp = plan_segments(1, bi.segments, bi.mcus, bi.bpm, mcu_raw);
    PIXO_TRY(ensure_dev(ctx, ctx->d_raw, sp.total));
    EntParams P;
    memset(&P, 0, sizeof P);
    HuffDev T;
    memset(&T, 0, sizeof T);
    PIXO_CUDA(ctx, cudaMemsetAsync(d_scratch, 0, 256, ctx->stream));
    *d_out_len = reinterpret_cast<uint64_t *>(d_scratch);
    *d_overflow = reinterpret_cast<uint32_t *>(d_scratch + 8);
    return launch_segmented(ctx, P, T, 1, g, sp, reinterpret_cast<uint8_t *>(ctx->d_raw.ptr), const_cast<uint8_t *>(d_raw),
                            d_out, out_cap, *d_out_len, *d_overflow, base_bit, base_tail, last, false, true);
}

// Stream-ordered band coding: predictors from device memory in, {bits, tail} and flags to device
// memory out, no host synchronisation.  Always uses the segment machinery (S >= 1) with the strings
// in the caller's buffer.
int launch_band_entropy_async(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb, const int16_t *d_cr,
                              const FrameGeometry &g, const HuffTables &t, const int *d_seed, uint8_t *d_raw,
                              uint64_t raw_cap, uint64_t *d_bits_tail, uint32_t *d_flags)
{
    const uint64_t bpm = g.y_per_mcu + (g.has_chroma ? 2 : 0);
    const uint64_t mcu_raw = (uint64_t)g.y_per_mcu * 64 * (g.has_chroma ? 3 : 1);
    uint32_t S = ctx->no_segments ? 1 : segments_for(1, g.total_mcus(), bpm);
    SegPlan sp = plan_segments(1, S, g.total_mcus(), bpm, mcu_raw);
    if (sp.raw_total > raw_cap && sp.S > 1) sp = plan_segments(1, 1, g.total_mcus(), bpm, mcu_raw);
    if (sp.raw_total > raw_cap)
        return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "raw capacity %llu too small (need %zu)",
                         (unsigned long long)raw_cap, sp.raw_total);
    if ((g.ny + 2 * g.nc) > 0xFFFFFFFFull)
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "entropy stage: too many blocks per call");
    PIXO_TRY(ensure_dev(ctx, ctx->d_raw, sp.total));
    auto *seg_scratch = reinterpret_cast<uint8_t *>(ctx->d_raw.ptr);
    EntParams P;
    memset(&P, 0, sizeof P);
    P.y = d_y; P.cb = d_cb; P.cr = d_cr;
    P.bpm = (uint32_t)bpm; P.y_per_mcu = g.y_per_mcu;
    P.dc_seed_dev = d_seed;
    HuffDev T;
    make_huff_dev(t, &T);
    pixo_b200_ctx::BandInfo bi;
    bi.segments = sp.S; bi.bpm = (uint32_t)bpm; bi.y_per_mcu = g.y_per_mcu; bi.has_chroma = g.has_chroma;
    bi.mcus = g.total_mcus();
    ctx->bands[d_raw] = bi;
    PIXO_TRY(launch_segmented(ctx, P, T, 1, g, sp, seg_scratch, d_raw, nullptr, 0, nullptr, nullptr, 0, 0, false, true, false));
    k_band_totals<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long *>(d_raw + sp.off_bits),
                                             reinterpret_cast<const unsigned long long *>(d_raw + sp.off_tails),
                                             reinterpret_cast<const uint32_t *>(seg_scratch + sp.off_ent + sp.ent.off_ovf), sp.S,
                                             reinterpret_cast<unsigned long long *>(d_bits_tail), d_flags);
    ctx->launches += 1;
    PIXO_CUDA(ctx, cudaGetLastError());
    return 0;
}

int launch_band_splice_async(pixo_b200_ctx *ctx, const uint8_t *d_raw, const uint64_t *d_offset, uint8_t *d_out,
                             uint64_t out_cap, uint64_t *d_out_len, uint32_t *d_flags)
{
    auto it = ctx->bands.find(d_raw);
    if (it == ctx->bands.end())
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "this raw buffer was not coded by pixo_b200_jpeg_band_entropy_dev_async");
    const pixo_b200_ctx::BandInfo bi = it->second;
    FrameGeometry g;
    g.y_per_mcu = bi.y_per_mcu; g.has_chroma = bi.has_chroma;
    const SegPlan sp = plan_segments(1, bi.segments, bi.mcus, bi.bpm, (uint64_t)g.y_per_mcu * 64 * (g.has_chroma ? 3 : 1));
    PIXO_TRY(ensure_dev(ctx, ctx->d_raw, sp.total));
    EntParams P;
    memset(&P, 0, sizeof P);
    HuffDev T;
    memset(&T, 0, sizeof T);
    return launch_segmented(ctx, P, T, 1, g, sp, reinterpret_cast<uint8_t *>(ctx->d_raw.ptr), const_cast<uint8_t *>(d_raw),
                            d_out, out_cap, d_out_len, d_flags, 0, 0, false, false, true, d_offset);
}

size_t band_raw_bytes_segmented(const FrameGeometry &g)
{
    const uint64_t bpm = g.y_per_mcu + (g.has_chroma ? 2 : 0);
    const uint32_t S = segments_for(1, g.total_mcus(), bpm);
    if (S < 2) return 0;
    return plan_segments(1, S, g.total_mcus(), bpm, (uint64_t)g.y_per_mcu * 64 * (g.has_chroma ? 3 : 1)).raw_total;
}

}  // namespace pixo
