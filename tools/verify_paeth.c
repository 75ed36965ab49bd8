/* Exhaustive check (2^24 byte triples, in all four byte lanes at once) of the SWAR Paeth predictor
 * used by png_filter.cu against the scalar definition (src/simd/fallback.rs:143-159).
 * Host emulation of the device primitives: VABSDIFF4.U8 and PRMT sign replication.
 *   gcc -O2 -o /tmp/verify_paeth tools/verify_paeth.c && /tmp/verify_paeth */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static uint32_t absdiff4(uint32_t x, uint32_t y)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        int a = (x >> (8 * i)) & 0xFF, b = (y >> (8 * i)) & 0xFF;
        r |= (uint32_t)abs(a - b) << (8 * i);
    }
    return r;
}
static uint32_t signrep4(uint32_t x) /* prmt x, 0, 0xBA98: every byte becomes 0xFF if its msb is set */
{
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i)
        if ((x >> (8 * i)) & 0x80) r |= 0xFFu << (8 * i);
    return r;
}
#define H 0x80808080u
/* bit 7 of every byte: x > y (unsigned).  (x | H) - (y & ~H): per byte 128 + x7 - y7, never borrows;
 * its bit 7 = (x7 >= y7), x7/y7 the low 7 bits.  x > y  <=>  msb(x) > msb(y), or equal msbs and x7 > y7;
 * written with >= on (x, y+... ) the device code uses the form below (gt7). */
static uint32_t gt7(uint32_t x, uint32_t y)
{
    /* x > y  <=>  !(y >= x);  ge(y, x) bit7 = (ym & ~xm) | (~(ym ^ xm) & t), t = (y|H) - (x&~H) */
    const uint32_t t = (y | H) - (x & ~H);
    const uint32_t ge = (y & ~x) | (~(y ^ x) & t);
    return ~ge & H;
}
static uint32_t sel4(uint32_t m, uint32_t x, uint32_t y) { return (x & m) | (y & ~m); }

static uint32_t paeth_swar(uint32_t a, uint32_t b, uint32_t c)
{
    const uint32_t pa = absdiff4(b, c), pb = absdiff4(a, c), dab = absdiff4(a, b);
    const uint32_t m1 = signrep4(gt7(pa, pb));            /* 0xFF where pa > pb: b is the nearer one */
    const uint32_t near = sel4(m1, b, a);
    const uint32_t mn = sel4(m1, pb, pa), mx = sel4(m1, pa, pb);
    const uint32_t adiff = mx - mn;                          /* per byte, no borrow: mx >= mn */
    /* c wins iff it lies within [min(a,b), max(a,b)] (mx <= dab) and the nearer endpoint is farther
     * from p than c is (mn > mx - mn) */
    const uint32_t cw = signrep4(gt7(mn, adiff) & ~gt7(mx, dab));
    return sel4(cw, c, near);
}
static int paeth_ref(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
int main(void)
{
    unsigned long long bad = 0, n = 0;
    for (int a = 0; a < 256; ++a)
        for (int b = 0; b < 256; ++b)
            for (int c = 0; c < 256; c += 4) {
                uint32_t A = (uint32_t)a * 0x01010101u, B = (uint32_t)b * 0x01010101u;
                uint32_t C = (uint32_t)c | ((uint32_t)(c + 1) << 8) | ((uint32_t)(c + 2) << 16) | ((uint32_t)(c + 3) << 24);
                uint32_t r = paeth_swar(A, B, C);
                for (int i = 0; i < 4; ++i, ++n)
                    if ((int)((r >> (8 * i)) & 0xFF) != paeth_ref(a, b, c + i)) ++bad;
            }
    printf("%llu triples, %llu mismatches\n", n, bad);
    return bad != 0;
}
