/* A plain C client of the C ABI: what a cgo / Rust-FFI / JNI binding sees.  Built and run by
 * tests/test_abi.py (strict C11, all warnings) - proves the header is usable outside C++ and that
 * the host-only entry points work without a device. */
#include <stdio.h>
#include <string.h>

#include "pixo_b200.h"

int main(void)
{
    unsigned char lum_zz[64], chr_zz[64];
    float lum[64], chr[64];
    size_t ny = 0, nc = 0;
    if (pixo_b200_version() != 0x000100) return 1;
    pixo_b200_quant_tables(50, lum_zz, chr_zz, lum, chr);
    if (lum_zz[0] != 16 || lum[0] != 16.0f) return 2;
    if (pixo_b200_jpeg_block_counts(3840, 2160, PIXO_B200_RGB, PIXO_B200_S420, &ny, &nc) != PIXO_B200_OK ||
        ny != 129600 || nc != 32400) return 3;
    /* host-only entropy coder: one 8x8 gray block of zeros -> a complete baseline JPEG */
    {
        short y[64];
        unsigned char out[2048];
        size_t n = 0;
        memset(y, 0, sizeof y);
        if (pixo_b200_jpeg_entropy_encode(NULL, y, NULL, NULL, 8, 8, PIXO_B200_GRAY, 80, PIXO_B200_S444, 0, 0, out,
                                          sizeof out, &n) != PIXO_B200_OK) return 4;
        if (n < 100 || out[0] != 0xFF || out[1] != 0xD8 || out[n - 2] != 0xFF || out[n - 1] != 0xD9) return 5;
    }
    {
        pixo_b200_ctx *ctx = NULL;
        const int rc = pixo_b200_ctx_create(0, &ctx);
        if (pixo_b200_device_count() == 0) {
            if (rc != PIXO_B200_ERR_CUDA || ctx != NULL) return 6;       /* no device: loud failure */
            if (strstr(pixo_b200_last_error(NULL), "no CPU fallback") == NULL) return 7;
        } else {
            if (rc != PIXO_B200_OK) return 8;
            pixo_b200_ctx_destroy(ctx);
        }
    }
    puts("abi_client ok");
    return 0;
}
