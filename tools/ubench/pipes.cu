// Microbenchmark: issue cost (cycles per warp instruction per SM sub-partition) of the
// instructions K1 is made of.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int MODE>
__global__ void k(unsigned *out, unsigned a, unsigned b)
{
    unsigned x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a * (i + 1) + threadIdx.x;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) asm volatile("dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
            if (MODE == 1) asm volatile("prmt.b32 %0, %0, %1, 0x7531;" : "+r"(x[i]) : "r"(b));
            if (MODE == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(x[i]) : "r"(b), "r"(a));
            if (MODE == 3) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
            if (MODE == 4) asm volatile("shf.r.wrap.b32 %0, %0, %1, 8;" : "+r"(x[i]) : "r"(b));
            if (MODE == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
            if (MODE == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
            if (MODE == 7) { float f = __uint_as_float(x[i]); f = __fmul_rn(f, 1.0000001f); x[i] = __float_as_uint(f); }
            if (MODE == 8) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
        }
        if (MODE == 9) {   // dp4a + packed fadd2 1:1 (different pipes?)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
#pragma unroll
            for (int i = 8; i < 16; i += 2) {
                unsigned long long v, w;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "r"(x[i]), "r"(x[i + 1]));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(w) : "r"(b), "r"(b));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=r"(x[i]), "=r"(x[i + 1]) : "l"(v));
            }
        }
        if (MODE == 10) {  // prmt + lop3 + dp4a mix (all "ALU"?)
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                asm volatile("dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
                asm volatile("prmt.b32 %0, %0, %1, 0x7531;" : "+r"(x[i + 1]) : "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(x[i + 2]) : "r"(b), "r"(a));
                asm volatile("max.u16x2 %0, %0, %1;" : "+r"(x[i + 3]) : "r"(b));
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double n_instr)
{
    unsigned *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 3, 5); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, 3, 5); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // warp-instructions per SMSP: 148*8 blocks * 8 warps / (148*4) = 16 warps per SMSP
    double winstr_per_smsp = 16.0 * ITER * n_instr;
    printf("%-26s %7.3f ms   %.2f cycles per warp-instr per SMSP (at 1.9 GHz)\n", name, ms, ms * 1e-3 * 1.9e9 / winstr_per_smsp);
    cudaFree(out);
}
int main()
{
    run<0>("IDP.4A", 16); run<1>("PRMT", 16); run<2>("LOP3", 16); run<3>("VIMNMX.U16x2", 16);
    run<4>("SHF", 16); run<5>("IADD", 16); run<6>("IMAD", 16); run<7>("FMUL", 16); run<8>("VABSDIFF4", 16);
    run<9>("8 IDP + 4 FADD2", 12); run<10>("IDP/PRMT/LOP3/VIMNMX mix", 16);
    return 0;
}
