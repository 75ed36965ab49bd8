// Do scalar FP32 ops (which may run on the "fmalite" half of the FP32 datapath) overlap with
// packed f32x2 / IDP / IMAD ops (fmaheavy)?  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm volatile("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int NS, int NP, int ND>   // per iteration: NS scalar FMUL, NP packed FADD2, ND dp4a  (independent chains)
__global__ void k(float *out, float a, unsigned b)
{
    float s[16]; u64 p[8]; unsigned d[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = a + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) { p[i] = pk(a + i, a - i); d[i] = b + i; }
    const u64 w = pk(a, a);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < NS) s[i] = __fmul_rn(s[i], 1.0000001f);
            if (i < NP) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i % 8]) : "l"(w));
            if (i < ND) asm volatile("dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(d[i % 8]) : "r"(b), "r"(b));
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += s[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += (float)(p[i] & 0xffff) + (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int NS, int NP, int ND> void run()
{
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NS, NP, ND><<<148 * 8, 256>>>(out, 1.0f, 5); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<NS, NP, ND><<<148 * 8, 256>>>(out, 1.0f, 5); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("scalar FMUL %2d + FADD2 %2d + IDP %2d per iter: %7.3f ms  = %.2f cycles/iter/warp-per-SMSP (16 warps)\n", NS, NP, ND, ms,
           ms * 1e-3 * 1.9e9 / (16.0 * ITER));
    cudaFree(out);
}
int main()
{
    run<16, 0, 0>(); run<0, 8, 0>(); run<0, 0, 8>();
    run<8, 4, 0>(); run<16, 8, 0>(); run<8, 8, 0>(); run<4, 8, 0>();
    run<8, 0, 8>(); run<8, 4, 4>(); run<0, 8, 8>();
    return 0;
}
