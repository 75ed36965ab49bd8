// Microbenchmark: does packed f32x2 arithmetic (add/mul/fma .f32x2, sm_100+) raise FP32
// throughput per issue slot on B200?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
template <int MODE>
__global__ void k(float *out, float a, float b)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + i + threadIdx.x;
    for (int it = 0; it < ITER; ++it) {
        if (MODE == 0) {  // scalar FADD (non-contracted)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fadd_rn(x[i], b);
        } else if (MODE == 1) {  // packed add
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v, w;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(w) : "f"(b), "f"(b));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
            }
        } else if (MODE == 2) {  // scalar FMUL
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fmul_rn(x[i], b);
        } else if (MODE == 3) {  // packed mul
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v, w;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(w) : "f"(b), "f"(b));
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
            }
        } else if (MODE == 4) {  // scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], b, a);
        } else if (MODE == 5) {  // packed fma
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v, w, u;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(w) : "f"(b), "f"(b));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(u) : "f"(a), "f"(a));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(w), "l"(u));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
            }
        } else if (MODE == 6) {  // mixed: scalar FADD + integer LOP3 interleaved 1:1
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __fadd_rn(x[i], b);
#pragma unroll
            for (int i = 8; i < 16; ++i) x[i] = __int_as_float((__float_as_int(x[i]) ^ 0x5555) + 3);
        } else if (MODE == 7) {  // mixed: packed add (8 floats) + integer ops on the other 8
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                unsigned long long v, w;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(w) : "f"(b), "f"(b));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
            }
#pragma unroll
            for (int i = 8; i < 16; ++i) x[i] = __int_as_float((__float_as_int(x[i]) ^ 0x5555) + 3);
        } else if (MODE == 8) {  // packed add with round-toward-zero
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v, w;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(w) : "f"(b), "f"(b));
                asm volatile("add.rz.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
            }
        } else if (MODE == 9) {  // F2I throughput
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __int_as_float(__float2int_rz(x[i]) + 0x3f800000);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, double ops_per_iter)
{
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 1.0f, 1.0000001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, 1.0f, 1.0000001f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = 148.0 * 8 * 256 * ITER * ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f Tlane-op/s  (%.1f lane-ops/clk/SM @1.9GHz)\n", name, ms,
           lane_ops / ms / 1e9, lane_ops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}

int main()
{
    run<0>("FADD scalar", 16);
    run<1>("add.f32x2", 16);
    run<2>("FMUL scalar", 16);
    run<3>("mul.f32x2", 16);
    run<4>("FFMA scalar", 16);
    run<5>("fma.f32x2", 16);
    run<6>("FADD + int 1:1 (16 ops)", 16);
    run<7>("add.f32x2 + int (16 ops)", 16);
    run<8>("add.rz.f32x2", 16);
    run<9>("F2I+IADD (32 ops)", 32);
    return 0;
}
