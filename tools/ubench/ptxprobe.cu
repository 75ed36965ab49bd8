// compile-only probe: which packed f32x2 forms does ptxas accept for sm_100a?
__global__ void k(unsigned long long *p, const double *c)
{
    unsigned long long a = p[0], b = p[1], d;
    asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); p[2] = d;
    asm volatile("add.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); p[3] = d;
    asm volatile("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); p[4] = d;
    asm volatile("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(p[7])); p[5] = d;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(*(const unsigned long long*)c)); p[6] = d;
}
