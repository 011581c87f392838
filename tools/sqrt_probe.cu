// Development probe: is sqrt.approx.f32 (MUFU.SQRT) correctly rounded on small integers?
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
__global__ void k(float *a, float *b, float *c, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = (float)i, r;
    asm volatile("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    a[i] = r;
    asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    c[i] = r;
    b[i] = __fsqrt_rn(x);
}
int main()
{
    const int n = 1 << 16;
    float *a, *b, *c;
    cudaMallocManaged(&a, n * 4); cudaMallocManaged(&b, n * 4); cudaMallocManaged(&c, n * 4);
    k<<<n / 256, 256>>>(a, b, c, n);
    cudaDeviceSynchronize();
    int bad = 0, badc = 0, first = -1, badhost = 0;
    for (int i = 0; i < n; ++i) {
        if (a[i] != b[i]) { bad++; if (first < 0) first = i; }
        if (c[i] != b[i]) badc++;
        if (b[i] != sqrtf((float)i)) badhost++;
    }
    int bad226 = 0;
    for (int i = 0; i <= 226; ++i) if (a[i] != b[i]) { bad226++; printf("  d2=%d approx=%.9g rn=%.9g\n", i, a[i], b[i]); }
    printf("approx != rn: %d of %d (first %d); ftz variant: %d; rn != host sqrtf: %d; within [0,226]: %d\n", bad, n, first, badc, badhost, bad226);
    return 0;
}
