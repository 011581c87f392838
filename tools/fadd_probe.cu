// tools/fadd_probe.cu -- how long does a sequential float sum out of shared memory take for ONE warp?
// (development probe for fastmatch_kernel's phase 2)   nvcc -arch=sm_100a -O3 -o tools/build/fadd_probe tools/fadd_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
struct Big { float v[960]; };
// probe2: the same loop in fastmatch_kernel's clothes -- __launch_bounds__(1024, 1), a 4 KB parameter block, the
// programmatic-dependent-launch instructions, register pressure
__global__ void __launch_bounds__(1024, 1) probe2(float *out, long long *cyc, int n, int nbp, const __grid_constant__ Big T)
{
    extern __shared__ __align__(16) float vals[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int i = threadIdx.x; i < 32 * nbp; i += blockDim.x) vals[i] = 1.0f + (i % 7) * 0.125f + T.v[i % 960];
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x < 27) {
        const float *v = vals + threadIdx.x * nbp;
        long long t0 = clock64();
        float s = 0.f;
        int i = 0;
        float4 a = *reinterpret_cast<const float4 *>(v), b = *reinterpret_cast<const float4 *>(v + 4);
        for (; i + 16 <= n; i += 8) {
            const float4 na = *reinterpret_cast<const float4 *>(v + i + 8), nb = *reinterpret_cast<const float4 *>(v + i + 12);
            s = __fadd_rn(s, a.x); s = __fadd_rn(s, a.y); s = __fadd_rn(s, a.z); s = __fadd_rn(s, a.w);
            s = __fadd_rn(s, b.x); s = __fadd_rn(s, b.y); s = __fadd_rn(s, b.z); s = __fadd_rn(s, b.w);
            a = na; b = nb;
        }
        long long t1 = clock64();
        out[threadIdx.x] = s;
        if (threadIdx.x == 0) cyc[0] = t1 - t0;
    }
    __syncthreads();
}

// probe3: probe2 with ~60 live registers per thread (the register file of the SM completely allocated, as in
// fastmatch_kernel) and phase-1-like traffic before the loop
__global__ void __launch_bounds__(1024, 1) probe3(float *out, long long *cyc, int n, int nbp, const float *g, int stride)
{
    extern __shared__ __align__(16) float vals[];
    float keep[40];
#pragma unroll
    for (int q = 0; q < 40; ++q) keep[q] = g[(threadIdx.x * 40 + q) * stride];
    for (int i = threadIdx.x; i < 32 * nbp; i += blockDim.x) vals[i] = 1.0f + (i % 7) * 0.125f + __ldg(g + (i * 37) % 4096);
    __syncthreads();
    if (threadIdx.x < 27) {
        const float *v = vals + threadIdx.x * nbp;
        long long t0 = clock64();
        float s = 0.f;
        int i = 0;
        float4 a = *reinterpret_cast<const float4 *>(v), b = *reinterpret_cast<const float4 *>(v + 4);
        for (; i + 16 <= n; i += 8) {
            const float4 na = *reinterpret_cast<const float4 *>(v + i + 8), nb = *reinterpret_cast<const float4 *>(v + i + 12);
            s = __fadd_rn(s, a.x); s = __fadd_rn(s, a.y); s = __fadd_rn(s, a.z); s = __fadd_rn(s, a.w);
            s = __fadd_rn(s, b.x); s = __fadd_rn(s, b.y); s = __fadd_rn(s, b.z); s = __fadd_rn(s, b.w);
            a = na; b = nb;
        }
        long long t1 = clock64();
        out[threadIdx.x] = s;
        if (threadIdx.x == 0) cyc[0] = t1 - t0;
    }
    __syncthreads();
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 40; ++q) acc += keep[q];
    out[1024 + threadIdx.x] = acc;
}

__global__ void probe(float *out, long long *cyc, int n, int nbp, int mode, int nthreads_active)
{
    extern __shared__ __align__(16) float vals[];
    for (int i = threadIdx.x; i < 32 * nbp; i += blockDim.x) vals[i] = 1.0f + (i % 7) * 0.125f;
    __syncthreads();
    if (threadIdx.x < nthreads_active) {
        const float *v = vals + threadIdx.x * nbp;
        long long t0 = clock64();
        float s = 0.f;
        if (mode == 0) {                      // scalar loads
            for (int i = 0; i < n; ++i) s = __fadd_rn(s, v[i]);
        } else if (mode == 1) {               // float4 x2 per iteration, prefetched
            int i = 0;
            float4 a = *reinterpret_cast<const float4 *>(v), b = *reinterpret_cast<const float4 *>(v + 4);
            for (; i + 16 <= n; i += 8) {
                const float4 na = *reinterpret_cast<const float4 *>(v + i + 8), nb = *reinterpret_cast<const float4 *>(v + i + 12);
                s = __fadd_rn(s, a.x); s = __fadd_rn(s, a.y); s = __fadd_rn(s, a.z); s = __fadd_rn(s, a.w);
                s = __fadd_rn(s, b.x); s = __fadd_rn(s, b.y); s = __fadd_rn(s, b.z); s = __fadd_rn(s, b.w);
                a = na; b = nb;
            }
            s = __fadd_rn(s, a.x); s = __fadd_rn(s, a.y); s = __fadd_rn(s, a.z); s = __fadd_rn(s, a.w);
            s = __fadd_rn(s, b.x); s = __fadd_rn(s, b.y); s = __fadd_rn(s, b.z); s = __fadd_rn(s, b.w);
            i += 8;
            for (; i < n; ++i) s = __fadd_rn(s, v[i]);
        } else {                              // registers only: the bare dependent chain
            float x = v[0];
            for (int i = 0; i < n; ++i) s = __fadd_rn(s, x);
        }
        long long t1 = clock64();
        out[threadIdx.x] = s;
        if (threadIdx.x == 0) cyc[0] = t1 - t0;
    }
    __syncthreads();
}
int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 8);
    const int n = 1079, nbp = 1092;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int threads : {32, 1024})
        for (int active : {1, 27})
            for (int mode = 0; mode < 3; ++mode) {
                for (int rep = 0; rep < 2; ++rep) { probe<<<1, threads, 32 * nbp * 4>>>(out, cyc, n, nbp, mode, active); cudaDeviceSynchronize(); }
                printf("block %4d threads, %2d summing, mode %d: %lld cycles = %.1f per add\n", threads, active, mode, cyc[0], (double)cyc[0] / n);
            }
    {
        Big T = {};
        cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int pdl = 0; pdl < 2; ++pdl)
            for (int rep = 0; rep < 3; ++rep) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(1); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 32 * nbp * 4;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr; cfg.numAttrs = pdl;
                cudaLaunchKernelEx(&cfg, probe2, out, cyc, n, nbp, T);
                cudaDeviceSynchronize();
                printf("probe2 (1024 threads, launch bounds, 4 KB params, griddepcontrol, pdl attr %d): %lld cycles = %.1f per add\n", pdl, cyc[0], (double)cyc[0] / n);
            }
    }
    {
        float *g; cudaMalloc(&g, 1 << 22); cudaMemset(g, 0, 1 << 22);
        float *out3; cudaMalloc(&out3, 16384);
        cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int rep = 0; rep < 3; ++rep) {
            probe3<<<1, 1024, 32 * nbp * 4>>>(out3, cyc, n, nbp, g, 1);
            cudaDeviceSynchronize();
            printf("probe3 (64 registers per thread, register file full): %lld cycles = %.1f per add\n", cyc[0], (double)cyc[0] / n);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
