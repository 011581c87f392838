// tools/fadd_probe2.cu -- the summing warp's inner loop of fastmatch_kernel on its own: straight-line groups of
// 32 x (LDS.128 + 4 dependent FADD).  What does one add cost, and what changes it?
//   nvcc -arch=sm_100a -O3 -fmad=false -o tools/build/fadd_probe2 tools/fadd_probe2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(float *out, long long *cyc, int ngroup, int nbp, int active, int busy_warps, unsigned int busy_mask = 0, const float *g = nullptr, int spread = 0)
{
    extern __shared__ __align__(16) float vals[];
    for (int i = threadIdx.x; i < 32 * nbp; i += blockDim.x) vals[i] = 1.0f + (i % 7) * 0.125f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        const float4 *v = reinterpret_cast<const float4 *>(vals + (lane < active ? lane : 0) * nbp);
        long long t0 = clock64();
        float s = 0.f;
        for (int g = 0; g < ngroup; ++g) {
            const float4 *p = v + g * 32;
            if (MODE == 0) {
#pragma unroll
                for (int q = 0; q < 32; ++q) { const float4 x = p[q]; s = __fadd_rn(s, x.x); s = __fadd_rn(s, x.y); s = __fadd_rn(s, x.z); s = __fadd_rn(s, x.w); }
            } else if (MODE == 1) {           // all 8 loads of a chunk first, then its 32 adds
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 x[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) x[q] = p[c * 8 + q];
#pragma unroll
                    for (int q = 0; q < 8; ++q) { s = __fadd_rn(s, x[q].x); s = __fadd_rn(s, x[q].y); s = __fadd_rn(s, x[q].z); s = __fadd_rn(s, x[q].w); }
                }
            } else {                          // scalar loads
                const float *pf = reinterpret_cast<const float *>(p);
#pragma unroll
                for (int q = 0; q < 128; ++q) s = __fadd_rn(s, pf[q]);
            }
        }
        long long t1 = clock64();
        out[lane] = s;
        if (lane == 0) cyc[0] = t1 - t0;
    } else if (warp <= busy_warps || ((busy_mask >> warp) & 1u)) {          // other warps hammering shared memory with stores meanwhile
        float *w = vals + 28 * nbp;
        if (g) {                              // scattered global loads: `spread` floats between neighbouring lanes
            float acc = 0.f;
            for (int it = 0; it < 40; ++it) {
                float x[9];
#pragma unroll
                for (int q = 0; q < 9; ++q) x[q] = __ldg(g + ((it * 9 + q) * 4099 + warp * 977 + lane * spread) % (1 << 20));
#pragma unroll
                for (int q = 0; q < 9; ++q) { acc += x[q]; w[(q * 67 + warp * 32 + lane) % (3 * nbp)] = x[q]; }
            }
            out[64 + threadIdx.x] = acc;
        } else
        for (int it = 0; it < 400; ++it)
#pragma unroll
            for (int q = 0; q < 9; ++q) w[(q * 67 + warp * 32 + lane + it) % (3 * nbp)] = (float)it;
    }
    __syncthreads();
}
int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 8);
    const int ngroup = 9, nbp = 1156;
    auto run = [&](auto kern, const char *name, int threads, int active, int busy) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int rep = 0; rep < 2; ++rep) { kern<<<1, threads, 32 * nbp * 4>>>(out, cyc, ngroup, nbp, active, busy, 0u, nullptr, 0); cudaDeviceSynchronize(); }
        printf("%-28s block %4d, %2d lanes summing, %2d busy warps: %6lld cycles = %.2f per add (%s)\n", name, threads, active, busy, cyc[0], (double)cyc[0] / (ngroup * 128), cudaGetErrorString(cudaGetLastError()));
    };
    for (int threads : {32, 1024})
        for (int active : {1, 9, 27}) {
            run(probe<0>, "LDS.128 between the adds", threads, active, 0);
            run(probe<1>, "8 x LDS.128, then 32 adds", threads, active, 0);
            run(probe<2>, "scalar LDS", threads, active, 0);
        }
    for (int w = 1; w < 32; ++w) {
        for (int rep = 0; rep < 2; ++rep) { probe<0><<<1, 1024, 32 * nbp * 4>>>(out, cyc, ngroup, nbp, 9, 0, 1u << w, nullptr, 0); cudaDeviceSynchronize(); }
        printf("busy warp %2d alone: %.2f per add\n", w, (double)cyc[0] / (ngroup * 128));
    }
    for (unsigned int m : {0xeeeeeeeeu, 0xfffffff0u, 0xffff0000u, 0xaaaaaaaau, 0xfefefefeu}) {
        for (int rep = 0; rep < 2; ++rep) { probe<0><<<1, 1024, 32 * nbp * 4>>>(out, cyc, ngroup, nbp, 9, 0, m, nullptr, 0); cudaDeviceSynchronize(); }
        printf("busy mask %08x: %.2f per add\n", m, (double)cyc[0] / (ngroup * 128));
    }
    float *gbuf; cudaMalloc(&gbuf, 4 << 20); cudaMemset(gbuf, 0, 4 << 20);
    for (int spread : {1, 8, 32, 256}) {
        for (int rep = 0; rep < 2; ++rep) { probe<0><<<1, 1024, 32 * nbp * 4>>>(out, cyc, ngroup, nbp, 9, 0, 0xeeeeeeeeu, gbuf, spread); cudaDeviceSynchronize(); }
        printf("24 warps gathering (lanes %3d floats apart) on the other schedulers: %.2f per add\n", spread, (double)cyc[0] / (ngroup * 128));
    }
    for (int busy : {3, 12, 24}) run(probe<0>, "LDS.128 between the adds", 1024, 9, busy);
    return 0;
}
