// Development probe: minimal TMA / bulk-copy loads, to isolate descriptor / PTX issues.
// usage: tma_probe <experiment>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait0(uint64_t *bar)
{
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n"
                 ::"r"(smem_u32(bar)), "r"(0) : "memory");
}
__device__ __forceinline__ void init_bar(uint64_t *bar)
{
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
}
// descriptor as __grid_constant__ param
__global__ void probe_param(const __grid_constant__ CUtensorMap tmap, int *out, int n, int x, int y)
{
    extern __shared__ __align__(128) unsigned char smem[];
    int *dst = reinterpret_cast<int *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + n * 4);
    init_bar(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(&tmap), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
    }
    wait0(bar);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
}
// descriptor in global memory
__global__ void probe_gmem(const CUtensorMap *tmap, int *out, int n, int x, int y)
{
    extern __shared__ __align__(128) unsigned char smem[];
    int *dst = reinterpret_cast<int *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + n * 4);
    init_bar(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
    }
    wait0(bar);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
}
// plain 1-D bulk copy, no descriptor
__global__ void probe_bulk(const int *src, int *out, int n)
{
    extern __shared__ __align__(128) unsigned char smem[];
    int *dst = reinterpret_cast<int *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + n * 4);
    init_bar(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(n * 4), "r"(smem_u32(bar)) : "memory");
    }
    wait0(bar);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv)
{
    const int exp = argc > 1 ? atoi(argv[1]) : 0;
    const int rows = 100, cols = 320, pitch = 320;
    std::vector<int> h(rows * pitch);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < pitch; ++c) h[r * pitch + c] = r * 1000 + c;
    int *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, 1 << 20);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    int bx = 32, by = 4, x = 0, y = 0;
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_INT32;
    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (exp == 2) l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (exp == 3) dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (exp == 4) { x = -8; y = -2; }
    if (exp == 5) { bx = 192; by = 9; x = -10; y = -9; }
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t gstr[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {(cuuint32_t)bx, (cuuint32_t)by}; cuuint32_t es[2] = {1, 1};
    CUresult cr = enc(&tm, dt, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("exp %d: box %dx%d at (%d,%d) encode=%d\n", exp, bx, by, x, y, (int)cr);
    const unsigned char *tb = (const unsigned char *)&tm;
    for (int i = 0; i < 128; ++i) printf("%02x%s", tb[i], (i % 32 == 31) ? "\n" : "");
    const int n = bx * by;
    if (exp == 0) probe_bulk<<<1, 128, n * 4 + 64>>>(d, o, n);
    else if (exp == 1) {
        CUtensorMap *dtm; cudaMalloc(&dtm, 128); cudaMemcpy(dtm, &tm, 128, cudaMemcpyHostToDevice);
        probe_gmem<<<1, 128, n * 4 + 64>>>(dtm, o, n, x, y);
    } else probe_param<<<1, 128, n * 4 + 64>>>(tm, o, n, x, y);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  -> %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<int> r(n); cudaMemcpy(r.data(), o, r.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < by; ++j) for (int i = 0; i < bx; ++i) {
        int gx = x + i, gy = y + j; int want = (gx < 0 || gx >= cols || gy < 0 || gy >= rows) ? 0 : gy * 1000 + gx;
        if (exp == 0) want = (j * bx + i) / pitch * 1000 + (j * bx + i) % pitch;
        if (r[j * bx + i] != want) bad++;
    }
    printf("  mismatches: %d\n", bad);
    return 0;
}
