// edt.cu -- clamped exact Euclidean distance transform of an occupancy grid, sm_100a.
//
// Replaces euclidean_distance_transform{,2} of the reference
// (Subsystem_1/main.c:223-269, Subsystem_1/main_accelerated.c:215-283,
//  Submodule_2/Accelereated_Euclidean_Distance_Transform.c:1-69):
//     out[r][c] = d2min < max_dist^2 ? sqrtf(d2min) : max_dist
// where d2min is the INTEGER squared distance (main.c:216-220) to the nearest occupied
// cell.  The result only depends on d2min, and only cells within R = ceil(max_dist) - 1
// of a cell can produce d2min < max_dist^2, so the transform is computed exactly as
//     pass 1 (along x):  h(x,y)  = distance to the nearest occupied cell of row y in
//                                  [x-R, x+R]                     (bit tricks on ballots)
//     pass 2 (along y):  d2(x,y) = min_{|dy|<=R} dy^2 + h(x,y+dy)^2   (packed min-plus)
// followed by a table lookup d2 -> sqrtf(d2) (IEEE sqrt, so bit-identical to the host).
//
// edt_fused_kernel: one warp owns a strip of 64 output columns and marches down a chunk
// of rows.  Per row it reads three aligned 128-byte lines of int32 occupancy (the only
// global reads), turns them into three 32-bit ballots, and every lane extracts the
// 2R+1-bit neighbourhood of its two columns with one funnel shift each; the nearest set
// bit on either side comes from one BREV + one FLO.  The 4*h^2 values of the lane's two
// columns are packed as u16x2 and kept in a register window of 4R rows; pass 2 is then
// 2R+1 VIADDMNMX.U16x2 (min(a + imm, c) on both halves, a native sm_90+/sm_100 DPX
// instruction) per output row pair with immediates 4*dy^2.  No shared-memory staging of
// the grid, no intermediate in global memory: algorithmic traffic is 4 B read + 4 B
// written per cell, which is the HBM roofline the kernel is measured against.
#include "common.cuh"

namespace {

constexpr int EDT_WARPS = 4;                 // warps (column strips) per CTA
constexpr int EDT_THREADS = EDT_WARPS * 32;
constexpr int EDT_MAX_FUSED_R = 14;          // 2R+2 window bits must fit 32
constexpr int EDT_LUT_MAX = 232;             // >= (R+1)^2 + 1 for R = 14

template <int R>
__device__ __forceinline__ uint32_t h2x4(uint32_t X)
{
    // X bit k = occupancy of column (c - (R+1) + k), c = this lane's output column.
    // Fold the right-hand side onto the left so that bit (R+1-d) is set iff a cell at
    // horizontal distance d (either side) is occupied; the highest set bit is the
    // nearest one.  Returns (2h)^2 = 4h^2; no occupied cell in reach gives (2R+4)^2,
    // which clamps.
    constexpr uint32_t MASK = ((1u << (R + 1)) - 1u) << 1;
    const uint32_t M = (X | (__brev(X) >> (29 - 2 * R))) & MASK;
    const int t = 2 * __clz(M) + (2 * R - 60);
    return (uint32_t)(t * t);
}

template <int R>
__global__ void __launch_bounds__(EDT_THREADS)
edt_fused_kernel(const int32_t *__restrict__ occ, long occ_pitch, float *__restrict__ out,
                 long out_pitch, int rows, int cols, int chunk_rows, int nstrips, int t2,
                 float max_dist)
{
    constexpr int B = 2 * R;       // rows produced per batch
    constexpr int WN = 4 * R;      // window rows held in registers

    __shared__ float lut[EDT_LUT_MAX];
    for (int d = threadIdx.x; d <= t2; d += EDT_THREADS)
        lut[d] = d < t2 ? __fsqrt_rn((float)d) : max_dist;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int strip = blockIdx.x * EDT_WARPS + (threadIdx.x >> 5);
    if (strip >= nstrips) return;
    const int y0 = blockIdx.y * chunk_rows;
    const int y1 = min(y0 + chunk_rows, rows);

    // Ballot words start at column b (a multiple of 32, so every load instruction reads
    // one aligned 128-byte line); this lane outputs columns b+lane+R+1 and +32.
    const int b = strip * 64 - 32;
    const int cl0 = b + lane, cl1 = cl0 + 32, cl2 = cl0 + 64;
    const bool v0 = cl0 >= 0 && cl0 < cols, v1 = cl1 < cols, v2 = cl2 < cols;
    const int co0 = cl0 + R + 1, co1 = co0 + 32;
    const bool s0 = co0 >= 0 && co0 < cols, s1 = co1 < cols;
    const uint32_t clampv = (uint32_t)(4 * t2) * 0x00010001u;

    uint32_t win[WN];
#pragma unroll
    for (int i = 0; i < WN; ++i) win[i] = 0;

    for (int yb = y0 - B; yb < y1; yb += B) {
        // ---- global loads: rows yb+R .. yb+R+B-1, three aligned lines each ----------
        int ld[B][3];
#pragma unroll
        for (int r = 0; r < B; ++r) {
            const int row = yb + R + r;
            const bool rowok = row >= 0 && row < rows;
            const int32_t *p = occ + (long)row * occ_pitch;
            ld[r][0] = (rowok && v0) ? __ldg(p + cl0) : 0;
            ld[r][1] = (rowok && v1) ? __ldg(p + cl1) : 0;
            ld[r][2] = (rowok && v2 && lane < 2 * R + 2) ? __ldg(p + cl2) : 0;
        }
        // ---- pass 1: horizontal nearest-occupied distance from the ballots ----------
#pragma unroll
        for (int r = 0; r < B; ++r) {
            const uint32_t w0 = __ballot_sync(0xffffffffu, ld[r][0] != 0);
            const uint32_t w1 = __ballot_sync(0xffffffffu, ld[r][1] != 0);
            const uint32_t w2 = __ballot_sync(0xffffffffu, ld[r][2] != 0);
            const uint32_t X0 = __funnelshift_r(w0, w1, lane);
            const uint32_t X1 = __funnelshift_r(w1, w2, lane);
            win[B + r] = h2x4<R>(X0) | (h2x4<R>(X1) << 16);
        }
        // ---- pass 2: vertical min-plus with the parabola, two columns per op --------
        if (yb >= y0) {
#pragma unroll
            for (int r = 0; r < B; ++r) {
                const int y = yb + r;
                if (y < y1) {
                    uint32_t a0 = win[r + R];
                    uint32_t a1 = 0xffffffffu;
#pragma unroll
                    for (int d = 1; d <= R; ++d) {
                        const uint32_t k = (uint32_t)(4 * d * d) * 0x00010001u;
                        a0 = __viaddmin_u16x2(win[r + R - d], k, a0);
                        a1 = __viaddmin_u16x2(win[r + R + d], k, a1);
                    }
                    const uint32_t a = __vminu2(__vminu2(a0, a1), clampv);
                    const float f0 = *(const float *)((const char *)lut + (a & 0xffffu));
                    const float f1 = *(const float *)((const char *)lut + (a >> 16));
                    float *o = out + (long)y * out_pitch;
                    if (s0) o[co0] = f0;
                    if (s1) o[co1] = f1;
                }
            }
        }
        // ---- slide the window down by B rows ---------------------------------------
#pragma unroll
        for (int i = 0; i < B; ++i) win[i] = win[i + B];
    }
}

// ---- generic path for any radius: two plain passes through a u16 intermediate -------
__global__ void edt_generic_cols(const int32_t *__restrict__ occ, long occ_pitch,
                                 uint16_t *__restrict__ g, int rows, int cols, int R)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= cols) return;
    int best = 0xffff;
    const int lo = max(r - R, 0), hi = min(r + R, rows - 1);
    for (int j = lo; j <= hi; ++j)
        if (__ldg(occ + (long)j * occ_pitch + c)) best = min(best, abs(j - r));
    g[(long)r * cols + c] = (uint16_t)best;
}

__global__ void edt_generic_rows(const uint16_t *__restrict__ g, float *__restrict__ out,
                                 long out_pitch, int rows, int cols, int R, int t2,
                                 float max_dist)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= cols) return;
    int best = INT_MAX;
    const int lo = max(c - R, 0), hi = min(c + R, cols - 1);
    for (int i = lo; i <= hi; ++i) {
        const int gv = g[(long)r * cols + i];
        if (gv != 0xffff) best = min(best, (i - c) * (i - c) + gv * gv);
    }
    out[(long)r * out_pitch + c] = best < t2 ? __fsqrt_rn((float)best) : max_dist;
}

template <int R>
int launch_fused(b200slam_ctx *ctx, const int32_t *d_occ, int occ_pitch, float *d_field,
                 int field_pitch, int rows, int cols, int t2, float max_dist)
{
    constexpr int B = 2 * R;
    const int nstrips = (cols + 30 - R) / 64 + 1;
    const int gx = (nstrips + EDT_WARPS - 1) / EDT_WARPS;
    // Rows per chunk: a multiple of the batch height; tall chunks amortise the 2R halo
    // rows, short ones give small grids enough warps (aim for >= 16 per SM).
    const int want_warps = ctx->sm_count * 16;
    int chunks = (want_warps + nstrips - 1) / nstrips;
    int chunk_rows = (rows + chunks - 1) / chunks;
    chunk_rows = ((chunk_rows + B - 1) / B) * B;
    if (chunk_rows < 2 * B) chunk_rows = 2 * B;
    if (chunk_rows > 8 * B) chunk_rows = 8 * B;
    const int gy = (rows + chunk_rows - 1) / chunk_rows;
    dim3 grid(gx, gy);
    edt_fused_kernel<R><<<grid, EDT_THREADS, 0, ctx->stream>>>(
        d_occ, occ_pitch, d_field, field_pitch, rows, cols, chunk_rows, nstrips, t2, max_dist);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}

}  // namespace

int edt_launch(b200slam_ctx *ctx, const int32_t *d_occ, int occ_pitch, float *d_field,
               int field_pitch, int rows, int cols, float max_dist)
{
    if (rows <= 0 || cols <= 0) return B200SLAM_OK;
    if (!(max_dist > 0.0f) || max_dist > 255.0f)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "max_dist %g out of (0, 255]", max_dist);
    // Window radius and clamp index exactly as the float compare at main.c:235 decides.
    const float thr = max_dist * max_dist;
    int R = 0;
    while ((float)((R + 1) * (R + 1)) < thr) R++;
    int t2 = R * R;
    while ((float)t2 < thr) t2++;       // smallest integer d2 that is NOT < max_dist^2

    switch (R) {
#define CASE(RR) case RR: return launch_fused<RR>(ctx, d_occ, occ_pitch, d_field, field_pitch, rows, cols, t2, max_dist);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7)
        CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
#undef CASE
        default: break;
    }
    // R == 0 (max_dist <= 1) or R > 14: generic two-pass kernels.
    const size_t need = (size_t)rows * cols;
    if (need > ctx->edt_scratch_cap) {
        if (ctx->d_edt_scratch) cudaFree(ctx->d_edt_scratch);
        ctx->d_edt_scratch = nullptr;
        ctx->edt_scratch_cap = 0;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_edt_scratch, need * sizeof(uint16_t)));
        ctx->edt_scratch_cap = need;
    }
    dim3 block(128), grid((cols + 127) / 128, rows);
    edt_generic_cols<<<grid, block, 0, ctx->stream>>>(d_occ, occ_pitch, ctx->d_edt_scratch, rows, cols, R);
    LAUNCH_CHECK(ctx);
    edt_generic_rows<<<grid, block, 0, ctx->stream>>>(ctx->d_edt_scratch, d_field, field_pitch, rows,
                                                      cols, R, t2, max_dist);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}
