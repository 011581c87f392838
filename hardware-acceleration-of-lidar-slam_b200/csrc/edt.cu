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
// edt_tma_kernel -- one CTA owns a block of 64*NW columns and a chunk of rows:
//   * the int32 occupancy streams through a ring of shared-memory stages with TMA
//     (cp.async.bulk.tensor.2d, one box of R rows x (64*NW + 36) columns per stage, box
//     origin 16-byte aligned left of the strip; out-of-bounds cells are zero-filled by the
//     TMA unit, so the kernel has no load-side bounds checks).  There is no producer
//     warp: the last of the NW warps to finish reading a stage (shared-memory arrival
//     counter) re-arms its mbarrier and issues the refill, so nothing ever spins;
//   * NW consumer warps each own a strip of 64 output columns and march down the chunk.
//     Per row a warp turns three 32-cell words of the stage into ballots; every lane
//     extracts the 2R+2-bit neighbourhood of its two columns with one funnel shift each,
//     and the nearest set bit on either side comes from one BREV + one FLO.  64*h^2 of
//     the lane's two columns is packed as u16x2 and kept in a register window of 3R rows;
//     pass 2 is 2R VIADDMNMX.U16x2 (min(a + imm, c) on both halves, the sm_90+/sm_100 DPX
//     instruction) per output row with immediates 64*dy^2;
//   * 64*d2 is directly the byte offset of a 16-bank-replicated sqrt table in shared
//     memory (entry d2, bank lane%16: at most 2-way conflicts), and the two f32 results go
//     out as two aligned 128-byte warp stores.
// The ALU pipe (VIADDMNMX/SHF/LOP3, one warp instruction per two cycles per SM
// sub-partition) bounds pass 2; HBM traffic is the algorithmic 4 B read + 4 B written per
// cell (halo re-reads hit L2).
#include <cuda.h>

#include <climits>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int EDT_MAX_FUSED_R = 14;          // 2R+2 window bits must fit 32; sums fit u16

// ---- mbarrier / TMA primitives (inline PTX) -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Bounded: try_wait suspends the warp for a hardware-defined slice per attempt; a load that has not
// completed after 2^22 attempts (seconds) never will -> false, and the kernel reports DEV_ERR_TMA.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t tries = 0; tries < (1u << 22); ++tries) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}

constexpr int EDT_SCALE = 64;                // packed u16 values are EDT_SCALE * d2 = LUT byte offset

// 64 * h^2 for the window X: bit k = occupancy of column (c - (R+1) + k), c = this
// lane's output column.  The right-hand side is folded onto the left so that bit (R+1-d)
// is set iff a cell at horizontal distance d (either side) is occupied; the highest set
// bit is the nearest one.  No occupied cell in reach gives h = R+2, which clamps.
template <int R>
__device__ __forceinline__ uint32_t h2scaled(uint32_t X)
{
    constexpr uint32_t MASK = ((1u << (R + 1)) - 1u) << 1;
    const uint32_t M = (X | (__brev(X) >> (29 - 2 * R))) & MASK;
    const int z = __clz(M);                       // h = z - (30 - R)
    const int u = 8 * z + 8 * (R - 30);           // 8h
    return (uint32_t)(u * u);                     // 64 h^2
}

// Pass 2 for the R rows of one batch: window rows r .. r+2R feed output row r.
// CHECK = false is the interior path (no row / column predicates).
// NDEST > 1 (row-sharded transform on several GPUs): every result is also stored at the same
// offset of each peer's field (peer_delta[k] = byte distance from this GPU's field to peer k's,
// mapped through CUDA IPC), so the all-gather of the row blocks happens inside the kernel,
// row by row, over NVLink.
struct EdtPeers {
    long long delta[7];
    int n;                       // number of EXTRA destinations (0 = this GPU only)
};
template <int R>
struct EdtRowMask {                      // one bit per window row (3R rows)
    using type = typename std::conditional<(3 * R > 32), unsigned long long, uint32_t>::type;
};

// SKIP: rowmask has bit i set <=> window row i has an occupied cell within this warp's reach (warp uniform).
// An output row whose 2R+1 window rows are all empty is max_dist everywhere: no arithmetic at all.  When every
// window row is occupied (dense maps) the caller takes the SKIP = false copy, which has no branches.
template <int R, bool CHECK, bool MULTI, bool SKIP>
__device__ __forceinline__ void edt_emit_rows(const uint32_t (&win)[3 * R], typename EdtRowMask<R>::type rowmask,
                                              const unsigned char *lut, uint32_t lane4, uint32_t clampv, float max_dist,
                                              unsigned char *pb, uint32_t pitch_bytes, int rows_left, bool s0, bool s1,
                                              const EdtPeers &peers)
{
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float f0 = max_dist, f1 = max_dist;
        if (!SKIP || ((uint32_t)(rowmask >> r) & ((1u << (2 * R + 1)) - 1u))) {           // warp uniform
            uint32_t a0 = win[r + R];
            uint32_t a1 = 0xffffffffu;
#pragma unroll
            for (int d = 1; d <= R; ++d) {
                const uint32_t k = (uint32_t)(EDT_SCALE * d * d) * 0x00010001u;
                a0 = __viaddmin_u16x2(win[r + R - d], k, a0);
                a1 = __viaddmin_u16x2(win[r + R + d], k, a1);
            }
            const uint32_t a = __vimin3_u16x2(a0, a1, clampv);
            f0 = *reinterpret_cast<const float *>(lut + ((a & 0xffffu) | lane4));
            f1 = *reinterpret_cast<const float *>(lut + (__umulhi(a, 65536u) + lane4));
        }
        float *o = reinterpret_cast<float *>(pb + (uint64_t)pitch_bytes * (uint32_t)r);
        if (CHECK) {
            if (r < rows_left) {
                if (s0) o[0] = f0;
                if (s1) o[32] = f1;
            }
        } else {
            o[0] = f0;
            o[32] = f1;
        }
        if (MULTI) {
            for (int k = 0; k < peers.n; ++k) {
                float *q = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(o) + peers.delta[k]);
                if (!CHECK || r < rows_left) {
                    if (!CHECK || s0) q[0] = f0;
                    if (!CHECK || s1) q[32] = f1;
                }
            }
        }
    }
}

// OT: element type of the occupancy the kernel streams -- the int32 grid (the reference's type), or its byte
// shadow (1 byte per cell: a quarter of the read traffic of an HBM-bound kernel).
template <int R, int NW, typename OT = int32_t>
struct EdtCfg {
    // The TMA box origin is kept 16-byte aligned in global memory, so the stage starts SH
    // columns left of the CTA's first output column (SH = R+1 rounded up to 16 bytes) and the
    // ballot loads skip the DELTA = SH - (R+1) surplus columns.
    static constexpr int EB = (int)sizeof(OT);
    static constexpr int AL = 16 / EB;                       // elements per 16 bytes
    static constexpr int SH = (R + 1 + AL - 1) / AL * AL;
    static constexpr int DELTA = SH - (R + 1);
    // one box per stage (<= 256 columns, a multiple of 16 bytes): the last ballot word ends at 64 NW + 32 + DELTA
    static constexpr int BOX_COLS = EB == 4 ? 64 * NW + 32 + 4 : (64 * NW + 32 + DELTA + AL - 1) / AL * AL;
    static constexpr int STAGE_BYTES = ((R * BOX_COLS * EB) + 127) & ~127;
    static constexpr uint32_t TX_BYTES = R * BOX_COLS * EB;
    static constexpr int THREADS = 32 * NW;
    static_assert(BOX_COLS <= 256 && BOX_COLS >= 64 * NW + 32 + DELTA, "TMA box dimension limit");
};

// Dynamic shared memory: [stage ring][sqrt table][mbarriers][arrival counters]
// Output rows [row_begin, row_end) only (the whole grid, or one rank's block of a row-sharded
// transform); the input halo above and below comes from the full occupancy either way.
template <int R, int NW, int NST, bool MULTI, typename OT = int32_t>
__global__ void __launch_bounds__(EdtCfg<R, NW, OT>::THREADS)
edt_tma_kernel(const __grid_constant__ CUtensorMap tmap, float *__restrict__ out, uint32_t pitch_bytes,
               int row_begin, int row_end, int cols, int chunk_batches, int t2, float max_dist,
               const __grid_constant__ EdtPeers peers, unsigned int *__restrict__ error)
{
    const int rows = row_end;
    using C = EdtCfg<R, NW, OT>;
    constexpr int B = R;                          // rows per stage / batch
    constexpr int WN = 3 * R;                     // register window rows

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *lut = smem_raw + (size_t)NST * C::STAGE_BYTES;
    uint64_t *full = reinterpret_cast<uint64_t *>(lut + (size_t)(t2 + 1) * EDT_SCALE);
    int *arrivals = reinterpret_cast<int *>(full + NST);      // per stage, monotonically increasing

    // Consecutive transforms are independent of each other (each reads an occupancy no transform writes and
    // writes its own field), so the next one in the stream may start -- launch latency, table fill, its first
    // TMA loads -- while this one drains: programmatic dependent launch.  Every thread waits for the grid in
    // front right before it exits, so completion order (what later kernels in the stream rely on) is kept.
    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * (64 * NW);
    const int y0 = row_begin + blockIdx.y * chunk_batches * B;
    const int out_rows = min(chunk_batches * B, rows - y0);
    const int nb_out = (out_rows + B - 1) / B;
    const int nbl = nb_out + 2;                   // stages to stream: halo + chunk + halo
    const int nactive = min(NW, (cols - x0 + 63) / 64);

    for (int d = tid; d <= t2; d += C::THREADS) {
        const float v = d < t2 ? __fsqrt_rn((float)d) : max_dist;
        float4 *p = reinterpret_cast<float4 *>(lut + (size_t)d * EDT_SCALE);
#pragma unroll
        for (int q = 0; q < EDT_SCALE / 16; ++q) p[q] = make_float4(v, v, v, v);
    }
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            arrivals[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // prologue: fill the ring
        for (int t = 0; t < NST && t < nbl; ++t) {
            mbar_expect_tx(&full[t], C::TX_BYTES);
            tma_load_2d(smem_raw + (size_t)t * C::STAGE_BYTES, &tmap, &full[t], x0 - C::SH, y0 - R + t * B);
        }
    }
    __syncthreads();
    if (warp >= nactive) {
        pdl_wait_prior_grids();
        return;
    }

    // ---- consumers: warp j owns output columns [x0 + 64j, x0 + 64j + 64) ---------------
    // Stage column q is grid column x0 - SH + q.  Strip j ballots the three 32-cell words
    // starting at stage column 64j + DELTA, i.e. grid column x0 + 64j - (R+1): bit k of the
    // funnel-shifted window of output column c is then grid column c - (R+1) + k.  (The top
    // lanes of the third word may read past the row; those bits are never used.)
    const int woff = 64 * warp + C::DELTA + lane;
    const int c0 = x0 + 64 * warp + lane, c1 = c0 + 32;
    const bool s0 = c0 < cols, s1 = c1 < cols;
    const bool interior = x0 + 64 * warp + 64 <= cols;
    unsigned char *obase = reinterpret_cast<unsigned char *>(out + c0);
    const uint32_t lane4 = (uint32_t)(lane & 15) * 4u;
    const uint32_t clampv = (uint32_t)(EDT_SCALE * t2) * 0x00010001u;

    // h = R + 2 ("nothing within reach", what h2scaled returns for an empty window): clamps in pass 2
    constexpr uint32_t NONE2 = (uint32_t)(EDT_SCALE * (R + 2) * (R + 2)) * 0x00010001u;
    uint32_t win[WN];
#pragma unroll
    for (int i = 0; i < WN; ++i) win[i] = NONE2;
    typename EdtRowMask<R>::type rowmask = 0;     // bit i: window row i has an occupied cell in this warp's three words

    for (int t = 0; t < nbl; ++t) {
        const int s = t % NST;
        if (!mbar_wait(&full[s], (t / NST) & 1)) {
            if (lane == 0) atomicOr(error, DEV_ERR_TMA);
            pdl_wait_prior_grids();
            return;
        }
        const OT *st = reinterpret_cast<const OT *>(smem_raw + (size_t)s * C::STAGE_BYTES) + woff;
        // ---- pass 1: horizontal nearest-occupied distance from the ballots ------------
        int ld[B][3];
#pragma unroll
        for (int r = 0; r < B; ++r) {
#pragma unroll
            for (int i = 0; i < 3; ++i) ld[r][i] = st[r * C::BOX_COLS + 32 * i];
        }
#pragma unroll
        for (int r = 0; r < B; ++r) {
            const uint32_t w0 = __ballot_sync(0xffffffffu, ld[r][0] != 0);
            const uint32_t w1 = __ballot_sync(0xffffffffu, ld[r][1] != 0);
            const uint32_t w2 = __ballot_sync(0xffffffffu, ld[r][2] != 0);
            // Occupancy grids are sparse (walls): most rows of a 64-column strip hold nothing.  The ballots are
            // warp uniform, so an empty row costs the loads and the three votes and nothing else.
            uint32_t v = NONE2;
            if ((w0 | w1 | w2) != 0u) {
                const uint32_t X0 = __funnelshift_r(w0, w1, lane);
                const uint32_t X1 = __funnelshift_r(w1, w2, lane);
                v = h2scaled<R>(X0) + (h2scaled<R>(X1) << 16);
                rowmask |= (typename EdtRowMask<R>::type)1 << (2 * R + r);
            }
            win[2 * R + r] = v;
        }
        // ---- stage consumed: the last warp to get here refills it ----------------------
        __syncwarp();
        if (lane == 0 && t + NST < nbl) {
            const int old = atomicAdd(&arrivals[s], 1);
            if (old + 1 == nactive * (t / NST + 1)) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&full[s], C::TX_BYTES);
                tma_load_2d(smem_raw + (size_t)s * C::STAGE_BYTES, &tmap, &full[s], x0 - C::SH,
                            y0 - R + (t + NST) * B);
            }
        }
        // ---- pass 2: vertical min-plus with the parabola, two columns per op ----------
        if (t >= 2) {
            const int yb = y0 + (t - 2) * B;
            unsigned char *pb = obase + (size_t)(uint32_t)yb * pitch_bytes;
            constexpr typename EdtRowMask<R>::type FULL = ((typename EdtRowMask<R>::type)1 << (3 * R)) - 1;
            if (interior && yb + B <= rows) {
                if (rowmask == FULL)
                    edt_emit_rows<R, false, MULTI, false>(win, rowmask, lut, lane4, clampv, max_dist, pb, pitch_bytes, 0, true, true, peers);
                else
                    edt_emit_rows<R, false, MULTI, true>(win, rowmask, lut, lane4, clampv, max_dist, pb, pitch_bytes, 0, true, true, peers);
            } else {
                edt_emit_rows<R, true, MULTI, true>(win, rowmask, lut, lane4, clampv, max_dist, pb, pitch_bytes, rows - yb, s0, s1, peers);
            }
        }
        rowmask >>= B;
        // ---- slide the window down by B rows ------------------------------------------
#pragma unroll
        for (int i = 0; i < 2 * R; ++i) win[i] = win[i + B];
    }
    pdl_wait_prior_grids();
}

// ---- generic path for any radius: two plain passes through a u16 intermediate -------
__global__ void edt_generic_cols(const int32_t *__restrict__ occ, long occ_pitch,
                                 uint16_t *__restrict__ g, int rows, int cols, int R)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {          // gridDim.y is capped at 65535
        int best = 0xffff;
        const int lo = max(r - R, 0), hi = min(r + R, rows - 1);
        for (int j = lo; j <= hi; ++j)
            if (__ldg(occ + (long)j * occ_pitch + c)) best = min(best, abs(j - r));
        g[(long)r * cols + c] = (uint16_t)best;
    }
}

__global__ void edt_generic_rows(const uint16_t *__restrict__ g, float *__restrict__ out,
                                 long out_pitch, int rows, int cols, int R, int t2,
                                 float max_dist)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int lo = max(c - R, 0), hi = min(c + R, cols - 1);
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        int best = INT_MAX;
        for (int i = lo; i <= hi; ++i) {
            const int gv = g[(long)r * cols + i];
            if (gv != 0xffff) best = min(best, (i - c) * (i - c) + gv * gv);
        }
        out[(long)r * out_pitch + c] = best < t2 ? __fsqrt_rn((float)best) : max_dist;
    }
}

// ---- host side ----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

template <int R, int NW, int NST, typename OT>
size_t edt_smem_bytes(int t2)
{
    using C = EdtCfg<R, NW, OT>;
    return (size_t)NST * C::STAGE_BYTES + (size_t)(t2 + 1) * EDT_SCALE + NST * (sizeof(uint64_t) + sizeof(int));
}

template <int R, int NW, int NST, bool MULTI, typename OT = int32_t>
int launch_tma(b200slam_ctx *ctx, const OT *d_occ, int occ_pitch, float *d_field, int field_pitch,
               int rows, int cols, int t2, float max_dist, int row_begin, int row_end, const EdtPeers &peers)
{
    using C = EdtCfg<R, NW, OT>;
    auto kern = edt_tma_kernel<R, NW, NST, MULTI, OT>;
    const size_t smem = edt_smem_bytes<R, NW, NST, OT>(t2);
    static int occupancy_dev[64] = {};        // resident CTAs per SM (per instantiation and device)
    static size_t smem_set_dev[64] = {};
    int &occupancy = occupancy_dev[ctx->device & 63];
    size_t &smem_set = smem_set_dev[ctx->device & 63];
    if (smem > smem_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
        occupancy = 0;
    }
    if (occupancy == 0) {
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, kern, C::THREADS, smem));
        if (occupancy < 1) occupancy = 1;
    }

    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return b200slam_set_error(ctx, B200SLAM_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)occ_pitch * sizeof(OT)};
    const cuuint32_t box[2] = {(cuuint32_t)C::BOX_COLS, (cuuint32_t)R};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tmap, sizeof(OT) == 4 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                      const_cast<OT *>(d_occ), gdim, gstr, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      getenv("B200SLAM_EDT_L2P") ? (CUtensorMapL2promotion)atoi(getenv("B200SLAM_EDT_L2P"))
                                                 : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS)
        return b200slam_set_error(ctx, B200SLAM_ERR_CUDA, "cuTensorMapEncodeTiled -> %d (pitch %d)", (int)cr,
                                  occ_pitch);

    // Rows per CTA.  Big grids: chunks of 14 batches -- several waves of CTAs balance better
    // than one wave of long ones (measured at 8192^2: 103 us vs 111 us) and the halo costs
    // 2/16.  Small grids: a chunk of `cb` batches costs cb + ~2.5 batch times (two halo stages
    // that only run pass 1, plus the pipeline fill); pick the cb that minimises waves x cost
    // for the number of CTAs the GPU holds at once.
    const int gx = (cols + 64 * NW - 1) / (64 * NW);
    const int nbatch = (row_end - row_begin + R - 1) / R;
    const long resident = (long)ctx->sm_count * occupancy;
    int best_cb = 1;
    double best_cost = 1e300;
    for (int cb = 1; cb <= 96 && cb <= nbatch; ++cb) {
        const long gy = (nbatch + cb - 1) / cb;
        const long waves = (gx * gy + resident - 1) / resident;
        const double cost = (double)waves * (cb + 2.5);
        if (cost < best_cost - 1e-9) { best_cost = cost; best_cb = cb; }
    }
    if (nbatch >= 14 && (long)gx * ((nbatch + 13) / 14) * 2 >= 5 * resident) best_cb = 14;
    else if (ctx->match_mode == B200SLAM_MATCH_THROUGHPUT) {
        // Many independent transforms in flight (they overlap through programmatic dependent launch): what
        // counts is the work per cell, i.e. the halo overhead (cb + 2) / cb, as long as one launch still
        // spreads over every SM a few times.  Measured at 2048^2, back to back: cb = 3 (the latency choice)
        // 7.6 us, cb = 5 6.6 us; one launch alone: 18.8 us either way.
        for (int cb = 14; cb >= 1; --cb)
            if (cb <= nbatch && (long)gx * ((nbatch + cb - 1) / cb) >= 3L * ctx->sm_count) { best_cb = cb > best_cb ? cb : best_cb; break; }
    }
    if (const char *e = getenv("B200SLAM_EDT_CB")) best_cb = max(1, min(atoi(e), nbatch));   // tuning knob
    const int gy = (nbatch + best_cb - 1) / best_cb;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, gy);
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // Only directly behind another transform, or behind a scan-matching kernel that reads a DIFFERENT field (the
    // serial step EDT(i) -> match(i) -> EDT(i+1): the transform of the next map starts under the match's
    // reduction tail; both kinds of kernel release their dependents at their first instruction and this one
    // waits for the grid in front before it exits).  Never for the multi-GPU variant (its peer stores are
    // bracketed by barrier kernels).
    const bool behind_other_match = ctx->prev_launch_was_lattice && ctx->prev_lattice_field != d_field;
    cfg.numAttrs = (ctx->use_pdl && (ctx->prev_launch_was_edt || behind_other_match) && !MULTI) ? 1 : 0;
    CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kern, tmap, d_field, (uint32_t)field_pitch * 4u, row_begin, row_end, cols, best_cb, t2,
                                     max_dist, peers, &ctx->d_match->error));
    LAUNCH_CHECK(ctx);
    ctx->prev_launch_was_edt = !MULTI;
    return B200SLAM_OK;
}

template <int R>
int launch_fused(b200slam_ctx *ctx, const int32_t *d_occ, int occ_pitch, float *d_field, int field_pitch,
                 int rows, int cols, int t2, float max_dist, int row_begin, int row_end, const EdtPeers &peers)
{
    if (peers.n > 0)
        return launch_tma<R, 3, 3, true>(ctx, d_occ, occ_pitch, d_field, field_pitch, rows, cols, t2, max_dist,
                                         row_begin, row_end, peers);
    if (const char *e = getenv("B200SLAM_EDT_NST")) {
        if (atoi(e) == 2)
            return launch_tma<R, 3, 2, false>(ctx, d_occ, occ_pitch, d_field, field_pitch, rows, cols, t2, max_dist,
                                              row_begin, row_end, peers);
        if (atoi(e) == 4)
            return launch_tma<R, 3, 4, false>(ctx, d_occ, occ_pitch, d_field, field_pitch, rows, cols, t2, max_dist,
                                              row_begin, row_end, peers);
    }
    return launch_tma<R, 3, 3, false>(ctx, d_occ, occ_pitch, d_field, field_pitch, rows, cols, t2, max_dist,
                                      row_begin, row_end, peers);
}

}  // namespace

// Window radius R and clamp index t2 exactly as the float compare at main.c:235 decides.
static void edt_radius(float max_dist, int *R_out, int *t2_out)
{
    const float thr = max_dist * max_dist;
    int R = 0;
    while ((float)((R + 1) * (R + 1)) < thr) R++;
    int t2 = R * R;
    while ((float)t2 < thr) t2++;       // smallest integer d2 that is NOT < max_dist^2
    *R_out = R; *t2_out = t2;
}

namespace {
__global__ void __launch_bounds__(256)
occ_pack_kernel(const int32_t *__restrict__ occ, uint8_t *__restrict__ occ8, int pitch, int rows)
{
    // 4 cells per thread: one 16-byte load, one 4-byte store (the pitch is a multiple of 32 cells)
    const long i4 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= (long)rows * pitch / 4) return;
    const int4 v = reinterpret_cast<const int4 *>(occ)[i4];
    reinterpret_cast<uint32_t *>(occ8)[i4] = (v.x != 0 ? 1u : 0u) | (v.y != 0 ? 0x100u : 0u) | (v.z != 0 ? 0x10000u : 0u) |
                                             (v.w != 0 ? 0x1000000u : 0u);
}
}  // namespace

int occ_pack_launch(b200slam_ctx *ctx, const b200slam_map *map)
{
    const long n4 = (long)map->rows * map->occ_pitch / 4;
    occ_pack_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, ctx->stream>>>(map->d_occ, map->d_occ8, map->occ_pitch, map->rows);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}

int edt_launch_bytes(b200slam_ctx *ctx, const uint8_t *d_occ8, int occ8_pitch, float *d_field,
                     int field_pitch, int rows, int cols, float max_dist)
{
    if (rows <= 0 || cols <= 0) return B200SLAM_OK;
    if (!(max_dist > 0.0f) || max_dist > 255.0f)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "max_dist %g out of (0, 255]", max_dist);
    int R, t2;
    edt_radius(max_dist, &R, &t2);
    EdtPeers peers = {};
    switch (R) {
#define CASE(RR) case RR: return launch_tma<RR, 3, 3, false, uint8_t>(ctx, d_occ8, occ8_pitch, d_field, field_pitch, rows, cols, t2, max_dist, 0, rows, peers);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7)
        CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
#undef CASE
        default: break;
    }
    return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "byte-shadow EDT needs 1 < max_dist <= 15 (radius %d)", R);
}

int edt_launch(b200slam_ctx *ctx, const int32_t *d_occ, int occ_pitch, float *d_field,
               int field_pitch, int rows, int cols, float max_dist)
{
    return edt_launch_rows(ctx, d_occ, occ_pitch, d_field, field_pitch, rows, cols, max_dist, 0, rows, nullptr, 0);
}

int edt_launch_rows(b200slam_ctx *ctx, const int32_t *d_occ, int occ_pitch, float *d_field,
                    int field_pitch, int rows, int cols, float max_dist, int row_begin, int row_end,
                    float *const *peer_fields, int npeers)
{
    if (rows <= 0 || cols <= 0) return B200SLAM_OK;
    if (row_begin < 0 || row_end > rows || npeers < 0 || npeers > 7)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "EDT row range [%d, %d) of %d rows / %d peers", row_begin,
                                  row_end, rows, npeers);
    if (row_end <= row_begin) return B200SLAM_OK;
    EdtPeers peers;
    peers.n = npeers;
    for (int k = 0; k < 7; ++k)
        peers.delta[k] = k < npeers ? (long long)(reinterpret_cast<const char *>(peer_fields[k]) -
                                                  reinterpret_cast<const char *>(d_field))
                                    : 0;
    if (!(max_dist > 0.0f) || max_dist > 255.0f)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "max_dist %g out of (0, 255]", max_dist);
    int R, t2;
    edt_radius(max_dist, &R, &t2);

    // The TMA path needs a 16-byte aligned base and pitch (always true for b200slam_map).
    const bool tma_ok = (occ_pitch % 4 == 0) && ((uintptr_t)d_occ % 16 == 0);
    if (tma_ok) {
        switch (R) {
#define CASE(RR) case RR: return launch_fused<RR>(ctx, d_occ, occ_pitch, d_field, field_pitch, rows, cols, t2, max_dist, row_begin, row_end, peers);
            CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7)
            CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
#undef CASE
            default: break;
        }
    }
    static_assert(EDT_MAX_FUSED_R == 14, "switch above covers 1..14");
    // R == 0 (max_dist <= 1) or R > 14: generic two-pass kernels (whole grid, this GPU only).
    if (row_begin != 0 || row_end != rows || npeers > 0)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "row-sharded EDT needs 1 < max_dist <= 15 (radius %d)", R);
    const size_t need = (size_t)rows * cols;
    if (need > ctx->edt_scratch_cap) {
        if (ctx->d_edt_scratch) cudaFree(ctx->d_edt_scratch);
        ctx->d_edt_scratch = nullptr;
        ctx->edt_scratch_cap = 0;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_edt_scratch, need * sizeof(uint16_t)));
        ctx->edt_scratch_cap = need;
    }
    dim3 block(128), grid((cols + 127) / 128, rows < 65535 ? rows : 65535);
    edt_generic_cols<<<grid, block, 0, ctx->stream>>>(d_occ, occ_pitch, ctx->d_edt_scratch, rows, cols, R);
    LAUNCH_CHECK(ctx);
    edt_generic_rows<<<grid, block, 0, ctx->stream>>>(ctx->d_edt_scratch, d_field, field_pitch, rows,
                                                      cols, R, t2, max_dist);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}
