// score.cu -- candidate-pose / particle scoring against the distance field, sm_100a.
//
// Replaces the scoring loops of FastMatch / FastMatch2 (Subsystem_1/main.c:381-596,
// 598-809).  Per candidate (theta, tx, ty) and beam i the reference computes
//     S_x = psx*ct + psy*st ;  S_y = psx*(-st) + psy*ct          (main.c:462-463)
//     Sx  = (int)roundf(S_x + Sx_temp[tx]) + 1                    (main.c:483)
//     Sy  = (int)roundf(S_y + Sy_temp[ty]) + 1                    (main.c:501)
//     if (1 < Sx < nCols && 1 < Sy < nRows) score += field[Sy-1][Sx-1]   (main.c:512-516)
// with psx = scan.x * ipixel (main.c:418), every product and sum rounded separately and
// the score accumulated sequentially in beam order.  All of that is reproduced exactly:
// __fmul_rn / __fadd_rn (never contracted), cos/sin supplied by the host libm, roundf
// lowered to FADD.RZ + F2I.TRUNC (exhaustively equal to glibc roundf), and ONE thread
// owning each candidate's sum so the additions happen in the reference's order.  Scores
// are therefore bit-identical, and so is the arg-min (strict `<` in loop order == lowest
// linear index among equal scores == min of the packed (score bits, index) key).
//
// lattice_kernel: the reference hoists the rotation per theta, the column index per
// (theta, tx) and the row index per (theta, ty) (main.c:459-503); so does the kernel.  A
// CTA owns one theta and a TXT x TYT tile of (tx, ty).  Per chunk of beams it builds two
// shared-memory tables -- col[beam][tx] (column index, or INVALID) and rowoff[beam][ty]
// (row * pitch, or INVALID) -- and the inner loop per evaluation is
//     off = max(col + rowoff, -1)   (one VIADDMNMX)      v = field[off]  (one LDG)
//     acc += v                      (one FADD)
// field[-1] is a zero pad, so out-of-bounds beams add +0.0f.  Lanes run along tx (0.5 px
// apart at the reference resolutions), so a warp's gather touches one or two 128-byte
// lines of the L1/L2-resident field.
//
// The arg-min lives in one 64-bit cell (atomicMin of the packed key).  The last CTA to
// finish (ticket counter) publishes it, replays the winner and the last candidate to
// produce what FastMatch leaves in FastMatchParameters (bestHits_size of the WINNER,
// bestHits[] of the LAST candidate: main.c:515,557), and resets the cell and the counter
// for the next launch -- so a match is exactly ONE kernel: no memset, no table upload (the
// lattice axis tables travel as kernel parameters), no separate reduction pass.
#include <climits>

#include "common.cuh"

namespace {

constexpr int INVALID_OFF = -(1 << 30);

__device__ __forceinline__ float rot_x(float psx, float psy, float ct, float st)
{
    return __fadd_rn(__fmul_rn(psx, ct), __fmul_rn(psy, st));            // main.c:462
}
__device__ __forceinline__ float rot_y(float psx, float psy, float ct, float st)
{
    return __fadd_rn(__fmul_rn(psx, -st), __fmul_rn(psy, ct));           // main.c:463
}
// 0-based cell index of (int)roundf(v) + 1 - 1, or -1 when the 1-based index fails
// `1 < S < n` (main.c:512).
__device__ __forceinline__ int cell_index(float v, int n)
{
    const int r = (int)roundf(v);
    return (r > 0 && r < n - 1) ? r : -1;
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, s);
        k = o < k ? o : k;
    }
    return k;
}

// ---- completion protocol shared by the lattice and pose-list kernels -------------------
// Every CTA folds its best key into match->work_key and takes a ticket; the CTA holding
// the last ticket sees the final key.
__device__ __forceinline__ bool cta_is_last(MatchDev *match, unsigned long long cta_best, unsigned total_ctas,
                                            int *flag_smem)
{
    if (threadIdx.x == 0) {
        if (cta_best != ~0ull) atomicMin(&match->work_key, cta_best);
        __threadfence();
        const unsigned ticket = atomicAdd(&match->tickets, 1u);
        *flag_smem = (ticket == total_ctas - 1);
    }
    __syncthreads();
    return *flag_smem != 0;
}

// Replays candidate (ct, st, sxt, syt) with the whole CTA: in-bounds field values are
// compacted in beam order into vals[] (main.c:515); returns the count in every thread.
// red: shared scratch of NT/32 + 1 ints.
template <int NT>
__device__ int trace_candidate(const float *__restrict__ field, int pitch, int rows, int cols,
                               const float *__restrict__ scan_x, const float *__restrict__ scan_y, int nbeams,
                               float ipixel, float ct, float st, float sxt, float syt, float *vals, int *red)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int count = 0;
    for (int i0 = 0; i0 < nbeams; i0 += NT) {
        const int i = i0 + tid;
        bool in = false;
        float v = 0.0f;
        if (i < nbeams) {
            const float psx = __fmul_rn(scan_x[i], ipixel);
            const float psy = __fmul_rn(scan_y[i], ipixel);
            const int c = cell_index(__fadd_rn(rot_x(psx, psy, ct, st), sxt), cols);
            const int r = cell_index(__fadd_rn(rot_y(psx, psy, ct, st), syt), rows);
            in = c >= 0 && r >= 0;
            if (in) v = field[(long)r * pitch + c];
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        __syncthreads();                       // red[] free again
        if (lane == 0) red[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
            const int c = red[w];
            before += w < warp ? c : 0;
            total += c;
        }
        if (in && vals) vals[count + before + __popc(m & ((1u << lane) - 1u))] = v;
        count += total;
    }
    return count;
}

constexpr int LATTICE_TABLE_FLOATS = 960;      // ct | st | sxt | syt as kernel parameters
struct LatticeTables {
    float v[LATTICE_TABLE_FLOATS];
};

struct LatticeArgs {
    const float *field;       // [0][0]; field[-1] == 0
    int pitch, rows, cols;
    const float *scan_x, *scan_y;
    int nbeams;
    float ipixel;
    const float *tables;      // device copy of ct|st|sxt|syt, or nullptr: use the parameter block
    int nth, ntx, nty;
    int th_first;             // first theta index covered by blockIdx.z
    long long row_begin, row_end;
    MatchDev *match;
    float *scores;            // optional
    float *hit_values;        // [2][hit_stride]: winner / last candidate
    int hit_stride;
    int cb;                   // beams per shared-memory chunk
    unsigned total_ctas;
};

// TYPT candidates (consecutive ty) per thread, WX warps along tx, WY warps along ty.
// Dynamic shared memory: colT[cb][TXT] | rowT[cb][TYT] | Sx[cb] | Sy[cb]
template <int TYPT, int WX, int WY>
__global__ void __launch_bounds__(32 * WX * WY)
lattice_kernel(const __grid_constant__ LatticeArgs A, const __grid_constant__ LatticeTables T)
{
    constexpr int TXT = 32 * WX;
    constexpr int TYT = TYPT * WY;
    constexpr int NT = 32 * WX * WY;
    extern __shared__ __align__(16) int lat_smem[];
    int *colT = lat_smem;
    int *rowT = colT + A.cb * TXT;
    float *Sx_s = reinterpret_cast<float *>(rowT + A.cb * TYT);
    float *Sy_s = Sx_s + A.cb;
    __shared__ unsigned long long red[WX * WY];
    __shared__ int tail_red[NT / 32 + 1];
    __shared__ int last_flag;

    const float *tab = A.tables ? A.tables : T.v;
    const float *ctT = tab, *stT = tab + A.nth, *sxtT = tab + 2 * A.nth, *sytT = sxtT + A.ntx;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wx = warp % WX, wy = warp / WX;
    const int ith = A.th_first + blockIdx.z;
    const int tx0 = blockIdx.y * TXT, ty0 = blockIdx.x * TYT;

    unsigned long long best = ~0ull;
    // whole tile outside this shard's (theta, tx) row range?  (still takes a ticket)
    const long long r_lo = (long long)ith * A.ntx + tx0;
    const long long r_hi = r_lo + min(TXT, A.ntx - tx0);
    if (r_hi > A.row_begin && r_lo < A.row_end) {
        const float ct = ctT[ith], st = stT[ith];
        const int txl = wx * 32 + lane;          // tile-local tx of this thread
        const int tyl = wy * TYPT;               // first tile-local ty of this thread

        float acc[TYPT];
#pragma unroll
        for (int j = 0; j < TYPT; ++j) acc[j] = 0.0f;            // main.c:507

        for (int c0 = 0; c0 < A.nbeams; c0 += A.cb) {
            const int cb = min(A.cb, A.nbeams - c0);
            __syncthreads();
            for (int i = tid; i < cb; i += NT) {
                const float psx = __fmul_rn(A.scan_x[c0 + i], A.ipixel);     // main.c:418
                const float psy = __fmul_rn(A.scan_y[c0 + i], A.ipixel);     // main.c:419
                Sx_s[i] = rot_x(psx, psy, ct, st);
                Sy_s[i] = rot_y(psx, psy, ct, st);
            }
            __syncthreads();
            for (int e = tid; e < cb * TXT; e += NT) {
                const int i = e / TXT, t = e % TXT;
                int v = INVALID_OFF;
                if (tx0 + t < A.ntx) {
                    const int c = cell_index(__fadd_rn(Sx_s[i], sxtT[tx0 + t]), A.cols);   // :483
                    if (c >= 0) v = c;
                }
                colT[e] = v;
            }
            for (int e = tid; e < cb * TYT; e += NT) {
                const int i = e / TYT, t = e % TYT;
                int v = INVALID_OFF;
                if (ty0 + t < A.nty) {
                    const int r = cell_index(__fadd_rn(Sy_s[i], sytT[ty0 + t]), A.rows);   // :501
                    if (r >= 0) v = r * A.pitch;
                }
                rowT[e] = v;
            }
            __syncthreads();
            const int *cp = colT + txl;
            const int *rp = rowT + tyl;
#pragma unroll 8
            for (int i = 0; i < cb; ++i) {
                const int c = cp[i * TXT];
                int ro[TYPT];
                if constexpr (TYPT % 4 == 0) {
#pragma unroll
                    for (int j = 0; j < TYPT; j += 4) {
                        const int4 q = *reinterpret_cast<const int4 *>(rp + i * TYT + j);
                        ro[j] = q.x; ro[j + 1] = q.y; ro[j + 2] = q.z; ro[j + 3] = q.w;
                    }
                } else if constexpr (TYPT == 2) {
                    const int2 q = *reinterpret_cast<const int2 *>(rp + i * TYT);
                    ro[0] = q.x; ro[1] = q.y;
                } else {
#pragma unroll
                    for (int j = 0; j < TYPT; ++j) ro[j] = rp[i * TYT + j];
                }
                float v[TYPT];
#pragma unroll
                for (int j = 0; j < TYPT; ++j) v[j] = __ldg(A.field + __viaddmax_s32(c, ro[j], -1));
#pragma unroll
                for (int j = 0; j < TYPT; ++j) acc[j] = __fadd_rn(acc[j], v[j]);      // main.c:516
            }
        }

        // ---- arg-min of this tile (lowest score, then lowest linear index) ----------
        const int itx = tx0 + txl;
        const long long row = (long long)ith * A.ntx + itx;
        if (itx < A.ntx && row >= A.row_begin && row < A.row_end) {
#pragma unroll
            for (int j = 0; j < TYPT; ++j) {
                const int ity = ty0 + tyl + j;
                if (ity < A.nty) {
                    const long long lin = row * A.nty + ity;
                    if (A.scores) A.scores[lin] = acc[j];
                    const unsigned long long k = pack_key(acc[j], (unsigned int)lin);
                    best = k < best ? k : best;
                }
            }
        }
    }
    best = warp_min_u64(best);
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (tid == 0)
        for (int w = 1; w < WX * WY; ++w) best = red[w] < best ? red[w] : best;

    // ---- last CTA: publish the winner, trace winner and last candidate, reset -----------
    if (!cta_is_last(A.match, best, A.total_ctas, &last_flag)) return;
    __threadfence();
    const unsigned long long key = *reinterpret_cast<volatile unsigned long long *>(&A.match->work_key);
    int best_hits = 0, last_hits = 0;
    if (key != ~0ull) {
        for (int which = 0; which < 2; ++which) {
            const long long lin = which == 0 ? (long long)(key & 0xffffffffull) : A.row_end * A.nty - 1;
            const int ity = (int)(lin % A.nty);
            const long long row = lin / A.nty;
            const int itx = (int)(row % A.ntx), jth = (int)(row / A.ntx);
            const int n = trace_candidate<NT>(A.field, A.pitch, A.rows, A.cols, A.scan_x, A.scan_y, A.nbeams,
                                              A.ipixel, ctT[jth], stT[jth], sxtT[itx], sytT[ity],
                                              A.hit_values + (size_t)which * A.hit_stride, tail_red);
            if (which == 0) best_hits = n; else last_hits = n;
        }
    }
    if (tid == 0) {
        A.match->key = key;
        A.match->best_hits = best_hits;
        A.match->last_hits = last_hits;
        A.match->work_key = ~0ull;
        A.match->tickets = 0u;
    }
}

// ---- arbitrary pose list (particles): one thread per pose, beams sequential ---------
struct PosesArgs {
    const float *field;
    int pitch, rows, cols;
    const float *scan_x, *scan_y;
    int nbeams;
    float ipixel, min_x, min_y;
    const float *px, *py, *ct, *st;    // [P]
    long long P, index_base;
    float *scores;                     // [P]
    int *hits;                         // [P]
    MatchDev *match;
    unsigned total_ctas;
};

constexpr int POSES_THREADS = 128;
constexpr int POSES_CB = 1024;

__global__ void __launch_bounds__(POSES_THREADS) poses_kernel(const __grid_constant__ PosesArgs A)
{
    __shared__ float2 ps[POSES_CB];
    __shared__ unsigned long long red[POSES_THREADS / 32];
    __shared__ int last_flag;
    const long long p = (long long)blockIdx.x * POSES_THREADS + threadIdx.x;
    const bool live = p < A.P;
    float ct = 1.0f, st = 0.0f, sxt = 0.0f, syt = 0.0f;
    if (live) {
        ct = A.ct[p];
        st = A.st[p];
        sxt = __fmul_rn(__fsub_rn(A.px[p], A.min_x), A.ipixel);          // main.c:436
        syt = __fmul_rn(__fsub_rn(A.py[p], A.min_y), A.ipixel);          // main.c:437
    }
    const float nst = -st;
    float score = 0.0f;
    int nh = 0;
    for (int c0 = 0; c0 < A.nbeams; c0 += POSES_CB) {
        const int cb = min(POSES_CB, A.nbeams - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cb; i += POSES_THREADS)
            ps[i] = make_float2(__fmul_rn(A.scan_x[c0 + i], A.ipixel),
                                __fmul_rn(A.scan_y[c0 + i], A.ipixel));
        __syncthreads();
        if (live) {
#pragma unroll 8
            for (int i = 0; i < cb; ++i) {
                const float2 q = ps[i];
                const float fx = __fadd_rn(__fadd_rn(__fmul_rn(q.x, ct), __fmul_rn(q.y, st)), sxt);
                const float fy = __fadd_rn(__fadd_rn(__fmul_rn(q.x, nst), __fmul_rn(q.y, ct)), syt);
                const int c = cell_index(fx, A.cols);
                const int r = cell_index(fy, A.rows);
                const bool in = (c >= 0) && (r >= 0);
                const int off = in ? r * A.pitch + c : -1;
                score = __fadd_rn(score, __ldg(A.field + off));
                nh += in ? 1 : 0;
            }
        }
    }
    unsigned long long best = ~0ull;
    if (live) {
        A.scores[p] = score;
        A.hits[p] = nh;
        best = pack_key(score, (unsigned int)(p + A.index_base));
    }
    best = warp_min_u64(best);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < POSES_THREADS / 32; ++w) best = red[w] < best ? red[w] : best;

    if (!cta_is_last(A.match, best, A.total_ctas, &last_flag)) return;
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long key = *reinterpret_cast<volatile unsigned long long *>(&A.match->work_key);
        int bh = 0, lh = 0;
        if (key != ~0ull) {
            const long long local = (long long)(key & 0xffffffffull) - A.index_base;
            bh = *reinterpret_cast<volatile int *>(&A.hits[local]);
            lh = *reinterpret_cast<volatile int *>(&A.hits[A.P - 1]);
        }
        A.match->key = key;
        A.match->best_hits = bh;
        A.match->last_hits = lh;
        A.match->work_key = ~0ull;
        A.match->tickets = 0u;
    }
}

template <int TYPT, int WX, int WY>
int launch_lattice_cfg(b200slam_ctx *ctx, LatticeArgs &A, const LatticeTables &T, int nth_cover)
{
    constexpr int TXT = 32 * WX, TYT = TYPT * WY;
    auto kern = lattice_kernel<TYPT, WX, WY>;
    // Beams per chunk: the whole scan when it fits ~60 KB of tables, else even chunks.
    const int per_beam = (TXT + TYT + 2) * 4;
    int cb = A.nbeams > 0 ? A.nbeams : 1;
    const int cap = (60 * 1024) / per_beam;
    if (cb > cap) {
        const int chunks = (cb + cap - 1) / cap;
        cb = (cb + chunks - 1) / chunks;
    }
    cb = (cb + 3) & ~3;                                   // keeps the int4 row loads aligned
    A.cb = cb;
    const size_t smem = (size_t)cb * per_beam;
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        smem_set = 64 * 1024;
    }
    dim3 grid((A.nty + TYT - 1) / TYT, (A.ntx + TXT - 1) / TXT, nth_cover);
    A.total_ctas = grid.x * grid.y * grid.z;
    kern<<<grid, 32 * WX * WY, smem, ctx->stream>>>(A, T);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}

}  // namespace

int lattice_launch(b200slam_ctx *ctx, const LatticeLaunch &L)
{
    const b200slam_map *m = L.map;
    LatticeArgs A;
    A.field = m->d_field; A.pitch = m->field_pitch; A.rows = m->rows; A.cols = m->cols;
    A.scan_x = ctx->d_scan_x; A.scan_y = ctx->d_scan_y; A.nbeams = ctx->nbeams;
    A.ipixel = 1 / m->pixel_size;                                        // main.c:383
    A.nth = L.nth; A.ntx = L.ntx; A.nty = L.nty;
    A.row_begin = L.row_begin; A.row_end = L.row_end;
    A.match = ctx->d_match;
    A.scores = L.d_scores;
    A.hit_values = ctx->d_hit_values;
    A.hit_stride = ctx->scan_cap;
    A.tables = L.d_tables;
    LatticeTables T;                     // parameter block (copied at launch)
    if (!L.d_tables) memcpy(T.v, L.h_tables, sizeof(float) * (2 * (size_t)L.nth + L.ntx + L.nty));
    if (L.row_end <= L.row_begin) {
        // empty shard: publish "nothing scored"
        CUDA_TRY(ctx, cudaMemsetAsync(&ctx->d_match->key, 0xff, sizeof(unsigned long long), ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(&ctx->d_match->best_hits, 0, 2 * sizeof(int), ctx->stream));
        return B200SLAM_OK;
    }
    const int th_first = (int)(L.row_begin / L.ntx);
    const int th_last = (int)((L.row_end - 1) / L.ntx);
    A.th_first = th_first;
    const int nth_cover = th_last - th_first + 1;
    const long long cands = (long long)(L.row_end - L.row_begin) * L.nty;

    // Candidates per thread: as many as still leave ~16 warps per SM; big register tiles
    // amortise the per-chunk tables on large sweeps, small lattices need every thread.
    const long long want_warps = 16LL * ctx->sm_count;
    if (cands >= want_warps * 32 * 16 && L.nty >= 64 && L.ntx >= 64)
        return launch_lattice_cfg<16, 2, 4>(ctx, A, T, nth_cover);       // 64 x 64 tile, 256 thr
    if (cands >= want_warps * 32 * 8 && L.nty >= 64)
        return launch_lattice_cfg<8, 1, 8>(ctx, A, T, nth_cover);        // 32 x 64 tile, 256 thr
    if (cands >= want_warps * 32 * 4 && L.nty >= 32)
        return launch_lattice_cfg<4, 1, 8>(ctx, A, T, nth_cover);        // 32 x 32 tile
    if (cands >= want_warps * 32 * 2 && L.nty >= 16)
        return launch_lattice_cfg<2, 1, 8>(ctx, A, T, nth_cover);        // 32 x 16 tile
    if (L.nty > 4)
        return launch_lattice_cfg<1, 1, 8>(ctx, A, T, nth_cover);        // 32 x 8 tile
    return launch_lattice_cfg<1, 1, 4>(ctx, A, T, nth_cover);            // 32 x 4 tile, 128 thr
}

int poses_launch(b200slam_ctx *ctx, const b200slam_map *m, int64_t P, int64_t index_base,
                 float *d_scores, int32_t *d_hits)
{
    PosesArgs A;
    A.field = m->d_field; A.pitch = m->field_pitch; A.rows = m->rows; A.cols = m->cols;
    A.scan_x = ctx->d_scan_x; A.scan_y = ctx->d_scan_y; A.nbeams = ctx->nbeams;
    A.ipixel = 1 / m->pixel_size;
    A.min_x = m->top_left_x; A.min_y = m->top_left_y;
    A.px = ctx->d_pose_soa; A.py = A.px + ctx->pose_cap; A.ct = A.py + ctx->pose_cap;
    A.st = A.ct + ctx->pose_cap;
    A.P = P; A.index_base = index_base;
    A.scores = d_scores; A.hits = d_hits;
    A.match = ctx->d_match;
    if (P <= 0) {
        CUDA_TRY(ctx, cudaMemsetAsync(&ctx->d_match->key, 0xff, sizeof(unsigned long long), ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(&ctx->d_match->best_hits, 0, 2 * sizeof(int), ctx->stream));
        return B200SLAM_OK;
    }
    const unsigned grid = (unsigned)((P + POSES_THREADS - 1) / POSES_THREADS);
    A.total_ctas = grid;
    poses_kernel<<<grid, POSES_THREADS, 0, ctx->stream>>>(A);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}
