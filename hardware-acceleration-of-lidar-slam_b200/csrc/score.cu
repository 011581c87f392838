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
// small shared-memory tables -- col[beam][tx] (column index, or INVALID) and
// rowoff[beam][ty] (row * pitch, or INVALID) -- and the inner loop per evaluation is
//     off = max(col + rowoff, -1)   (one VIADDMNMX)      v = field[off]  (one LDG)
//     acc += v                      (one FADD)
// field[-1] is a zero pad, so out-of-bounds beams add +0.0f.  Lanes run along tx (0.5 px
// apart at the reference resolutions), so a warp's gather touches one or two 128-byte
// lines of the L1/L2-resident field.
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

struct LatticeArgs {
    const float *field;       // [0][0]; field[-1] == 0
    int pitch, rows, cols;
    const float *scan_x, *scan_y;
    int nbeams;
    float ipixel;
    const float *ct, *st, *sxt, *syt;
    int nth, ntx, nty;
    int th_first;             // first theta index covered by blockIdx.z
    long long row_begin, row_end;
    unsigned long long *best_key;
    float *scores;            // optional
};

// TYPT candidates (consecutive ty) per thread, WX warps along tx, WY warps along ty.
template <int TYPT, int WX, int WY, int CB>
__global__ void __launch_bounds__(32 * WX * WY)
lattice_kernel(const LatticeArgs A)
{
    constexpr int TXT = 32 * WX;
    constexpr int TYT = TYPT * WY;
    constexpr int NT = 32 * WX * WY;
    __shared__ int colT[CB][TXT];
    __shared__ __align__(16) int rowT[CB][TYT];
    __shared__ float Sx_s[CB], Sy_s[CB];
    __shared__ unsigned long long red[WX * WY];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wx = warp % WX, wy = warp / WX;
    const int ith = A.th_first + blockIdx.z;
    const int tx0 = blockIdx.y * TXT, ty0 = blockIdx.x * TYT;

    // whole tile outside this shard's (theta, tx) row range?
    const long long r_lo = (long long)ith * A.ntx + tx0;
    const long long r_hi = r_lo + min(TXT, A.ntx - tx0);
    if (r_hi <= A.row_begin || r_lo >= A.row_end) return;

    const float ct = A.ct[ith], st = A.st[ith];
    const int txl = wx * 32 + lane;          // tile-local tx of this thread
    const int tyl = wy * TYPT;               // first tile-local ty of this thread

    float acc[TYPT];
#pragma unroll
    for (int j = 0; j < TYPT; ++j) acc[j] = 0.0f;            // main.c:507

    for (int c0 = 0; c0 < A.nbeams; c0 += CB) {
        const int cb = min(CB, A.nbeams - c0);
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
                const int c = cell_index(__fadd_rn(Sx_s[i], A.sxt[tx0 + t]), A.cols);   // :483
                if (c >= 0) v = c;
            }
            colT[i][t] = v;
        }
        for (int e = tid; e < cb * TYT; e += NT) {
            const int i = e / TYT, t = e % TYT;
            int v = INVALID_OFF;
            if (ty0 + t < A.nty) {
                const int r = cell_index(__fadd_rn(Sy_s[i], A.syt[ty0 + t]), A.rows);   // :501
                if (r >= 0) v = r * A.pitch;
            }
            rowT[i][t] = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int i = 0; i < cb; ++i) {
            const int c = colT[i][txl];
            int ro[TYPT];
            if constexpr (TYPT % 4 == 0) {
#pragma unroll
                for (int j = 0; j < TYPT; j += 4) {
                    const int4 q = *reinterpret_cast<const int4 *>(&rowT[i][tyl + j]);
                    ro[j] = q.x; ro[j + 1] = q.y; ro[j + 2] = q.z; ro[j + 3] = q.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < TYPT; ++j) ro[j] = rowT[i][tyl + j];
            }
            float v[TYPT];
#pragma unroll
            for (int j = 0; j < TYPT; ++j) v[j] = __ldg(A.field + __viaddmax_s32(c, ro[j], -1));
#pragma unroll
            for (int j = 0; j < TYPT; ++j) acc[j] = __fadd_rn(acc[j], v[j]);      // main.c:516
        }
    }

    // ---- arg-min of this tile (lowest score, then lowest linear index) --------------
    unsigned long long best = ~0ull;
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
    best = warp_min_u64(best);
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < WX * WY; ++w) best = red[w] < best ? red[w] : best;
        if (best != ~0ull) atomicMin(A.best_key, best);
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
    int *hits;                         // optional [P]
    unsigned long long *best_key;
};

constexpr int POSES_THREADS = 128;
constexpr int POSES_CB = 1024;

__global__ void __launch_bounds__(POSES_THREADS) poses_kernel(const PosesArgs A)
{
    __shared__ float2 ps[POSES_CB];
    __shared__ unsigned long long red[POSES_THREADS / 32];
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
#pragma unroll 4
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
        if (A.hits) A.hits[p] = nh;
        best = pack_key(score, (unsigned int)(p + A.index_base));
    }
    best = warp_min_u64(best);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < POSES_THREADS / 32; ++w) best = red[w] < best ? red[w] : best;
        if (best != ~0ull) atomicMin(A.best_key, best);
    }
}

// ---- trace: hit count / hit values of the winner and of the last candidate ----------
// FastMatch leaves bestHits_size of the WINNER and bestHits[] of the LAST candidate
// (main.c:515,557; SURVEY.md section 7 hard part 4).  Block 0 traces the winner (decoded
// from the packed key), block 1 the last candidate of the row range; one warp each,
// compaction in beam order through ballots.
struct TraceArgs {
    const float *field;
    int pitch, rows, cols;
    const float *scan_x, *scan_y;
    int nbeams;
    float ipixel;
    const float *ct, *st, *sxt, *syt;
    int nth, ntx, nty;
    long long last_lin;
    const unsigned long long *keys;    // nkeys packed keys (1, or one per rank)
    int nkeys;
    MatchDev *out;
    float *hit_values;                 // [2][hit_stride]
    int hit_stride;
};

__global__ void __launch_bounds__(32) trace_kernel(const TraceArgs A)
{
    const int lane = threadIdx.x;
    unsigned long long key = ~0ull;
    for (int k = 0; k < A.nkeys; ++k) key = A.keys[k] < key ? A.keys[k] : key;
    long long lin = blockIdx.x == 0 ? (long long)(key & 0xffffffffull) : A.last_lin;
    if (blockIdx.x == 0 && key == ~0ull) {           // empty shard: nothing was scored
        if (lane == 0) { A.out->key = key; A.out->best_hits = 0; }
        return;
    }
    const int ity = (int)(lin % A.nty);
    const long long row = lin / A.nty;
    const int itx = (int)(row % A.ntx), ith = (int)(row / A.ntx);
    const float ct = A.ct[ith], st = A.st[ith], sxt = A.sxt[itx], syt = A.syt[ity];
    float *vals = A.hit_values + (size_t)blockIdx.x * A.hit_stride;
    int count = 0;
    for (int i0 = 0; i0 < A.nbeams; i0 += 32) {
        const int i = i0 + lane;
        bool in = false;
        float v = 0.0f;
        if (i < A.nbeams) {
            const float psx = __fmul_rn(A.scan_x[i], A.ipixel);
            const float psy = __fmul_rn(A.scan_y[i], A.ipixel);
            const int c = cell_index(__fadd_rn(rot_x(psx, psy, ct, st), sxt), A.cols);
            const int r = cell_index(__fadd_rn(rot_y(psx, psy, ct, st), syt), A.rows);
            in = c >= 0 && r >= 0;
            if (in) v = A.field[(long)r * A.pitch + c];
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (in) vals[count + __popc(m & ((1u << lane) - 1u))] = v;      // main.c:515
        count += __popc(m);
    }
    if (lane == 0) {
        if (blockIdx.x == 0) { A.out->key = key; A.out->best_hits = count; }
        else A.out->last_hits = count;
    }
}

template <int TYPT, int WX, int WY, int CB>
int launch_lattice_cfg(b200slam_ctx *ctx, const LatticeArgs &A, int nth_cover)
{
    constexpr int TXT = 32 * WX, TYT = TYPT * WY;
    dim3 grid((A.nty + TYT - 1) / TYT, (A.ntx + TXT - 1) / TXT, nth_cover);
    lattice_kernel<TYPT, WX, WY, CB><<<grid, 32 * WX * WY, 0, ctx->stream>>>(A);
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
    A.ct = L.d_ct; A.st = L.d_st; A.sxt = L.d_sxt; A.syt = L.d_syt;
    A.nth = L.nth; A.ntx = L.ntx; A.nty = L.nty;
    A.row_begin = L.row_begin; A.row_end = L.row_end;
    A.best_key = &ctx->d_match->key;
    A.scores = L.d_scores;
    CUDA_TRY(ctx, cudaMemsetAsync(&ctx->d_match->key, 0xff, sizeof(unsigned long long), ctx->stream));
    if (L.row_end <= L.row_begin) return B200SLAM_OK;
    const int th_first = (int)(L.row_begin / L.ntx);
    const int th_last = (int)((L.row_end - 1) / L.ntx);
    A.th_first = th_first;
    const int nth_cover = th_last - th_first + 1;
    const long long cands = (long long)(L.row_end - L.row_begin) * L.nty;

    // Tile shape: big register tiles amortise the per-chunk tables on large sweeps;
    // small lattices need many short threads to fill 148 SMs.
    if (cands >= (1ll << 21) && L.nty >= 64 && L.ntx >= 64)
        return launch_lattice_cfg<16, 2, 4, 64>(ctx, A, nth_cover);      // 64 x 64 tile, 256 thr
    if (cands >= (1ll << 18) && L.nty >= 32)
        return launch_lattice_cfg<8, 1, 4, 64>(ctx, A, nth_cover);       // 32 x 32 tile, 128 thr
    if (L.nty >= 8)
        return launch_lattice_cfg<2, 1, 4, 128>(ctx, A, nth_cover);      // 32 x 8 tile, 128 thr
    return launch_lattice_cfg<1, 1, 4, 128>(ctx, A, nth_cover);          // 32 x 4 tile
}

int trace_launch(b200slam_ctx *ctx, const LatticeLaunch &L, bool use_gathered_keys)
{
    const b200slam_map *m = L.map;
    TraceArgs T;
    T.field = m->d_field; T.pitch = m->field_pitch; T.rows = m->rows; T.cols = m->cols;
    T.scan_x = ctx->d_scan_x; T.scan_y = ctx->d_scan_y; T.nbeams = ctx->nbeams;
    T.ipixel = 1 / m->pixel_size;
    T.ct = L.d_ct; T.st = L.d_st; T.sxt = L.d_sxt; T.syt = L.d_syt;
    T.nth = L.nth; T.ntx = L.ntx; T.nty = L.nty;
    T.last_lin = L.row_end > L.row_begin ? (long long)L.row_end * L.nty - 1 : 0;
    T.keys = use_gathered_keys ? ctx->d_keys : &ctx->d_match->key;
    T.nkeys = use_gathered_keys ? ctx->nranks : 1;
    T.out = ctx->d_match;
    T.hit_values = ctx->d_hit_values;
    T.hit_stride = ctx->scan_cap;
    trace_kernel<<<2, 32, 0, ctx->stream>>>(T);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
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
    A.best_key = &ctx->d_match->key;
    CUDA_TRY(ctx, cudaMemsetAsync(&ctx->d_match->key, 0xff, sizeof(unsigned long long), ctx->stream));
    if (P <= 0) return B200SLAM_OK;
    const unsigned grid = (unsigned)((P + POSES_THREADS - 1) / POSES_THREADS);
    poses_kernel<<<grid, POSES_THREADS, 0, ctx->stream>>>(A);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}
