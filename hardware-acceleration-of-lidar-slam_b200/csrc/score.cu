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
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "trig.cuh"

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
// (int)roundf(v) for every v whose result can pass the validity test below: for v >= 0 adding 0.5
// with round-toward-zero and truncating IS round-half-away-from-zero (exhaustively equal to glibc
// roundf); for v < 0, NaN and out-of-range values both this and roundf give something <= 0 or
// saturated, i.e. an invalid cell either way.
__device__ __forceinline__ int round_cell(float v)
{
    return __float2int_rz(__fadd_rz(v, 0.5f));
}
// 0-based cell index of (int)roundf(v) + 1 - 1, or -1 when the 1-based index fails
// `1 < S < n` (main.c:512).
__device__ __forceinline__ int cell_index(float v, int n)
{
    const int r = round_cell(v);
    return ((unsigned)(r - 1) < (unsigned)(n - 2)) ? r : -1;
}

// Read-only gather whose position in the instruction stream the compiler must keep
// (volatile asm statements are not reordered against each other).
__device__ __forceinline__ float ldg_ordered(const float *p)
{
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
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

// ---- multi-GPU: per-rank results travel through NVLink peer memory ----------------------
// POST (tail of a scoring kernel): this rank's {key, best_hits, last_hits} of exchange number
// `epoch` is stored into EVERY peer's buffer.  Fire and forget: nobody waits here, so ranks
// are not forced into lockstep.
__device__ __forceinline__ void exchange_post(const XchgArgs &X, unsigned int epoch, unsigned long long key,
                                              int best_hits, int last_hits, unsigned int *words_smem)
{
    xchg_post_words(X, epoch, (unsigned int)key, (unsigned int)(key >> 32), (unsigned int)best_hits, (unsigned int)last_hits,
                    words_smem);
}

// COLLECT: waits until every rank's words of the oldest uncollected exchange have landed in OUR
// buffer and publishes the GLOBAL result (lowest key wins; the last candidate of the whole
// lattice belongs to the highest rank that scored anything) in match->g*.  Runs either as its
// own one-CTA kernel or in the tail of the NEXT scoring kernel (a step after the posts, when they
// have long arrived), so ranks drift by up to a step instead of running in lockstep.  Four
// epochs of slots are enough: a rank posts exchange k only after collecting k-2, which needed
// every peer's post of k-2, which those peers issued after collecting k-4.
// Called by a whole CTA (>= 64 threads); got: shared scratch of 4 * XCHG_MAX_RANKS words.
__device__ __forceinline__ void exchange_collect(const XchgArgs &X, MatchDev *match, unsigned int epoch,
                                                 unsigned int *got)
{
    const int tid = threadIdx.x;
    // Bounded wait: a peer that died or was queued out of order costs one timeout and a sticky error bit
    // (B200SLAM_ERR_STATE from b200slam_match_fetch / b200slam_sync), not a GPU that spins forever.  Once
    // the bit is set later collects do not wait at all.
    for (int i = tid; i < 4 * X.nranks; i += blockDim.x) {
        const int r = i >> 2, w = i & 3;
        unsigned int v;
        if (!xchg_wait_word(X, epoch, r, w, &v, &match->error, DEV_ERR_EXCHANGE))
            v = w < 2 ? 0xffffffffu : 0u;                   // "nothing scored" for the missing rank
        got[i] = v;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long k = ~0ull;
        int bh = 0, lh = 0;
        for (int r = 0; r < X.nranks; ++r) {
            const unsigned long long kr = (unsigned long long)got[4 * r] | ((unsigned long long)got[4 * r + 1] << 32);
            if (kr == ~0ull) continue;
            if (kr < k) { k = kr; bh = (int)got[4 * r + 2]; }
            lh = (int)got[4 * r + 3];
        }
        match->gkey = k;
        match->gbest_hits = bh;
        match->glast_hits = lh;
        match->collected = epoch;
    }
}

__global__ void __launch_bounds__(64) exchange_collect_kernel(MatchDev *match, const XchgArgs X)
{
    __shared__ unsigned int got[4 * XCHG_MAX_RANKS];
    __shared__ unsigned int todo[3];
    __shared__ unsigned int words_s[4];
    if (threadIdx.x == 0) {
        todo[0] = *reinterpret_cast<volatile unsigned int *>(&match->collected);
        todo[1] = *reinterpret_cast<volatile unsigned int *>(&match->epoch);
        todo[2] = *reinterpret_cast<volatile unsigned int *>(&match->posted);
    }
    __syncthreads();
    // results recorded in the outbox by a burst of deferred-post matches go out first
    for (unsigned int e = todo[2] + 1; e <= todo[1]; ++e) {
        const MatchDev::Outbox *ob = &match->outbox[e % XCHG_EPOCHS];
        exchange_post(X, e, *reinterpret_cast<const volatile unsigned long long *>(&ob->key),
                      *reinterpret_cast<const volatile int *>(&ob->best_hits),
                      *reinterpret_cast<const volatile int *>(&ob->last_hits), words_s);
        __syncthreads();
    }
    if (threadIdx.x == 0) match->posted = todo[1];
    for (unsigned int e = todo[0] + 1; e <= todo[1]; ++e) {      // everything posted so far
        exchange_collect(X, match, e, got);
        __syncthreads();
    }
}

// Empty shard of a multi-GPU match: nothing to score, but the peers expect this rank's post -- through the
// same channel the scoring kernel's tail would have used: recorded in the outbox when the burst's posts are
// deferred (the collect kernel sends the outbox in epoch order, so nothing recorded earlier is skipped), else
// stored into the peers' buffers right away, merging the previous exchange here when the caller deferred
// that to the next tail (so a rank that only ever posts stays within the ring like everybody else).
__global__ void __launch_bounds__(64) exchange_only_kernel(MatchDev *match, const XchgArgs X, int post_deferred,
                                                           int collect_prev)
{
    __shared__ unsigned int words_s[4];
    __shared__ unsigned int got[4 * XCHG_MAX_RANKS];
    __shared__ unsigned int epoch_s, done_s;
    if (threadIdx.x == 0) {
        epoch_s = *reinterpret_cast<volatile unsigned int *>(&match->epoch) + 1;
        done_s = *reinterpret_cast<volatile unsigned int *>(&match->collected);
    }
    __syncthreads();
    if (post_deferred) {
        if (threadIdx.x == 0) {
            MatchDev::Outbox *ob = &match->outbox[epoch_s % XCHG_EPOCHS];
            ob->key = ~0ull; ob->best_hits = 0; ob->last_hits = 0;
        }
    } else {
        exchange_post(X, epoch_s, ~0ull, 0, 0, words_s);
        if (threadIdx.x == 0) match->posted = epoch_s;
        if (collect_prev && done_s + 1 < epoch_s) exchange_collect(X, match, done_s + 1, got);
    }
    if (threadIdx.x == 0) {
        match->key = ~0ull;
        match->best_hits = 0;
        match->last_hits = 0;
        match->epoch = epoch_s;
    }
}

// Replays two candidates at once with the whole CTA (the winner and the last one scored):
// in-bounds field values are compacted in beam order into vals0 / vals1 (main.c:515);
// the counts come back in every thread.  red: shared scratch of 2 * NT/32 ints.
struct TraceCand {
    float ct, st, sxt, syt;
};
// lo0 / lo1: only compacted positions >= lo are stored (the staircase behind the last candidate).
template <int NT>
__device__ void trace_pair(const float *__restrict__ field, int pitch, int rows, int cols,
                           const float *__restrict__ scan_x, const float *__restrict__ scan_y, int nbeams,
                           float ipixel, const TraceCand (&cand)[2], float *vals0, float *vals1, int *red,
                           int (&count)[2], int lo0 = 0, int lo1 = 0)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    count[0] = count[1] = 0;
    for (int i0 = 0; i0 < nbeams; i0 += NT) {
        const int i = i0 + tid;
        bool in[2] = {false, false};
        float v[2] = {0.0f, 0.0f};
        if (i < nbeams) {
            const float psx = __fmul_rn(scan_x[i], ipixel);
            const float psy = __fmul_rn(scan_y[i], ipixel);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int c = cell_index(__fadd_rn(rot_x(psx, psy, cand[k].ct, cand[k].st), cand[k].sxt), cols);
                const int r = cell_index(__fadd_rn(rot_y(psx, psy, cand[k].ct, cand[k].st), cand[k].syt), rows);
                in[k] = c >= 0 && r >= 0;
                v[k] = field[in[k] ? (long)r * pitch + c : -1];
            }
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, in[0]);
        const unsigned m1 = __ballot_sync(0xffffffffu, in[1]);
        __syncthreads();                       // red[] free again
        if (lane == 0) { red[warp] = __popc(m0); red[NT / 32 + warp] = __popc(m1); }
        __syncthreads();
        int before0 = 0, total0 = 0, before1 = 0, total1 = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
            const int c0 = red[w], c1 = red[NT / 32 + w];
            before0 += w < warp ? c0 : 0; total0 += c0;
            before1 += w < warp ? c1 : 0; total1 += c1;
        }
        const unsigned below = (1u << lane) - 1u;
        const int p0 = count[0] + before0 + __popc(m0 & below), p1 = count[1] + before1 + __popc(m1 & below);
        if (in[0] && p0 >= lo0) vals0[p0] = v[0];
        if (in[1] && p1 >= lo1) vals1[p1] = v[1];
        count[0] += total0;
        count[1] += total1;
    }
}

constexpr int LATTICE_TABLE_FLOATS = 960;      // ct | st | sxt | syt as kernel parameters
struct LatticeTables {
    float v[LATTICE_TABLE_FLOATS];
};

// Gather pipeline shape: U beams per group, NBUF register buffers (NBUF - 1 groups in flight);
// GUARD = table rows appended for padding (a multiple of NBUF*U) and for the prefetches that
// run past the end.
template <int TYPT, int UM = 1, int Q = 0, int NB = 0>
struct LatticePipe {
    static constexpr int U = UM * (TYPT >= 16 ? 1 : 16 / TYPT);      // UM: deeper groups for low-occupancy shapes
    // 16 candidates per thread: two buffers.  With row reuse that is 16 sums + 2 x 9 gathered values in 64
    // registers = 4 CTAs (32 warps) per SM; the kernel is latency bound, and measured on config 3 occupancy beats
    // a third buffer: 648 us against 684 us with three buffers at 3 CTAs per SM (80 registers), 1386 us with three
    // buffers squeezed into 64 registers (spills), 1011 us at 2 CTAs per SM.
    static constexpr int NBUF = NB > 0 ? NB : (TYPT >= 16 ? 2 : 3);
    static constexpr int PAD = NBUF * U;
    static constexpr int GUARD = PAD - 1 + (NBUF - 1) * U;
};

struct LatticeArgs {
    const float *field;       // [0][0]; field[-1] == 0, and so are the B200SLAM_FIELD_PAD_ROWS rows of floats in front of it
    int pitch, rows, cols;
    int rr_clamp;             // row reuse: index of an invalid beam, -(8 * pitch + 1): its K <= 9 row reads all land in the zero region
    const float *scan_x, *scan_y;
    int nbeams;
    const int *nbeams_dev;    // not null: the scan's size as the device knows it (asynchronous scan loop; nbeams is an upper bound)
    MatchHost *host_result;   // not null: the tail also writes the result block into mapped host memory, seq last
    unsigned long long host_seq;
    const int *mp_n_dev;      // the map points' size on the device (rides along in the host result block)
    int seeded;               // 3 x 3 x 3 lattice centred on the winner of the match in front (b200slam_fastmatch_pair_async):
                              // the parameter block holds ct[3][3] | st[3][3] | sxt[3][3] | syt[3][3], one row per possible seed
    float ipixel;
    const float *tables;      // device copy of ct|st|sxt|syt, or nullptr: use the parameter block
    int nth, ntx, nty;
    int th_first;             // first theta index covered by blockIdx.z == first entry of ct / st
    int nth_tab;              // number of theta entries in the tables
    long long row_begin, row_end;
    MatchDev *match;
    float *scores;            // optional
    float *hit_values;        // [2][hit_stride]: winner / last candidate
    int hit_stride;
    int cb;                   // beams per shared-memory chunk
    unsigned total_ctas;
    int collect_prev;         // exchange: also merge the previous, still uncollected exchange in this tail
    int post_deferred;        // exchange: record the result in the outbox, the next collect kernel posts it
    XchgArgs xchg;            // peers == nullptr: single GPU / no exchange
};

// TYPT candidates (consecutive ty) per thread, WX warps along tx, WY warps along ty.
// Dynamic shared memory: colT[cb][TXT] | rowT[cb][TYT] | Sx[cb] | Sy[cb]
// COUNT (FastMatch-sized lattices only): every thread also counts its candidate's in-bounds beams.
// Q > 0: ROW REUSE.  When the ty step is pixel / Q, the TYPT consecutive ty candidates of a thread fall
// into only K = TYPT / Q + 1 consecutive rows of the field: row(ty_j) = r0 + floor((j + p) / Q) for one
// "phase" p in [0, Q) that depends on the beam (and the warp's ty group) alone.  The row table then holds
// {r0 * pitch, p} per (beam, ty group) instead of TYPT offsets, a thread gathers its column from the K rows
// once -- K loads instead of TYPT, same 1-2 lines each -- and the adds pick the value by compile-time index
// inside a warp-uniform switch on p.  Everything is decided from the exact per-candidate row indices: a
// (beam, ty group) whose rows do not follow the pattern (float fuzz at a rounding boundary, rows at the edge
// of the grid) is flagged and takes the per-candidate path for that beam, so results stay bit-identical.
// PITCH (row reuse only): the field's row pitch in floats when it is known at compile time (the power-of-two map
// sizes of the BASELINE configurations), 0 = read it from the arguments.  The K row reads of a beam are then ONE
// address computation and K loads at immediate offsets k * PITCH * 4 instead of a chain of K IMAD.WIDEs.
template <int TYPT, int WX, int WY, int UM = 1, bool COUNT = false, int Q = 0, int PITCH = 0, int NB = 0>
__global__ void __launch_bounds__(32 * WX * WY, TYPT >= 16 ? (Q > 0 ? 4 : 3) : 1)
lattice_kernel(const __grid_constant__ LatticeArgs A, const __grid_constant__ LatticeTables T)
{
    static_assert(!COUNT || TYPT == 1, "per-candidate hit counts: one candidate per thread");
    static_assert(Q == 0 || (TYPT % Q == 0 && !COUNT), "row reuse: TYPT a multiple of Q");
    constexpr int K = Q > 0 ? TYPT / Q + 1 : TYPT;       // values a thread gathers per beam
    constexpr int RW = Q > 0 ? WY : TYPT * WY;           // row-table ints per beam (Q > 0: r0 * pitch | phase per ty group)
    constexpr int TXT = 32 * WX;
    constexpr int TYT = TYPT * WY;
    constexpr int NT = 32 * WX * WY;
    using P = LatticePipe<TYPT, UM, Q, NB>;
    constexpr int U = P::U, NBUF = P::NBUF, GUARD = P::GUARD;
    static_assert(NT % TXT == 0 && NT % TYT == 0, "tile shape");
    extern __shared__ __align__(16) int lat_smem[];
    int *colT = lat_smem;
    int *rowT = colT + (A.cb + GUARD) * TXT;             // chunk rows + padding/guard rows
    float *Sx_s = reinterpret_cast<float *>(rowT + (A.cb + GUARD) * RW);
    float *Sy_s = Sx_s + A.cb;
    float *syt_s = Sy_s + A.cb;                          // Q > 0: this tile's TYT ty offsets (per-candidate path)
    __shared__ unsigned long long red[WX * WY];
    __shared__ int tail_red[2 * (NT / 32)];
    __shared__ int last_flag;
    __shared__ unsigned int xchg_words[4];
    __shared__ unsigned int epoch_s;

    pdl_launch_dependents();                      // the next scan-matching kernel may start
    const float *tab = A.tables ? A.tables : T.v;
    const float *ctT = tab, *stT = tab + A.nth_tab, *sxtT = tab + 2 * A.nth_tab, *sytT = sxtT + A.ntx;
    const int nbeams = A.nbeams_dev ? *A.nbeams_dev : A.nbeams;
    if constexpr (COUNT) {
        if (A.seeded) {
            // The lattice is centred on the winner of the match in front (main.c:909-918: FastMatch2 starts from
            // FastMatch's result).  Each axis of that winner is one of three values, so the host has sent the axis
            // tables of all three; wait for the match in front, read its winner and pick.  Its key stays in
            // match->seed_key (this kernel's tail overwrites match->key).
            pdl_wait_prior_grids();
            const unsigned long long k1 = *reinterpret_cast<volatile unsigned long long *>(&A.match->key);
            const int lin1 = k1 == ~0ull ? 13 : (int)(k1 & 0xffffffffull);          // nothing scored cannot happen for a full lattice
            const int ith1 = lin1 / 9, itx1 = (lin1 / 3) % 3, ity1 = lin1 % 3;
            ctT = tab + 3 * ith1; stT = tab + 9 + 3 * ith1; sxtT = tab + 18 + 3 * itx1; sytT = tab + 27 + 3 * ity1;
            if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) A.match->seed_key = k1;
        }
    }

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wx = warp % WX, wy = warp / WX;
    const int ith = A.th_first + blockIdx.z;
    const int tx0 = blockIdx.y * TXT, ty0 = blockIdx.x * TYT;

    unsigned long long best = ~0ull;
    [[maybe_unused]] int nhits = 0;                              // COUNT only: this thread's candidate's in-bounds beams
    [[maybe_unused]] long long count_lin = -1;                   // COUNT only: that candidate's linear index
    // whole tile outside this shard's (theta, tx) row range?  (still takes a ticket)
    const long long r_lo = (long long)ith * A.ntx + tx0;
    const long long r_hi = r_lo + min(TXT, A.ntx - tx0);
    if (r_hi > A.row_begin && r_lo < A.row_end) {
        const float ct = ctT[blockIdx.z], st = stT[blockIdx.z];
        const int txl = wx * 32 + lane;          // tile-local tx of this thread
        const int tyl = wy * TYPT;               // first tile-local ty of this thread

        float acc[TYPT];
#pragma unroll
        for (int j = 0; j < TYPT; ++j) acc[j] = 0.0f;            // main.c:507

        for (int c0 = 0; c0 < nbeams; c0 += A.cb) {
            const int cb = min(A.cb, nbeams - c0);
            __syncthreads();
            for (int i = tid; i < cb; i += NT) {
                const float psx = __fmul_rn(A.scan_x[c0 + i], A.ipixel);     // main.c:418
                const float psy = __fmul_rn(A.scan_y[c0 + i], A.ipixel);     // main.c:419
                Sx_s[i] = rot_x(psx, psy, ct, st);
                Sy_s[i] = rot_y(psx, psy, ct, st);
            }
            __syncthreads();
            const int cbp = (cb + P::PAD - 1) / P::PAD * P::PAD;             // padded with INVALID rows
            const int cbg = cbp + (NBUF - 1) * U;                            // + guard rows for the prefetches
            {
                // column table: thread owns tile column t (its axis value is loop invariant)
                // and walks the beams NT/TXT at a time
                const int t = tid % TXT;
                const bool tok = tx0 + t < A.ntx;
                const float sx = tok ? sxtT[tx0 + t] : 0.0f;
#pragma unroll 4
                for (int i = tid / TXT; i < cbg; i += NT / TXT) {
                    int v = INVALID_OFF;
                    if (tok && i < cb) {
                        const int c = cell_index(__fadd_rn(Sx_s[i], sx), A.cols);       // main.c:483
                        if (c >= 0) v = c;
                    }
                    colT[i * TXT + t] = v;
                }
            }
            if constexpr (Q == 0) {
                const int t = tid % TYT;
                const bool tok = ty0 + t < A.nty;
                const float sy = tok ? sytT[ty0 + t] : 0.0f;
#pragma unroll 4
                for (int i = tid / TYT; i < cbg; i += NT / TYT) {
                    int v = INVALID_OFF;
                    if (tok && i < cb) {
                        const int r = cell_index(__fadd_rn(Sy_s[i], sy), A.rows);       // main.c:501
                        if (r >= 0) v = r * A.pitch;
                    }
                    rowT[i * TYT + t] = v;
                }
            } else {
                if (c0 == 0) {
                    for (int t = tid; t < TYT; t += NT) syt_s[t] = ty0 + t < A.nty ? sytT[ty0 + t] : 0.0f;
                    __syncthreads();
                }
                // one entry per (beam, ty group): exact rows of the group's TYPT candidates -> {r0 * pitch, phase},
                // phase == Q: the rows do not follow r0 + floor((j + p) / Q) (or leave the grid): per-candidate path
                for (int e = tid; e < cbg * WY; e += NT) {
                    const int i = e / WY, g = e % WY;
                    int off0 = INVALID_OFF, phase = 0;
                    if (i < cb) {
                        const float sy = Sy_s[i];
                        int r[TYPT];
#pragma unroll
                        for (int j = 0; j < TYPT; ++j) r[j] = cell_index(__fadd_rn(sy, syt_s[g * TYPT + j]), A.rows);   // main.c:501
                        const int nj = min(TYPT, A.nty - (ty0 + g * TYPT));      // candidates of this group inside the lattice
                        phase = Q;
                        if (nj > 0 && r[0] >= 0 && r[0] + K - 1 < A.rows) {
#pragma unroll
                            for (int p = 0; p < Q; ++p) {
                                bool ok = true;
#pragma unroll
                                for (int j = 0; j < TYPT; ++j) ok = ok && (j >= nj || r[j] == r[0] + (j + p) / Q);
                                if (ok && phase == Q) phase = p;
                            }
                            if (phase < Q) off0 = r[0] * A.pitch;
                        }
                        if (nj <= 0) phase = 0;              // nothing to score here: adds +0.0f to unused sums
                    }
                    rowT[e] = off0 | phase;                  // the pitch is a multiple of 32 and INVALID_OFF of 2^30: low bits free
                }
            }
            __syncthreads();
            // Software-pipelined gather: warps issue in order, so the loads of the next group
            // of U beams are put in flight before the (sequential, in-order) additions of
            // the current group.  Table rows [cb, cbp) are INVALID and add +0.0f.
            const int *cp = colT + txl;
            const int *rp = rowT + (Q > 0 ? wy : tyl);
            [[maybe_unused]] int phase_q[NBUF][U];           // Q > 0: phase of every beam in flight
            // cq / rq: this thread's column-table / row-table entries of the group's first beam.  An invalid column
            // or row makes the sum hugely negative -> clamped to rr_clamp, from where all K row reads hit the zero
            // rows in front of the field: no select on the row stride, no branch.
            [[maybe_unused]] auto gather_rr = [&](const int *cq, const int *rq, float (&dst)[U][K], int (&ph)[U]) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = cq[u * TXT];
                    const int rv = rq[u * RW];                                           // r0 * pitch | phase, warp uniform
                    const int idx = __viaddmax_s32(c, rv & ~7, A.rr_clamp);
                    ph[u] = rv & 7;
                    const float *p0 = A.field + idx;
#pragma unroll
                    for (int k = 0; k < K; ++k) dst[u][k] = ldg_ordered(PITCH > 0 ? p0 + k * PITCH : p0 + (long)k * A.pitch);
                }
#pragma unroll
                for (int j = 0; j < TYPT; ++j) asm volatile("" : "+f"(acc[j]));
            };
            // adds of one beam, in scan order (main.c:516): candidate j takes row (j + phase) / Q of the K gathered
            // ones.  The phase is warp uniform, so each phase gets its own copy of the adds with compile-time row
            // indices -- no selects: measured on config 3, 684 us against 692 us with FSELs and 784 us with packed
            // FADD2 adds (whose even-aligned register pairs made the allocator spill).
            [[maybe_unused]] auto accumulate_rr = [&](const float (&src)[U][K], const int (&ph)[U], int i0) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    constexpr int QQ = Q > 0 ? Q : 1;
                    if (ph[u] < QQ) {                        // warp uniform; almost always
                        auto add_phase = [&](auto pc) {
                            constexpr int p = decltype(pc)::value;
#pragma unroll
                            for (int j = 0; j < TYPT; ++j) acc[j] = __fadd_rn(acc[j], src[u][(j + p) / QQ]);
                        };
                        switch (ph[u]) {
                            case 0: add_phase(std::integral_constant<int, 0>()); break;
                            case 1: add_phase(std::integral_constant<int, 1 % QQ>()); break;
                            case 2: add_phase(std::integral_constant<int, 2 % QQ>()); break;
                            default: add_phase(std::integral_constant<int, 3 % QQ>()); break;
                        }
                    } else {
                        // per-candidate path for this beam (rows off the pattern or at the grid's edge)
                        const int c = cp[(i0 + u) * TXT];
                        const float sy = Sy_s[i0 + u];
#pragma unroll
                        for (int j = 0; j < TYPT; ++j) {
                            const int r = cell_index(__fadd_rn(sy, syt_s[tyl + j]), A.rows);           // main.c:501
                            const int off = (c >= 0 && r >= 0 && ty0 + tyl + j < A.nty) ? r * A.pitch + c : -1;
                            acc[j] = __fadd_rn(acc[j], __ldg(A.field + off));
                        }
                    }
                }
            };
            auto gather = [&](int i0, float (&dst)[U][TYPT]) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u;
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
#pragma unroll
                    for (int j = 0; j < TYPT; ++j) {
                        const int off = __viaddmax_s32(c, ro[j], -1);
                        dst[u][j] = ldg_ordered(A.field + off);
                        if constexpr (COUNT) nhits += off >= 0 ? 1 : 0;
                    }
                }
                // order fence for the compiler: the additions below stay behind these loads
#pragma unroll
                for (int j = 0; j < TYPT; ++j) asm volatile("" : "+f"(acc[j]));
            };
            auto accumulate = [&](const float (&src)[U][TYPT]) {
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < TYPT; ++j) acc[j] = __fadd_rn(acc[j], src[u][j]);      // main.c:516
            };
            if constexpr (Q > 0) {
                float v[NBUF][U][K];
                const int *cq = cp, *rq = rp;                // table entries of beam i0: every load below is [register + immediate]
#pragma unroll
                for (int b = 0; b < NBUF - 1; ++b) gather_rr(cq + b * U * TXT, rq + b * U * RW, v[b], phase_q[b]);
                for (int i0 = 0; i0 < cbp; i0 += NBUF * U, cq += NBUF * U * TXT, rq += NBUF * U * RW) {
#pragma unroll
                    for (int b = 0; b < NBUF; ++b) {
                        gather_rr(cq + (b + NBUF - 1) * U * TXT, rq + (b + NBUF - 1) * U * RW, v[(b + NBUF - 1) % NBUF],
                                  phase_q[(b + NBUF - 1) % NBUF]);
                        accumulate_rr(v[b], phase_q[b], i0 + b * U);
                    }
                }
            } else {
                // NBUF rotating register buffers: group g+NBUF-1 is requested before group g is added
                float v[NBUF][U][TYPT];
#pragma unroll
                for (int b = 0; b < NBUF - 1; ++b) gather(b * U, v[b]);
                for (int i0 = 0; i0 < cbp; i0 += NBUF * U) {
#pragma unroll
                    for (int b = 0; b < NBUF; ++b) {
                        gather(i0 + (b + NBUF - 1) * U, v[(b + NBUF - 1) % NBUF]);   // past the end: guard rows
                        accumulate(v[b]);
                    }
                }
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
                    if constexpr (COUNT) count_lin = lin;
                    const unsigned long long k = pack_key(acc[j], (unsigned int)lin);
                    best = k < best ? k : best;
                }
            }
        }
    }
    // Scores and tables above depend on the scan, the launch parameters and the field only.  The
    // match state below is shared with the kernel in front (its last CTA resets it).
    pdl_wait_prior_grids();
    // cand_hits[] belongs to the match state: the kernel in front may still be walking it in its
    // bestHits staircase until it has completed, so it is written only behind the wait
    if constexpr (COUNT)
        if (count_lin >= 0) A.match->cand_hits[count_lin] = nhits;
    best = warp_min_u64(best);
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (tid == 0)
        for (int w = 1; w < WX * WY; ++w) best = red[w] < best ? red[w] : best;

    // ---- last CTA: publish the winner, trace winner and last candidate, reset -----------
    if (!cta_is_last(A.match, best, A.total_ctas, &last_flag)) return;
    __threadfence();
    const unsigned long long key = *reinterpret_cast<volatile unsigned long long *>(&A.match->work_key);
    if (tid == 0) epoch_s = *reinterpret_cast<volatile unsigned int *>(&A.match->epoch) + 1;
    int hits[2] = {0, 0};
    if (key != ~0ull) {
        TraceCand cand[2];
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const long long lin = which == 0 ? (long long)(key & 0xffffffffull) : A.row_end * A.nty - 1;
            const int ity = (int)(lin % A.nty);
            const long long row = lin / A.nty;
            const int itx = (int)(row % A.ntx), jth = (int)(row / A.ntx);
            cand[which].ct = ctT[jth - A.th_first]; cand[which].st = stT[jth - A.th_first];
            cand[which].sxt = sxtT[itx]; cand[which].syt = sytT[ity];
        }
        trace_pair<NT>(A.field, A.pitch, A.rows, A.cols, A.scan_x, A.scan_y, nbeams, A.ipixel, cand,
                       A.hit_values, A.hit_values + A.hit_stride, tail_red, hits);
    }
    int written = hits[1];
    if constexpr (COUNT) {
        // main.c:515: every candidate overwrites bestHits[] from index 0 in loop order, so behind the
        // last candidate's hits[1] values the array holds those of the most recent candidate that had
        // more hits, and so on: walk the candidates backwards and let each one that is longer than what
        // has been written so far supply the entries it alone still owns.
        if (key != ~0ull) {
            const int ncand = A.nth * A.ntx * A.nty;
            for (int c = ncand - 2; c >= 0; --c) {
                const int nc = *reinterpret_cast<volatile int *>(&A.match->cand_hits[c]);
                if (nc <= written) continue;                                       // uniform across the CTA
                const int ity = c % A.nty, itx = (c / A.nty) % A.ntx, jth = c / A.nty / A.ntx;
                TraceCand cand[2];
                cand[0].ct = ctT[jth - A.th_first]; cand[0].st = stT[jth - A.th_first];
                cand[0].sxt = sxtT[itx]; cand[0].syt = sytT[ity];
                cand[1] = cand[0];
                int n2[2];
                float *tail = A.hit_values + A.hit_stride;
                __syncthreads();
                trace_pair<NT>(A.field, A.pitch, A.rows, A.cols, A.scan_x, A.scan_y, nbeams, A.ipixel, cand, tail,
                               tail, tail_red, n2, written, written);
                written = nc;
            }
        }
    }
    const int best_hits = hits[0], last_hits = hits[1];
    const unsigned long long out_key = key;
    if (A.xchg.peers && A.post_deferred) {
        if (tid == 0) {
            MatchDev::Outbox *ob = &A.match->outbox[epoch_s % XCHG_EPOCHS];
            ob->key = key; ob->best_hits = best_hits; ob->last_hits = last_hits;
        }
    } else if (A.xchg.peers) {
        __syncthreads();
        exchange_post(A.xchg, epoch_s, key, best_hits, last_hits, xchg_words);
        if (tid == 0) A.match->posted = epoch_s;
        // merge the PREVIOUS exchange here (posted a whole step ago) when the caller deferred it
        if (A.collect_prev) {
            __shared__ unsigned int got[4 * XCHG_MAX_RANKS];
            __shared__ unsigned int done_s;
            if (tid == 0) done_s = *reinterpret_cast<volatile unsigned int *>(&A.match->collected);
            __syncthreads();
            if (done_s + 1 < epoch_s) exchange_collect(A.xchg, A.match, done_s + 1, got);
        }
    }
    if (tid == 0) {
        A.match->key = out_key;
        A.match->best_hits = best_hits;
        A.match->last_hits = last_hits;
        A.match->written_hits = written;
        A.match->work_key = ~0ull;
        A.match->tickets = 0u;
        if (A.xchg.peers) A.match->epoch = epoch_s;
        if (A.host_result) {
            volatile MatchHost *h = A.host_result;
            h->key = out_key;
            h->seed_key = *reinterpret_cast<volatile unsigned long long *>(&A.match->seed_key);
            h->best_hits = best_hits; h->last_hits = last_hits; h->written_hits = written;
            h->scan_n = nbeams;
            h->mp_n = A.mp_n_dev ? *A.mp_n_dev : 0;
            h->error = *reinterpret_cast<volatile unsigned int *>(&A.match->error);
            __threadfence_system();
            h->seq = A.host_seq;
        }
    }
}

// ---- arbitrary pose list (particles): one thread per pose, beams sequential ---------
struct PosesArgs {
    const float *field11;     // &field[1][1]
    int zero_off;             // -(pitch + 2)
    const float *field;
    int pitch, rows, cols;
    const float *scan_x, *scan_y;
    int nbeams;
    const int *nbeams_dev;    // not null: the device's own scan size
    float ipixel, min_x, min_y;
    const float *px, *py, *ct, *st;    // [P]
    long long P, index_base;
    float *scores;                     // [P]
    int *hits;                         // [P]
    MatchDev *match;
    unsigned total_ctas;
    XchgArgs xchg;                     // peers != nullptr: sharded particle filter, the tail posts {key, P} to every rank
};

constexpr int POSES_THREADS = 128;
constexpr int POSES_CB = 1024;

__global__ void __launch_bounds__(POSES_THREADS) poses_kernel(const __grid_constant__ PosesArgs A)
{
    __shared__ float2 ps[POSES_CB + 64];
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
    // S_x = x ct + y st + sxt ; S_y = x (-st) + y ct + syt  (main.c:462-463, 483, 501), as (S_x, S_y) pairs
    const f32x2_t rot_a = f2_pack(ct, -st), rot_b = f2_pack(st, ct), shift = f2_pack(sxt, syt), half2 = f2_pack(0.5f, 0.5f);
    float score = 0.0f;
    int nh = 0;
    const unsigned cols_m2 = (unsigned)(A.cols - 2), rows_m2 = (unsigned)(A.rows - 2);
    const f32x2_t magic2 = f2_pack(8388608.0f, 8388608.0f);
    const unsigned upitch = (unsigned)A.pitch;
    const float *field11 = A.field11;                    // &field[1][1] (from the host: one uniform base pointer)
    const int zero_off = A.zero_off;                     // field11 + zero_off == field[-1] == 0
    const int nbeams = A.nbeams_dev ? *A.nbeams_dev : A.nbeams;
    for (int c0 = 0; c0 < nbeams; c0 += POSES_CB) {
        const int cb = min(POSES_CB, nbeams - c0);
        constexpr int U = 8, NBUF = 3;
        const int cbp = (cb + NBUF * U - 1) / (NBUF * U) * (NBUF * U);
        __syncthreads();
        // beams [cb, cbp + (NBUF-1)*U) are padding: points that land outside every grid and add +0.0f
        for (int i = threadIdx.x; i < cbp + (NBUF - 1) * U; i += POSES_THREADS)
            ps[i] = i < cb ? make_float2(__fmul_rn(A.scan_x[c0 + i], A.ipixel), __fmul_rn(A.scan_y[c0 + i], A.ipixel))
                           : make_float2(-1.0e30f, -1.0e30f);
        __syncthreads();
        if (live) {
            // software-pipelined like the lattice kernel: the gathers of the next beams are in
            // flight while the current ones are added in beam order (main.c:516)
            auto gather = [&](int i0, float (&dst)[U]) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    // (x, y) halves of one packed op each: FMUL2 / FADD2 / FADD2.RZ, every half rounded
                    // like the scalar instruction it replaces (10 scalar ops -> 6)
                    const float2 q = ps[i0 + u];
                    // The two products are summed with SCALAR adds: ptxas contracts mul.rn.f32x2 + add.rn.f32x2
                    // into FFMA2 (one rounding) whatever --fmad says, which would break bit-exactness.
                    float ax, ay, bx, by;
                    f2_unpack(f2_mul_rn(rot_a, f2_pack(q.x, q.x)), ax, ay);         // x ct | x (-st)
                    f2_unpack(f2_mul_rn(rot_b, f2_pack(q.y, q.y)), bx, by);         // y st | y ct
                    const f32x2_t f = f2_add_rn(f2_pack(__fadd_rn(ax, bx), __fadd_rn(ay, by)), shift);
                    // (int)roundf(v) (main.c:483, 501) = truncate(v + 0.5 toward zero) for every v that can pass
                    // the range test; the truncation is a second toward-zero add of 2^23, which leaves the
                    // integer in the mantissa (no F2I: that pipe runs at a quarter rate).  Negative, huge and
                    // NaN inputs give bit patterns that fail the unsigned test below (dimensions < 2^22).
                    float hx, hy;
                    f2_unpack(f2_add_rz(f2_add_rz(f, half2), magic2), hx, hy);
                    const unsigned cm1 = __float_as_uint(hx) - 0x4B000001u;         // column index - 1 (1-based S - 2)
                    const unsigned rm1 = __float_as_uint(hy) - 0x4B000001u;
                    const bool in = (cm1 < cols_m2) & (rm1 < rows_m2);              // 1 < S < n, main.c:512
                    const int off = in ? (int)(rm1 * upitch + cm1) : zero_off;      // relative to field[1][1]
                    dst[u] = __ldg(field11 + off);
                    nh += in ? 1 : 0;
                }
            };
            auto accumulate = [&](const float (&src)[U]) {
#pragma unroll
                for (int u = 0; u < U; ++u) score = __fadd_rn(score, src[u]);
            };
            float v[NBUF][U];
#pragma unroll
            for (int b = 0; b < NBUF - 1; ++b) gather(b * U, v[b]);
            for (int i0 = 0; i0 < cbp; i0 += NBUF * U) {
#pragma unroll
                for (int b = 0; b < NBUF; ++b) {
                    gather(i0 + (b + NBUF - 1) * U, v[(b + NBUF - 1) % NBUF]);
                    accumulate(v[b]);
                }
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
    __shared__ unsigned long long key_s;
    __shared__ unsigned int epoch_s, words_s[4];
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
        key_s = key;
        epoch_s = *reinterpret_cast<volatile unsigned int *>(&A.match->epoch) + 1;
    }
    if (A.xchg.peers) {
        // sharded particle filter, first exchange of the step: this rank's best (score, global index) and its
        // particle count go to every rank (the weights need the GLOBAL minimum score)
        __syncthreads();
        xchg_post_words(A.xchg, epoch_s, (unsigned int)key_s, (unsigned int)(key_s >> 32), (unsigned int)A.P, 0u, words_s);
        if (threadIdx.x == 0) {
            A.match->epoch = epoch_s;
            A.match->posted = epoch_s;
        }
    }
}


// ---- FastMatch-sized lattices: one thread-block cluster, one CTA per theta, sums in reference order -----------
// The reference's own call -- 27 candidates x <= 1079 beams (main.c:381-596) -- is 29 000 evaluations: a latency
// problem, not a throughput one.  The lattice kernel above gives every candidate one thread that gathers AND adds
// (27 live threads, ~45 us).  Here the work is split three ways (hard part 2: the north_star's
// "beams across lanes", with the exactness of design A kept):
//   CTAs     a cluster of min(n_theta, 8) CTAs; CTA r owns the candidates of theta r, r + C, ...  A divergent gather
//            costs the SM's L1 about two cycles per 128-byte line touched, ~8 lines per warp-load for neighbouring
//            beams: 27 candidates on one SM are bound by that (measured ~10 000 cycles), three SMs take a third each;
//   gather   the warps not on warp 0's scheduler; work item = (32-beam chunk, theta), chunks in scan order: rotate
//            once, n_tx column and n_ty row indices, the n_tx * n_ty loads -> vals[candidate][beam] in shared
//            memory (+0.0f for an out-of-bounds or padding beam, which leaves a non-negative sum unchanged), one
//            ballot per candidate -> in-bounds bit mask of the chunk, then the chunk's progress counter.
//            For the reference's 3 x 3 translations the loops are compile-time and the 9 loads in flight together;
//   sums     warp 0, one thread per candidate of the CTA, follows the gather chunk by chunk (a progress counter per chunk) and
//            adds vals[c][0 .. nbeams) SEQUENTIALLY in beam order (main.c:516) -- the same additions in the same
//            order as the reference, so the score is bit-identical.  One dependent FADD per beam (4 cycles) is the
//            floor of the kernel; warp 0 has its scheduler to itself so that nothing else delays the chain;
//   hits     the gather warps, after their last chunk: hits in front of every chunk (warp scan of the mask
//            popcounts) and per candidate;
//   exchange every CTA stores its candidates' hit counts and its best (score, index) key into the shared memory of
//            all CTAs (DSMEM), one cluster barrier; then each CTA knows the winner (arg-min: lowest score, then
//            lowest index) and all counts, and writes its share of bestHits[] exactly as the reference's loop leaves
//            it (main.c:515: every candidate overwrites the array from index 0, so behind the last candidate's hits
//            it holds those of the most recent candidate that had more -- a "staircase" of suffix maxima over the
//            counts): every (stair, chunk) pair is an independent compaction out of shared memory.
// No atomics, no second pass over the field, no global-memory protocol between the CTAs.
constexpr int FM_THREADS = 1024;
constexpr int FM_WARPS = FM_THREADS / 32;
constexpr int FM_GATHER_WARPS = FM_WARPS - FM_WARPS / 4;   // warps 4, 8, ... share warp 0's scheduler and stay idle
constexpr int FM_MAX_CAND = 32;
constexpr int FM_MAX_CLUSTER = 8;              // portable cluster size
constexpr int FM_MAX_BEAMS = 1536;
constexpr int FM_MAX_CHUNKS = FM_MAX_BEAMS / 32;
constexpr int FM_GROUP = 4;                    // chunks the summing warp takes at a time; scans are padded to whole groups
constexpr int FM_TAB_B = 128;                  // offset of the second pass's tables in the parameter block

struct FmMap {
    const float *field;       // [0][0]; field[-1] == 0
    int pitch, rows, cols;
    float ipixel;
    float tlx, tly;           // top-left corner (chain mode: the axis tables are computed on the device)
};

// One launch = optional readAScan + one or two matches ("passes").  Pass 0 reads its axis tables ct | st | sxt | syt
// packed at the front of the parameter block -- or, seeded0, all three candidate table sets per axis (36 floats),
// picked by the winner the kernel IN FRONT left in match->key.  Pass 1 (b200slam_scan_step_async: FastMatch2 from
// FastMatch's result, main.c:909-918) is always seeded, by pass 0's own winner, from the 36 floats at FM_TAB_B.
struct FmArgs {
    FmMap map[2];
    int npass;
    int seeded0;
    const float *scan_x, *scan_y;
    int nbeams;               // upper bound when nbeams_dev / read_scan is set
    const int *nbeams_dev;
    int nth, ntx, nty;
    int nbp;                  // row pitch of vals (beam capacity rounded up to a multiple of 32, plus 4)
    // chain mode (b200slam_scan_chain_step_async): the lattice centre comes from the device's own pose state, the axis
    // tables (main.c:424-437, cosf / sinf as glibc computes them: trig.cuh) are built here, the tail commits the pose
    ChainDev *chain;
    ChainSlot *chain_ring;
    int scan_index;
    int chain_count;          // scans this launch runs one after the other (their ranges lidar_n floats apart), while the chain lasts
    float step_a[3], step_b[3];
    float mini_dt, mini_dr;
    MatchDev *match;
    float *hit_values;        // the bestHits[] twin ([1] of the context's buffer)
    MatchHost *host_result;
    unsigned long long host_seq;
    const int *mp_n_dev;
    // readAScan in front (main.c:71-95), nullptr: the scan is already there
    const float *ranges, *cos_a, *sin_a;
    int lidar_n, max_range;
    float range_min;
    float *scan_x_out, *scan_y_out;
    b200slam_ctx::FrontOut *front;
    int trace;                // diagnostics (B200SLAM_FM_TRACE): SM cycle counter at the phase boundaries -> host_result->trace
};

inline int fastmatch_cluster_size(int nth) { return nth < FM_MAX_CLUSTER ? nth : FM_MAX_CLUSTER; }

// Progress counters in shared memory, one per group of FM_GROUP chunks: gather warps add one per (chunk, theta) item
// behind their stores, the summing warp polls before it starts on a group.
__device__ __forceinline__ void fm_group_done(unsigned int *counter)
{
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" :: "r"((unsigned int)__cvta_generic_to_shared(counter)) : "memory");
}
__device__ __forceinline__ unsigned int fm_group_progress(const unsigned int *counter)
{
    unsigned int v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"((unsigned int)__cvta_generic_to_shared(counter)) : "memory");
    return v;
}
__device__ __forceinline__ void fm_gather_barrier()               // the gather warps of one CTA among themselves
{
    asm volatile("bar.sync 1, %0;" :: "n"(FM_GATHER_WARPS * 32) : "memory");
}
__device__ __forceinline__ unsigned int fm_cluster_rank()
{
    unsigned int r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned int fm_cluster_size()
{
    unsigned int r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// address of this CTA's shared-memory object `p` in CTA `rank` of the cluster
__device__ __forceinline__ unsigned int fm_peer_smem(const void *p, unsigned int rank)
{
    unsigned int a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"((unsigned int)__cvta_generic_to_shared(p)), "r"(rank));
    return a;
}
__device__ __forceinline__ void fm_cluster_sync()                 // releases the DSMEM stores in front of it
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// cell of (int)roundf(v) when the 1-based index passes `1 < S < n` (main.c:512), else a value so negative that
// its sum with any valid offset stays negative
__device__ __forceinline__ int cell_or_invalid(float v, int n_minus_2)
{
    const int r = round_cell(v);
    return ((unsigned)(r - 1) < (unsigned)n_minus_2) ? r : INVALID_OFF;
}

// What a CTA of the cluster owns: thetas rank, rank + C, ...; its l-th theta's candidates are local l * per_th + q.
struct FmShare {
    int rank, C, nth_loc, nloc, per_th;
    __device__ __forceinline__ int owner(int c) const { return (c / per_th) % C; }
    __device__ __forceinline__ int local(int c) const { return (c / per_th) / C * per_th + c % per_th; }
    __device__ __forceinline__ int global(int lc) const { return ((lc / per_th) * C + rank) * per_th + lc % per_th; }
};

// The gathers of one pass by gather warp g.  NX, NY > 0: translation counts known at compile time.
template <int NX, int NY>
__device__ __forceinline__ void fm_gather(const FmArgs &A, const FmMap &M, const FmShare &S, int nbeams, int nchunk, const float *tab_s,
                                          const float2 *ps_s, float *vals, unsigned int *balm, unsigned int *group_done, int g, int lane)
{
    const int ntx = NX > 0 ? NX : A.ntx, nty = NY > 0 ? NY : A.nty;
    const float *field = M.field;
    const int pitch = M.pitch, cols2 = M.cols - 2, rows2 = M.rows - 2, nbp = A.nbp;
    constexpr int NXA = NX > 0 ? NX : 1, NYA = NY > 0 ? NY : 1;
    [[maybe_unused]] float sxt[NXA], syt[NYA];
    if constexpr (NX > 0 && NY > 0) {
#pragma unroll
        for (int k = 0; k < NX; ++k) sxt[k] = tab_s[2 * FM_MAX_CAND + k];
#pragma unroll
        for (int k = 0; k < NY; ++k) syt[k] = tab_s[3 * FM_MAX_CAND + k];
    }
    // work items (chunk k, local theta l) in chunk order, FM_GATHER_WARPS apart
    const int dk = FM_GATHER_WARPS / S.nth_loc, dl = FM_GATHER_WARPS - dk * S.nth_loc;
    int k = g / S.nth_loc, l = g - k * S.nth_loc;
    for (; k < nchunk; k += dk, l += dl) {
        if (l >= S.nth_loc) { l -= S.nth_loc; if (++k >= nchunk) break; }
        const int i = (k << 5) + lane, j = l * S.C + S.rank;
        const float ct = tab_s[j], st = tab_s[FM_MAX_CAND + j];
        const float2 ps = ps_s[i];
        const float Sx = rot_x(ps.x, ps.y, ct, st), Sy = rot_y(ps.x, ps.y, ct, st);                     // main.c:462-463
        float *dst = vals + (size_t)l * S.per_th * nbp + i;
        unsigned int *bm = balm + k * S.nloc + l * S.per_th;
        const bool beam = i < nbeams;                                                                   // else padding of the last group
        if constexpr (NX > 0 && NY > 0) {
            int c[NX], ro[NY];
#pragma unroll
            for (int kx = 0; kx < NX; ++kx) {
                c[kx] = cell_or_invalid(__fadd_rn(Sx, sxt[kx]), cols2);                                 // main.c:483
                if (!beam) c[kx] = INVALID_OFF;
            }
#pragma unroll
            for (int ky = 0; ky < NY; ++ky) {
                const int r = cell_or_invalid(__fadd_rn(Sy, syt[ky]), rows2);                           // main.c:501
                ro[ky] = r < 0 ? INVALID_OFF : r * pitch;
            }
            float v[NX * NY];
            unsigned int bal[NX * NY];
#pragma unroll
            for (int kx = 0; kx < NX; ++kx)
#pragma unroll
                for (int ky = 0; ky < NY; ++ky) {
                    const int idx = max(ro[ky] + c[kx], -1);                                            // either invalid: field[-1] == +0.0f
                    v[kx * NY + ky] = __ldg(field + idx);
                    bal[kx * NY + ky] = __ballot_sync(0xffffffffu, idx >= 0);                           // main.c:512
                }
#pragma unroll
            for (int q = 0; q < NX * NY; ++q) dst[q * nbp] = v[q];
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < NX * NY; ++q) bm[q] = bal[q];
            }
        } else {
            for (int kx = 0; kx < ntx; ++kx) {
                int c = cell_or_invalid(__fadd_rn(Sx, tab_s[2 * FM_MAX_CAND + kx]), cols2);             // main.c:483
                if (!beam) c = INVALID_OFF;
                for (int ky = 0; ky < nty; ++ky) {
                    const int r = cell_or_invalid(__fadd_rn(Sy, tab_s[3 * FM_MAX_CAND + ky]), rows2);   // main.c:501
                    const int idx = max((r < 0 ? INVALID_OFF : r * pitch) + c, -1);
                    const int q = kx * nty + ky;
                    dst[q * nbp] = __ldg(field + idx);
                    const unsigned int bal = __ballot_sync(0xffffffffu, idx >= 0);                      // main.c:512
                    if (lane == 0) bm[q] = bal;
                }
            }
        }
        __syncwarp();
        if (lane == 0) fm_group_done(group_done + k / FM_GROUP);
    }
}

__global__ void __launch_bounds__(FM_THREADS, 1) fastmatch_kernel(const __grid_constant__ FmArgs A, const __grid_constant__ LatticeTables T)
{
    extern __shared__ __align__(16) unsigned char fm_smem[];
    const int ncand = A.nth * A.ntx * A.nty;
    FmShare S;
    S.rank = (int)fm_cluster_rank(); S.C = (int)fm_cluster_size(); S.per_th = A.ntx * A.nty;
    S.nth_loc = (A.nth - S.rank + S.C - 1) / S.C; S.nloc = S.nth_loc * S.per_th;
    const int nloc_cap = (A.nth + S.C - 1) / S.C * S.per_th;                           // the same carve-up in every CTA
    const int chunk_cap = A.nbp >> 5;                                                  // 32-beam chunks per scan, upper bound
    float *vals = reinterpret_cast<float *>(fm_smem);                                 // [nloc][nbp]
    float2 *raw_s = reinterpret_cast<float2 *>(vals + (size_t)nloc_cap * A.nbp);        // [nbp]: the scan
    float2 *ps_s = raw_s + A.nbp;                                                      // [nbp]: the scan in pixels of this pass's map
    unsigned int *balm = reinterpret_cast<unsigned int *>(ps_s + A.nbp);               // [chunk][nloc]: in-bounds beams of a chunk
    int *pre = reinterpret_cast<int *>(balm + nloc_cap * chunk_cap);                   // [nloc][chunk_cap]: hits in front of a chunk
    __shared__ unsigned int group_done[FM_MAX_CHUNKS / FM_GROUP];                      // (chunk, theta) items of a group gathered so far, over the passes
    __shared__ __align__(8) unsigned long long xkey_s[2][FM_MAX_CLUSTER];              // per pass parity: best key of every CTA
    __shared__ int xcnt_s[2][FM_MAX_CAND];                                             // per pass parity: hits of every candidate
    __shared__ float tab_s[4 * FM_MAX_CAND];
    __shared__ int nscan_s;
    __shared__ int wcount_s[FM_WARPS];
    __shared__ long long trace_s[16];             // kept on chip while the kernel runs: a store to host memory per phase would perturb it

    pdl_launch_dependents();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool lead = S.rank == 0;
    int ntrace = 0;
#define FM_TRACE() do { if (A.trace && lead && tid == 0 && ntrace < 12) trace_s[ntrace++] = clock64(); } while (0)
    FM_TRACE();
    if (tid < FM_MAX_CHUNKS / FM_GROUP) group_done[tid] = 0;
    // Everything this kernel reads (scan, beam count, match state) and writes (bestHits[] twin, match state) may
    // belong to the kernel in front until it has completed; being resident already saves the launch latency.
    pdl_wait_prior_grids();
    // Chain mode: one launch runs chain_count consecutive scans (a cluster barrier between them carries the committed
    // pose from the first CTA to the others); everything else runs the body once.
    for (int it = 0;; ++it) {
    const int scan_index = A.scan_index + it;
    [[maybe_unused]] float guess[3] = {0.0f, 0.0f, 0.0f};
    [[maybe_unused]] ChainSlot out = {};              // chain mode, first CTA's thread 0: this scan's result for the host
    if (A.chain) {
        // the chain's state only changes in a tail, behind a kernel boundary or the cluster barrier at the end of the
        // loop: the same answer in every CTA
        const volatile ChainDev *cs = A.chain;
        if (cs->stop != 0 || cs->next_scan != scan_index) return;
        if (it > 0 && tid < FM_MAX_CHUNKS / FM_GROUP) group_done[tid] = 0;       // (a __syncthreads follows before its first use)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            // constant-velocity motion model, main.c:875-898: pose_guess = pose + DiffPose(previous_pose, pose)
            const float p = cs->pose[i];
            guess[i] = cs->have_prev ? __fadd_rn(p, __fsub_rn(p, cs->prev[i])) : p;
        }
    }
    int nbeams = A.nbeams_dev ? *A.nbeams_dev : A.nbeams;
    const float *ranges = A.ranges ? A.ranges + (size_t)it * A.lidar_n : nullptr;
    if (A.ranges) {
        // ---- readAScan (main.c:71-95): drop r < range_min | r > max_range, x = r cos, y = r sin, compacted in
        // beam order (the same arithmetic as scan_read_kernel, frontend.cu); every CTA for itself, the first one
        // for the kernels behind -------------------------------------------------------------------------------
        const float maxr = (float)A.max_range;
        int base = 0;
        for (int i0 = 0; i0 < A.lidar_n; i0 += FM_THREADS) {
            const int i = i0 + tid;
            float r = 0.0f;
            bool keep = false;
            if (i < A.lidar_n) {
                r = ranges[i];
                keep = !((r < A.range_min) | (r > maxr));                             // main.c:78
            }
            const unsigned int m = __ballot_sync(0xffffffffu, keep);
            __syncthreads();
            if (lane == 0) wcount_s[warp] = __popc(m);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll 8
            for (int w = 0; w < FM_WARPS; ++w) {
                const int c = wcount_s[w];
                before += w < warp ? c : 0;
                total += c;
            }
            if (keep) {
                const int slot = base + before + __popc(m & ((1u << lane) - 1u));
                const float x = __fmul_rn(r, A.cos_a[i]), y = __fmul_rn(r, A.sin_a[i]);   // main.c:90-91
                raw_s[slot] = make_float2(x, y);
                if (lead) { A.scan_x_out[slot] = x; A.scan_y_out[slot] = y; }
            }
            base += total;
        }
        if (tid == 0) {
            nscan_s = base;
            if (lead) { A.front->count = base; A.front->scan_n = base; }
        }
        __syncthreads();
        nbeams = nscan_s;
    } else {
        for (int i = tid; i < nbeams; i += FM_THREADS) raw_s[i] = make_float2(A.scan_x[i], A.scan_y[i]);
    }
    const int nchunk = (nbeams + 31) >> 5;                                            // 32-beam chunks with beams in them
    const int ngroup = (nchunk + FM_GROUP - 1) / FM_GROUP, nchunk_pad = ngroup * FM_GROUP;   // gathered and added: whole groups
    FM_TRACE();

    unsigned long long seed = ~0ull, key = ~0ull;
    int bh = 0, lh = 0, written = 0;
    for (int pass = 0; pass < A.npass; ++pass) {
        const FmMap M = A.map[pass];
        const int par = pass & 1;
        // ---- axis tables of this pass and the scan in pixels ---------------------------------------------------
        {
            const float *ctT = T.v, *stT = T.v + A.nth, *sxtT = T.v + 2 * A.nth, *sytT = sxtT + A.ntx;
            if (pass == 1 || A.seeded0) {
                // centred on the winner of the match in front (main.c:909-918): the host sent the axis tables of all
                // three possible centres per axis; see lattice_kernel
                const float *tb = T.v + (pass == 1 ? FM_TAB_B : 0);
                seed = pass == 0 ? *reinterpret_cast<volatile unsigned long long *>(&A.match->key) : key;
                const int lin1 = seed == ~0ull ? 13 : (int)(seed & 0xffffffffull);
                const int ith1 = lin1 / 9, itx1 = (lin1 / 3) % 3, ity1 = lin1 % 3;
                ctT = tb + 3 * ith1; stT = tb + 9 + 3 * ith1; sxtT = tb + 18 + 3 * itx1; sytT = tb + 27 + 3 * ity1;
            }
            __syncthreads();                          // raw_s filled; vals / balm / pre / tab_s of the previous pass done with
            if (A.chain) {
                // b200slam_lattice_value: p + (float)(k - n / 2) * s, the product rounded on its own (main.c:424-426)
                auto lv = [](float p, float st, int k) { return __fadd_rn(p, __fmul_rn((float)(k - 1), st)); };
                float c[3] = {guess[0], guess[1], guess[2]};
                if (pass == 1) {                      // FastMatch2 starts from FastMatch's result (main.c:918)
                    seed = key;
                    const int lin1 = (int)(key & 0xffffffffull);
                    c[0] = lv(guess[0], A.step_a[0], (lin1 / 3) % 3);
                    c[1] = lv(guess[1], A.step_a[1], lin1 % 3);
                    c[2] = lv(guess[2], A.step_a[2], lin1 / 9);
                }
                const float *stp = pass == 0 ? A.step_a : A.step_b;
                // 12 table entries, 12 threads (a cosf / sinf is ~150 dependent double-precision operations: the six of
                // them side by side, on the critical path of every pass)
                if (tid < 12) {
                    const int k = tid % 3, what = tid / 3;
                    if (what < 2) {
                        const float th = lv(c[2], stp[2], k);
                        tab_s[what * FM_MAX_CAND + k] = glibc_trig::sincos(th, what == 0 ? 1 : 0);   // main.c:434-435
                    } else if (what == 2) {
                        tab_s[2 * FM_MAX_CAND + k] = __fmul_rn(__fsub_rn(lv(c[0], stp[0], k), M.tlx), M.ipixel);   // main.c:436
                    } else {
                        tab_s[3 * FM_MAX_CAND + k] = __fmul_rn(__fsub_rn(lv(c[1], stp[1], k), M.tly), M.ipixel);   // main.c:437
                    }
                }
            } else {
                if (tid < A.nth) { tab_s[tid] = ctT[tid]; tab_s[FM_MAX_CAND + tid] = stT[tid]; }
                if (tid < A.ntx) tab_s[2 * FM_MAX_CAND + tid] = sxtT[tid];
                if (tid < A.nty) tab_s[3 * FM_MAX_CAND + tid] = sytT[tid];
            }
            for (int i = tid; i < (nchunk_pad << 5); i += FM_THREADS) {
                const float2 p = i < nbeams ? raw_s[i] : make_float2(0.0f, 0.0f);                    // padding of the last group: any finite value
                ps_s[i] = make_float2(__fmul_rn(p.x, M.ipixel), __fmul_rn(p.y, M.ipixel));           // main.c:418-419
            }
            __syncthreads();
        }
        FM_TRACE();

        if (warp == 0) {
            // ---- sums: one thread per candidate of this CTA, beams in scan order (main.c:516), chunk by chunk behind
            // the gather; half a chunk ahead in registers ----------------------------------------------------------
            const float4 *v = reinterpret_cast<const float4 *>(vals + (size_t)(lane < S.nloc ? lane : 0) * A.nbp);   // 16-byte aligned: nbp % 4 == 0
            float s = 0.0f;                                                           // main.c:507
            unsigned int spins = 0;
            const unsigned int target = (unsigned int)(FM_GROUP * S.nth_loc * (pass + 1));   // the counters run on over the passes
            for (int gq = 0; gq < ngroup; ++gq) {
                // the only branch on the chain: once per FM_GROUP chunks; the adds of a group are straight-line code
                // with its loads spread between them (a dependent FADD issues every 4 cycles, 3 slots are free)
                while (fm_group_progress(group_done + gq) < target)
                    if (++spins > (1u << 24)) { spins = ~0u - (1u << 25); break; }
                const float4 *p = v + gq * (FM_GROUP * 8);
#pragma unroll
                for (int q = 0; q < FM_GROUP * 8; ++q) {
                    const float4 x = p[q];
                    s = __fadd_rn(s, x.x); s = __fadd_rn(s, x.y); s = __fadd_rn(s, x.z); s = __fadd_rn(s, x.w);
                }
            }
            if (spins > (1u << 24) && lane == 0) atomicOr(&A.match->error, DEV_ERR_BARRIER);   // a gather warp never arrived
            // ---- this CTA's best (lowest score, then lowest index: strict `<` in loop order, main.c:549) -> every CTA --
            unsigned long long mine = lane < S.nloc ? pack_key(s, (unsigned int)S.global(lane)) : ~0ull;
            mine = warp_min_u64(mine);
            if (lane < S.C)
                asm volatile("st.shared::cluster.u64 [%0], %1;" :: "r"(fm_peer_smem(&xkey_s[par][S.rank], lane)), "l"(mine) : "memory");
        } else if ((warp & 3) != 0) {
            const int g = warp - 1 - (warp >> 2);                                     // 0 .. FM_GATHER_WARPS - 1
            // ---- gather ---------------------------------------------------------------------------------------------
            if (A.ntx == 3 && A.nty == 3) fm_gather<3, 3>(A, M, S, nbeams, nchunk_pad, tab_s, ps_s, vals, balm, group_done, g, lane);
            else fm_gather<0, 0>(A, M, S, nbeams, nchunk_pad, tab_s, ps_s, vals, balm, group_done, g, lane);
            if (A.trace && lead && lane == 0 && g == 0) trace_s[12 + pass] = clock64();
            fm_gather_barrier();
            // ---- hits: per candidate, the in-bounds beams (main.c:512-516) in front of every chunk and in total,
            // the total to every CTA ---------------------------------------------------------------------------------
            for (int lc = g; lc < S.nloc; lc += FM_GATHER_WARPS) {
                int total = 0;
                for (int k0 = 0; k0 < nchunk; k0 += 32) {
                    const int k = k0 + lane;
                    const int x = k < nchunk ? __popc(balm[k * S.nloc + lc]) : 0;
                    int incl = x;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += o;
                    }
                    if (k < nchunk) pre[lc * chunk_cap + k] = total + incl - x;
                    total += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane < S.C)
                    asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(fm_peer_smem(&xcnt_s[par][S.global(lc)], lane)), "r"(total) : "memory");
            }
        }
        fm_cluster_sync();
        FM_TRACE();
        // ---- the winner, and the hit counts of all candidates ------------------------------------------------------
        key = ~0ull;
        for (int r = 0; r < S.C; ++r) key = min(key, xkey_s[par][r]);
        const int cnt = lane < ncand ? xcnt_s[par][lane] : -1;
        bh = xcnt_s[par][(int)(key & 0xffffffffull)];                                 // bestHits_size: the winner's (main.c:557)
        lh = xcnt_s[par][ncand - 1];
        // main.c:515: every candidate overwrites bestHits[] from index 0, so the array ends up holding, behind the last
        // candidate's hits, those of the most recent candidate that had more.  Candidate c still owns entries
        // [suffix max of the counts behind it, its own count): lane c works that out by shuffles
        int sfx = cnt;                                                                // max over lanes >= this one
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_down_sync(0xffffffffu, sfx, d);
            if (lane + d < 32) sfx = max(sfx, o);
        }
        written = __shfl_sync(0xffffffffu, sfx, 0);                                   // the longest candidate
        if (warp != 0 && (warp & 3) != 0) {
            const int g = warp - 1 - (warp >> 2);
            int behind = __shfl_down_sync(0xffffffffu, sfx, 1);                       // max over lanes > this one
            if (lane == 31) behind = -1;
            const bool stair = lane < ncand && (lane == ncand - 1 || cnt > behind) && S.owner(lane) == S.rank;
            const int lo_l = lane == ncand - 1 ? 0 : behind;
            unsigned int stairs = __ballot_sync(0xffffffffu, stair);                  // this CTA's share
            while (stairs) {
                const int c = __ffs(stairs) - 1;
                stairs &= stairs - 1;
                const int lo = __shfl_sync(0xffffffffu, lo_l, c), lc = S.local(c);
                const float *v = vals + (size_t)lc * A.nbp;
                for (int k = g; k < nchunk; k += FM_GATHER_WARPS) {
                    const unsigned int bal = balm[k * S.nloc + lc];
                    const int pos = pre[lc * chunk_cap + k] + __popc(bal & ((1u << lane) - 1u));
                    if (((bal >> lane) & 1u) && pos >= lo) A.hit_values[pos] = v[(k << 5) + lane];
                }
            }
        }
    }
    if (lead && tid == 0) {
        A.match->key = key;
        A.match->best_hits = bh;
        A.match->last_hits = lh;
        A.match->written_hits = written;
        if (A.npass > 1 || A.seeded0) A.match->seed_key = seed;
        if (A.host_result) {
            volatile MatchHost *h = A.host_result;
            h->key = key; h->seed_key = (A.npass > 1 || A.seeded0) ? seed : ~0ull;
            h->best_hits = bh; h->last_hits = lh; h->written_hits = written;
            h->scan_n = nbeams;
            h->mp_n = A.mp_n_dev ? *A.mp_n_dev : 0;
            h->error = *reinterpret_cast<volatile unsigned int *>(&A.match->error);
            FM_TRACE();
            if (A.trace)                          // [0, 12): phase boundaries seen by thread 0; [12 + pass]: end of gather warp 0's last chunk
                for (int i = 0; i < 16; ++i) h->trace[i] = (i < ntrace || (i >= 12 && i < 12 + A.npass)) ? trace_s[i] : 0;
            __threadfence_system();
            h->seq = A.host_seq;
        }
        if (A.chain) {
            // ---- commit: the refined pose (main.c:592-594 of both matches, :920-922), the path's previous entry, the
            // mini-update test (main.c:928-940); the result goes to the host through the ring, seq last ------------------
            auto lv = [](float p, float st, int k) { return __fadd_rn(p, __fmul_rn((float)(k - 1), st)); };
            ChainDev *cs = A.chain;
            const int l1 = (int)(seed & 0xffffffffull), l2 = (int)(key & 0xffffffffull);
            const int ia[3] = {(l1 / 3) % 3, l1 % 3, l1 / 9}, ib[3] = {(l2 / 3) % 3, l2 % 3, l2 / 9};
            bool update = false;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float pa = lv(guess[i], A.step_a[i], ia[i]);
                const float pb = lv(pa, A.step_b[i], ib[i]);
                const float dp = fabsf(__fsub_rn(pb, cs->map_pose[i]));                              // DiffPose(map.pose, pose)
                update = update || dp > (i < 2 ? A.mini_dt : A.mini_dr);
                cs->prev[i] = cs->pose[i];
                cs->pose[i] = pb;
                out.pose_a[i] = pa; out.pose_b[i] = pb;
            }
            cs->have_prev = 1;
            cs->next_scan = scan_index + 1;
            if (update) cs->stop = 1;
            out.scan_n = nbeams; out.best_hits = bh;
            out.mp_n = A.mp_n_dev ? *A.mp_n_dev : 0;
            out.stopped = update ? 1 : 0;
            out.error = *reinterpret_cast<volatile unsigned int *>(&A.match->error);
        }
    }
    const bool more = A.chain && it + 1 < A.chain_count;
    if (more) fm_cluster_sync();                      // the commit above is visible to every CTA; shared memory is free again
    if (A.chain && lead && tid == 0) {
        // the host's copy travels off the critical path: the other warps are already reading the next scan
        volatile ChainSlot *slot = A.chain_ring + (scan_index % CHAIN_RING);
#pragma unroll
        for (int i = 0; i < 3; ++i) { slot->pose_a[i] = out.pose_a[i]; slot->pose_b[i] = out.pose_b[i]; }
        slot->scan_n = out.scan_n; slot->best_hits = out.best_hits; slot->mp_n = out.mp_n; slot->stopped = out.stopped;
        slot->error = out.error;
        __threadfence_system();
        slot->seq = (unsigned long long)scan_index + 1ull;
    }
    if (!more) break;
    }
#undef FM_TRACE
}

// beams per row of vals: whole groups of chunks, + 4 so that rows stay 16-byte aligned and start 4 banks apart
inline int fastmatch_row_pitch(int nbeams) { return ((nbeams + 32 * FM_GROUP - 1) / (32 * FM_GROUP)) * (32 * FM_GROUP) + 4; }

// dynamic shared memory of one CTA of the cluster
size_t fastmatch_smem_bytes(int ncand, int nth, int nbeams)
{
    const int C = fastmatch_cluster_size(nth), per_th = ncand / nth;
    const size_t nloc = (size_t)((nth + C - 1) / C) * per_th;
    const size_t nbp = fastmatch_row_pitch(nbeams);
    return sizeof(float) * ((nloc + 4) * nbp + 2 * nloc * (nbp >> 5));
}

int fastmatch_launch_args(b200slam_ctx *ctx, FmArgs &A, const LatticeTables &T)
{
    A.nbp = fastmatch_row_pitch(A.nbeams);
    A.trace = getenv("B200SLAM_FM_TRACE") ? 1 : 0;
    const size_t smem = fastmatch_smem_bytes(A.nth * A.ntx * A.nty, A.nth, A.nbeams);
    static bool smem_set[64] = {};
    if (!smem_set[ctx->device & 63]) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(fastmatch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        smem_set[ctx->device & 63] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(fastmatch_cluster_size(A.nth));  // one cluster, one CTA per theta (up to 8)
    cfg.blockDim = dim3(FM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cfg.gridDim.x; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (ctx->use_pdl && ctx->prev_launch_was_lattice) ? 2 : 1;
    CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, fastmatch_kernel, A, T));
    LAUNCH_CHECK(ctx);
    ctx->prev_launch_was_lattice = true;
    ctx->prev_lattice_field = A.map[A.npass - 1].field;
    return B200SLAM_OK;
}

int launch_fastmatch(b200slam_ctx *ctx, const LatticeArgs &L, const LatticeTables &T)
{
    FmArgs A = {};
    A.map[0].field = L.field; A.map[0].pitch = L.pitch; A.map[0].rows = L.rows; A.map[0].cols = L.cols; A.map[0].ipixel = L.ipixel;
    A.map[1] = A.map[0];
    A.npass = 1;
    A.seeded0 = L.seeded;
    A.scan_x = L.scan_x; A.scan_y = L.scan_y; A.nbeams = L.nbeams; A.nbeams_dev = L.nbeams_dev;
    A.nth = L.nth; A.ntx = L.ntx; A.nty = L.nty;
    A.match = L.match;
    A.hit_values = L.hit_values + L.hit_stride;
    A.host_result = L.host_result; A.host_seq = L.host_seq; A.mp_n_dev = L.mp_n_dev;
    A.ranges = nullptr;
    return fastmatch_launch_args(ctx, A, T);
}

template <int TYPT, int WX, int WY, int UM = 1, bool COUNT = false, int Q = 0, int PITCH = 0, int NB = 0>
int launch_lattice_cfg(b200slam_ctx *ctx, LatticeArgs &A, const LatticeTables &T, int nth_cover)
{
    constexpr int TXT = 32 * WX, TYT = TYPT * WY;
    auto kern = lattice_kernel<TYPT, WX, WY, UM, COUNT, Q, PITCH, NB>;
    // Beams per chunk: the whole scan when its tables fit the budget, else even chunks.
    constexpr int GUARD = LatticePipe<TYPT, UM, Q, NB>::GUARD;
    const int per_beam = (TXT + (Q > 0 ? WY : TYT) + 2) * 4;
    // 16 candidates per thread: chunks that leave room for 4 (row reuse) / 3 resident CTAs per SM
    const int budget = (TYPT >= 16 ? (Q > 0 ? 50 : 56) * 1024 : 100 * 1024) - (Q > 0 ? TYT * 4 : 0);
    int cb = A.nbeams > 0 ? A.nbeams : 1;
    const int cap = budget / per_beam - GUARD;
    if (cb > cap) {
        const int chunks = (cb + cap - 1) / cap;
        cb = (cb + chunks - 1) / chunks;
    }
    cb = (cb + 3) & ~3;                                   // keeps the int4 row loads aligned
    A.cb = cb;
    const size_t smem = (size_t)(cb + GUARD) * per_beam + (Q > 0 ? TYT * 4 : 0);
    // Opt in to large dynamic shared memory once per instantiation and device -- unconditionally: the
    // 48 KB default applies to static + dynamic together, so a chunk just under 48 KB needs it too.
    static bool smem_set[64] = {};
    if (!smem_set[ctx->device & 63]) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024));
        smem_set[ctx->device & 63] = true;
    }
    dim3 grid((A.nty + TYT - 1) / TYT, (A.ntx + TXT - 1) / TXT, nth_cover);
    A.total_ctas = grid.x * grid.y * grid.z;
    // PDL only behind another scan-matching kernel: everything else this kernel follows (EDT,
    // rasterisation, copies) produces data its early part reads, and keeps full stream order.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(32 * WX * WY);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (ctx->use_pdl && ctx->prev_launch_was_lattice && !A.scores) ? 1 : 0;
    CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kern, A, T));
    LAUNCH_CHECK(ctx);
    ctx->prev_launch_was_lattice = true;
    ctx->prev_lattice_field = A.field;
    return B200SLAM_OK;
}

}  // namespace

int lattice_launch(b200slam_ctx *ctx, const LatticeLaunch &L)
{
    const b200slam_map *m = L.map;
    LatticeArgs A;
    A.field = m->d_field; A.pitch = m->field_pitch; A.rows = m->rows; A.cols = m->cols;
    A.rr_clamp = -((B200SLAM_FIELD_PAD_ROWS - 1) * m->field_pitch + 1);
    A.scan_x = ctx->d_scan_x; A.scan_y = ctx->d_scan_y; A.nbeams = ctx->nbeams;
    A.nbeams_dev = ctx->scan_n_dev ? &ctx->d_front->scan_n : nullptr;
    A.seeded = L.seeded ? 1 : 0;
    A.host_result = L.host_result ? ctx->h_result : nullptr;
    A.host_seq = ctx->result_seq;
    A.mp_n_dev = (ctx->mp_n_dev && ctx->d_front) ? &ctx->d_front->mp_n : nullptr;
    A.ipixel = 1 / m->pixel_size;                                        // main.c:383
    A.nth = L.nth; A.ntx = L.ntx; A.nty = L.nty;
    A.row_begin = L.row_begin; A.row_end = L.row_end;
    A.match = ctx->d_match;
    A.scores = L.d_scores;
    A.hit_values = ctx->d_hit_values;
    A.hit_stride = ctx->scan_cap;
    A.tables = L.d_tables;
    A.xchg.peers = nullptr; A.xchg.nranks = 1; A.xchg.rank = 0; A.xchg.timeout_ns = ctx->spin_timeout_ns;
    A.collect_prev = L.collect_prev ? 1 : 0;
    A.post_deferred = L.post_deferred ? 1 : 0;
    if (L.exchange && ctx->p2p_ready) A.xchg = xchg_args(ctx);
    LatticeTables T;                     // parameter block (copied at launch)
    A.nth_tab = L.nth_tab;
    if (!L.d_tables) memcpy(T.v, L.h_tables, sizeof(float) * (L.seeded ? 36 : 2 * (size_t)L.nth_tab + L.ntx + L.nty));
    if (L.row_end <= L.row_begin) {
        if (A.xchg.peers) {             // nothing to score, but the peers wait for this rank's post
            exchange_only_kernel<<<1, 64, 0, ctx->stream>>>(ctx->d_match, A.xchg, A.post_deferred, A.collect_prev);
            LAUNCH_CHECK(ctx);
            return B200SLAM_OK;
        }
        // empty shard: publish "nothing scored"
        CUDA_TRY(ctx, cudaMemsetAsync(&ctx->d_match->key, 0xff, sizeof(unsigned long long), ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(&ctx->d_match->best_hits, 0, 2 * sizeof(int), ctx->stream));
        return B200SLAM_OK;
    }
    A.th_first = L.th_first;
    const int nth_cover = L.nth_tab;
    const long long cands = (long long)(L.row_end - L.row_begin) * L.nty;

    // Tile shape.  A thread owns TYPT candidates (consecutive ty); per warp and beam the inner loop
    // costs 1 (column LDS) + TYPT/4 (row LDS.128) + 1.5 * TYPT (gathers, ~1.5 lines each) L1
    // wavefronts, so big register tiles are cheaper per evaluation (the `per_eval` factors below,
    // relative to TYPT = 16) -- but a single sweep is only as fast as its busiest SM.  Latency
    // mode (default): every SM runs ceil(CTAs / SMs) tiles one after the other; pick the shape
    // with the lowest tiles-per-SM x candidates-per-tile x per-evaluation cost.  Throughput mode
    // (b200slam_set_match_mode: many independent matches in flight): total work only, CTAs x
    // candidates-per-tile x per-evaluation cost -- the other kernels in flight fill the SMs one
    // launch leaves idle.  Measured on config 1 (64 x 32 x 32 x 360 beams; one launch alone /
    // EDT+match step with the next EDT running under the match): 32 x 16 tiles with 16-beam
    // groups 24.9 / 31.6 us, 32 x 8 tiles 27.1 / 27.4 us, 32 x 32 tiles 33.3 / 23.2 us.
    // FastMatch-sized lattices (the reference's 3 x 3 x 3): one candidate per thread with per-candidate
    // hit counts, so that the tail can leave bestHits[] exactly as the reference's loop does.
    if ((L.seeded || !getenv("B200SLAM_LATTICE_CFG")) && (long long)L.nth * L.ntx * L.nty <= MATCH_SMALL && L.row_begin == 0 &&
        L.row_end == (int64_t)L.nth * L.ntx && !A.xchg.peers && L.nth_tab == L.nth) {
        // the reference's own size class: one CTA, gathers across all threads (fastmatch_kernel); scans longer than
        // its shared memory holds, > 32 candidates or a wanted score table take the general kernel
        const size_t fm_smem = fastmatch_smem_bytes(L.nth * L.ntx * L.nty, L.nth, A.nbeams);
        if ((long long)L.nth * L.ntx * L.nty <= FM_MAX_CAND && A.nbeams <= FM_MAX_BEAMS && fm_smem <= 216 * 1024 && !A.scores &&
            !L.d_tables && !getenv("B200SLAM_NO_FASTMATCH_KERNEL"))
            return launch_fastmatch(ctx, A, T);
        return launch_lattice_cfg<1, 1, 4, 1, true>(ctx, A, T, nth_cover);
    }
    if (L.seeded) return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "a seeded match is a whole 3 x 3 x 3 lattice on one GPU");
    // Row reuse (template parameter Q): usable when the ty step is pixel / Q, measured on the ty axis table
    // itself (host copy).  The kernel verifies every (beam, ty group) exactly, so a wrong guess here only
    // costs time, never a bit of the result.
    int q_ok = 0;
    if (!getenv("B200SLAM_LATTICE_NO_RR") && L.nty >= 4 && L.h_tables) {
        const float *syt = L.h_tables + 2 * (size_t)L.nth_tab + L.ntx;
        const double d = ((double)syt[L.nty - 1] - (double)syt[0]) / (L.nty - 1);
        for (int q : {2, 4})
            if (d > 0 && fabs(d * q - 1.0) < 1e-3) q_ok = q;
    }
    struct Shape { int typt, wx, wy, um, q; double per_eval; };
    static const Shape shapes[] = {
        {16, 2, 4, 1, 0, 1.00}, {8, 1, 8, 1, 0, 1.05}, {4, 1, 8, 1, 0, 1.13}, {2, 1, 8, 2, 0, 1.40}, {1, 1, 8, 1, 0, 1.95},
        {1, 1, 4, 1, 0, 1.95},
        // row reuse: K = TYPT / Q + 1 gathers per TYPT candidates.  Measured on config 3 (256 x 128 x 128 x 1080,
        // step = pixel / 2): 64 x 64 tiles 648 us against 1059 us without reuse, 32 x 64 tiles 1033 us; with 4
        // candidates per thread the per-beam bookkeeping outweighs the saved gathers (1503 us), so those
        // shapes are compiled (and tested) but never chosen.  The Q = 4 factors are estimates.
        {16, 2, 4, 1, 2, 0.50}, {8, 1, 8, 1, 2, 0.98},
        {16, 2, 4, 1, 4, 0.55}, {8, 1, 8, 1, 4, 0.85},
    };
    int pick_t = 0, pick_x = 0, pick_y = 0, pick_m = 1, pick_q = 0;
    if (const char *e = getenv("B200SLAM_LATTICE_CFG"))                   // tuning / test aid: "TYPT,WX,WY[,UM[,Q]]"
        if (sscanf(e, "%d,%d,%d,%d,%d", &pick_t, &pick_x, &pick_y, &pick_m, &pick_q) < 3) pick_t = 0;
    if (!pick_t) {
        const double covered = (double)(L.row_end - L.row_begin) / ((double)nth_cover * L.ntx);   // row shards: part of the grid
        double best_cost = 1e300;
        for (const Shape &sh : shapes) {
            if (sh.q != 0 && sh.q != q_ok) continue;
            const int txt = 32 * sh.wx, tyt = sh.typt * sh.wy;
            double ctas = (double)nth_cover * ((L.ntx + txt - 1) / txt) * ((L.nty + tyt - 1) / tyt) * covered;
            if (ctas < 1.0) ctas = 1.0;
            const double per_sm = (double)((long long)((ctas + ctx->sm_count - 1) / ctx->sm_count));
            const double units = ctx->match_mode == B200SLAM_MATCH_THROUGHPUT ? ctas : (per_sm < 1.0 ? 1.0 : per_sm);
            const double cost = units * txt * tyt * sh.per_eval;
            if (cost < best_cost * 0.999) {                                // ties: the bigger register tile
                best_cost = cost;
                pick_t = sh.typt; pick_x = sh.wx; pick_y = sh.wy; pick_m = sh.um; pick_q = sh.q;
            }
        }
    }
#define B200SLAM_CFG(T_, X_, Y_, M_, Q_) \
    if (pick_t == T_ && pick_x == X_ && pick_y == Y_ && pick_m == M_ && pick_q == Q_) \
        return launch_lattice_cfg<T_, X_, Y_, M_, false, Q_>(ctx, A, T, nth_cover);
    if (pick_t == 16 && pick_x == 2 && pick_y == 4 && pick_m == 1 && pick_q == 2 && !getenv("B200SLAM_LATTICE_NO_PITCH")) {
        // row reuse on a map whose pitch is a compile-time constant of the kernel (see PITCH)
        // (a third buffer spills at 64 registers: 232 bytes in the loop)
        if (A.pitch == 16384) return launch_lattice_cfg<16, 2, 4, 1, false, 2, 16384>(ctx, A, T, nth_cover);
        if (A.pitch == 8192) return launch_lattice_cfg<16, 2, 4, 1, false, 2, 8192>(ctx, A, T, nth_cover);
        if (A.pitch == 4096) return launch_lattice_cfg<16, 2, 4, 1, false, 2, 4096>(ctx, A, T, nth_cover);
        if (A.pitch == 2048) return launch_lattice_cfg<16, 2, 4, 1, false, 2, 2048>(ctx, A, T, nth_cover);
        if (A.pitch == 1024) return launch_lattice_cfg<16, 2, 4, 1, false, 2, 1024>(ctx, A, T, nth_cover);
    }
    B200SLAM_CFG(16, 2, 4, 1, 0)        // 64 x 64 tile, 256 threads
    B200SLAM_CFG(8, 1, 8, 1, 0)         // 32 x 64
    B200SLAM_CFG(4, 1, 8, 1, 0)         // 32 x 32
    B200SLAM_CFG(2, 1, 8, 2, 0)         // 32 x 16, groups of 16 beams
    B200SLAM_CFG(2, 1, 8, 1, 0)
    B200SLAM_CFG(1, 1, 8, 1, 0)         // 32 x 8
    B200SLAM_CFG(1, 1, 4, 1, 0)         // 32 x 4, 128 threads
    B200SLAM_CFG(16, 2, 4, 1, 2) B200SLAM_CFG(8, 1, 8, 1, 2) B200SLAM_CFG(4, 1, 8, 1, 2)      // row reuse, step = pixel / 2
    B200SLAM_CFG(16, 2, 4, 1, 4) B200SLAM_CFG(8, 1, 8, 1, 4) B200SLAM_CFG(4, 1, 8, 1, 4)      // row reuse, step = pixel / 4
#undef B200SLAM_CFG
    return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "B200SLAM_LATTICE_CFG names no compiled tile shape");
}

// readAScan + FastMatch + FastMatch2 as ONE kernel (b200slam_scan_step_async).  tables: 12 floats of the first
// lattice at [0], the 36 floats of the seeded second lattice at [FM_TAB_B].
int scan_step_launch(b200slam_ctx *ctx, const b200slam_map *ma, const b200slam_map *mb, const float *tables12,
                     const float *tables36, const float *d_ranges, int max_range)
{
    FmArgs A = {};
    const b200slam_map *mm[2] = {ma, mb};
    for (int p = 0; p < 2; ++p) {
        A.map[p].field = mm[p]->d_field; A.map[p].pitch = mm[p]->field_pitch; A.map[p].rows = mm[p]->rows;
        A.map[p].cols = mm[p]->cols; A.map[p].ipixel = 1 / mm[p]->pixel_size;                 // main.c:383
    }
    A.npass = 2;
    A.seeded0 = 0;
    A.scan_x = ctx->d_scan_x; A.scan_y = ctx->d_scan_y;
    A.nbeams = ctx->lidar_n; A.nbeams_dev = nullptr;
    A.nth = A.ntx = A.nty = 3;
    A.match = ctx->d_match;
    A.hit_values = ctx->d_hit_values + ctx->scan_cap;
    A.host_result = ctx->h_result; A.host_seq = ctx->result_seq;
    A.mp_n_dev = (ctx->mp_n_dev && ctx->d_front) ? &ctx->d_front->mp_n : nullptr;
    A.ranges = d_ranges; A.cos_a = ctx->d_lidar; A.sin_a = ctx->d_lidar + ctx->lidar_n;
    A.lidar_n = ctx->lidar_n; A.max_range = max_range; A.range_min = ctx->lidar_range_min;
    A.scan_x_out = ctx->d_scan_x; A.scan_y_out = ctx->d_scan_y;
    A.front = ctx->d_front;
    if (A.nbeams > FM_MAX_BEAMS || fastmatch_smem_bytes(27, 3, A.nbeams) > 216 * 1024)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "scan of %d beams exceeds the fused scan step (%d)", A.nbeams, FM_MAX_BEAMS);
    LatticeTables T;
    memcpy(T.v, tables12, sizeof(float) * 12);
    memcpy(T.v + FM_TAB_B, tables36, sizeof(float) * 36);
    return fastmatch_launch_args(ctx, A, T);
}

// One scan of the device-resident loop: the fused kernel in chain mode (no tables from the host, no pose argument).
int scan_chain_launch(b200slam_ctx *ctx, const b200slam_map *ma, const b200slam_map *mb, const ChainLaunch &C,
                      const float *d_ranges, int max_range)
{
    FmArgs A = {};
    const b200slam_map *mm[2] = {ma, mb};
    for (int p = 0; p < 2; ++p) {
        A.map[p].field = mm[p]->d_field; A.map[p].pitch = mm[p]->field_pitch; A.map[p].rows = mm[p]->rows;
        A.map[p].cols = mm[p]->cols; A.map[p].ipixel = 1 / mm[p]->pixel_size;                 // main.c:383
        A.map[p].tlx = mm[p]->top_left_x; A.map[p].tly = mm[p]->top_left_y;
    }
    A.npass = 2;
    A.seeded0 = 0;
    A.scan_x = ctx->d_scan_x; A.scan_y = ctx->d_scan_y;
    A.nbeams = ctx->lidar_n; A.nbeams_dev = nullptr;
    A.nth = A.ntx = A.nty = 3;
    A.match = ctx->d_match;
    A.hit_values = ctx->d_hit_values + ctx->scan_cap;
    A.host_result = nullptr; A.host_seq = 0;
    A.mp_n_dev = (ctx->mp_n_dev && ctx->d_front) ? &ctx->d_front->mp_n : nullptr;
    A.ranges = d_ranges; A.cos_a = ctx->d_lidar; A.sin_a = ctx->d_lidar + ctx->lidar_n;
    A.lidar_n = ctx->lidar_n; A.max_range = max_range; A.range_min = ctx->lidar_range_min;
    A.scan_x_out = ctx->d_scan_x; A.scan_y_out = ctx->d_scan_y;
    A.front = ctx->d_front;
    A.chain = ctx->d_chain; A.chain_ring = ctx->h_chain_ring; A.scan_index = C.scan_index; A.chain_count = C.count;
    for (int i = 0; i < 3; ++i) { A.step_a[i] = C.step_a[i]; A.step_b[i] = C.step_b[i]; }
    A.mini_dt = ctx->chain_mini_dt; A.mini_dr = ctx->chain_mini_dr;
    if (A.nbeams > FM_MAX_BEAMS || fastmatch_smem_bytes(27, 3, A.nbeams) > 216 * 1024)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "scan of %d beams exceeds the fused scan step (%d)", A.nbeams, FM_MAX_BEAMS);
    LatticeTables T = {};
    // Programmatic dependent launch behind the scan in front: the kernel may become resident early (launch latency,
    // shared-memory set-up), but it reads the chain's state -- which the kernel in front commits in its tail -- and
    // everything else only behind its wait for that kernel's completion.
    return fastmatch_launch_args(ctx, A, T);
}

namespace {
__global__ void trig_probe_kernel(const float *__restrict__ in, int n, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = glibc_trig::sincos(in[i], 0);
    out[n + i] = glibc_trig::sincos(in[i], 1);
}
}  // namespace

// The device's restatement of glibc's sinf / cosf against the libm this process runs on, on a sample of the range the
// chain serves: every 2^-10 step of [-16, 16], the floats next to every multiple of pi/4 (the branch points of the
// range reduction) and 2^16 hashed values.  (tools/trig_check.cpp does the whole range on the CPU.)
int chain_trig_selftest(b200slam_ctx *ctx)
{
    if (ctx->chain_trig_ok) return ctx->chain_trig_ok > 0 ? B200SLAM_OK
                                   : b200slam_set_error(ctx, B200SLAM_ERR_STATE, "the device's cosf / sinf differ from this host's libm");
    const int N1 = 32 * 1024 + 1, N2 = 21 * 16, N3 = 1 << 16, N = N1 + N2 + N3;
    float *h = (float *)malloc(sizeof(float) * 3 * (size_t)N);
    if (!h) return B200SLAM_ERR_NOMEM;
    int n = 0;
    for (int i = 0; i < N1; ++i) h[n++] = -16.0f + (float)i * (1.0f / 1024.0f);
    for (int k = -10; k <= 10; ++k) {
        float v = (float)((double)k * 0.78539816339744830962);
        for (int j = 0; j < 8; ++j) v = nextafterf(v, -100.0f);
        for (int j = 0; j < 16; ++j) { h[n++] = v; v = nextafterf(v, 100.0f); }
    }
    unsigned long long z = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < N3; ++i) {
        z ^= z << 13; z ^= z >> 7; z ^= z << 17;
        h[n++] = (float)((double)(z >> 11) * (1.0 / 9007199254740992.0) * 32.0 - 16.0);
    }
    float *d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(float) * 3 * (size_t)N);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d, h, sizeof(float) * N, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        trig_probe_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(d, N, d + N);
        ctx->launches++;
        e = cudaMemcpyAsync(h + N, d + N, sizeof(float) * 2 * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) {
        free(h);
        return b200slam_set_error(ctx, B200SLAM_ERR_CUDA, "trig self-test -> %s", cudaGetErrorString(e));
    }
    int bad = 0;
    for (int i = 0; i < N; ++i) {
        const float s = sinf(h[i]), c = cosf(h[i]);
        if (memcmp(&s, &h[N + i], 4) || memcmp(&c, &h[2 * N + i], 4)) bad++;
    }
    free(h);
    ctx->chain_trig_ok = bad ? -1 : 1;
    return bad ? b200slam_set_error(ctx, B200SLAM_ERR_STATE, "the device's cosf / sinf differ from this host's libm (%d of %d samples)", bad, N)
               : B200SLAM_OK;
}

int exchange_collect_launch(b200slam_ctx *ctx)
{
    exchange_collect_kernel<<<1, 64, 0, ctx->stream>>>(ctx->d_match, xchg_args(ctx));
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}

int poses_launch(b200slam_ctx *ctx, const b200slam_map *m, int64_t P, int64_t index_base,
                 float *d_scores, int32_t *d_hits, bool post_sharded)
{
    PosesArgs A;
    A.field = m->d_field; A.pitch = m->field_pitch; A.rows = m->rows; A.cols = m->cols;
    A.field11 = m->d_field + m->field_pitch + 1; A.zero_off = -(m->field_pitch + 2);
    A.scan_x = ctx->d_scan_x; A.scan_y = ctx->d_scan_y; A.nbeams = ctx->nbeams;
    A.nbeams_dev = ctx->scan_n_dev ? &ctx->d_front->scan_n : nullptr;
    A.ipixel = 1 / m->pixel_size;
    A.min_x = m->top_left_x; A.min_y = m->top_left_y;
    A.px = ctx->d_pose_soa; A.py = A.px + ctx->pose_cap; A.ct = A.py + ctx->pose_cap;
    A.st = A.ct + ctx->pose_cap;
    A.P = P; A.index_base = index_base;
    A.scores = d_scores; A.hits = d_hits;
    A.match = ctx->d_match;
    A.xchg.peers = nullptr; A.xchg.nranks = 1; A.xchg.rank = 0; A.xchg.timeout_ns = ctx->spin_timeout_ns;
    if (post_sharded) {
        if (!ctx->p2p_ready || P <= 0)
            return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "a sharded particle set needs NVLink peer memory and particles on every rank");
        A.xchg = xchg_args(ctx);
    }
    if (P <= 0) {
        CUDA_TRY(ctx, cudaMemsetAsync(&ctx->d_match->key, 0xff, sizeof(unsigned long long), ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(&ctx->d_match->best_hits, 0, 2 * sizeof(int), ctx->stream));
        return B200SLAM_OK;
    }
    if (m->rows >= (1 << 22) || m->cols >= (1 << 22))
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "pose-list scoring needs map dimensions below 2^22 (%d x %d)",
                                  m->rows, m->cols);
    const unsigned grid = (unsigned)((P + POSES_THREADS - 1) / POSES_THREADS);
    A.total_ctas = grid;
    poses_kernel<<<grid, POSES_THREADS, 0, ctx->stream>>>(A);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}
