// particles.cu -- FastSLAM-style particle weights, normalisation and systematic
// resampling on top of the pose-list scores (extension asked for by BASELINE.json; the
// reference has no particle filter -- SURVEY.md section 0).  The definition is the one in
// include/b200slam.h / oracle/slam_oracle.c (orc_weights_resample):
//   w_i = exp_det(-beta * (score_i - score_min)),  q_i = (uint64)(w_i * 2^32),
//   W = sum q_i,  weight_i = (float)((double)q_i / W),
//   T_k = U + floor(k*W/N),  ancestor_k = first i with inclusive prefix C_i > T_k.
// Weights are integers, so the prefix sum is associative and the resampled indices are
// the same for any blocking, any GPU count and the CPU oracle.  exp_det uses only IEEE
// basic operations in double (no FMA contraction), so it is bit-identical to the host.
#include "common.cuh"

namespace {

constexpr int PT_THREADS = 256;
constexpr int PT_ITEMS = 4;
constexpr int PT_BLOCK = PT_THREADS * PT_ITEMS;     // particles per CTA

__device__ __forceinline__ float exp_det(float x)
{
    double xd = (double)x;
    if (!(xd > -80.0)) return 0.0f;
    if (xd > 0.0) xd = 0.0;
    const double LOG2E = 1.4426950408889634;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    const double kf = rint(__dmul_rn(xd, LOG2E));
    double r = __dsub_rn(xd, __dmul_rn(kf, LN2_HI));
    r = __dsub_rn(r, __dmul_rn(kf, LN2_LO));
    const double c[14] = {
        1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040,
        1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600,
        1.0 / 6227020800.0};
    double p = c[13];
#pragma unroll
    for (int i = 12; i >= 0; --i) p = __dadd_rn(__dmul_rn(p, r), c[i]);
    const int k = (int)kf;
    const double two_k = __longlong_as_double((long long)(1023 + k) << 52);
    return __double2float_rn(__dmul_rn(p, two_k));
}

__device__ __forceinline__ unsigned long long warp_incl_scan(unsigned long long v, int lane)
{
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, v, s);
        if (lane >= s) v += o;
    }
    return v;
}

// q_i and per-CTA sums.  keys: packed arg-min keys (one per rank); their min carries
// score_min in its upper 32 bits.
// Sharded resident set (X.peers != nullptr): the keys are the ones every rank's poses_kernel posted into
// OUR exchange buffer as exchange number match->epoch; every CTA waits for them itself (local polling,
// bounded), so no collect kernel sits between scoring and weighting.
__global__ void __launch_bounds__(PT_THREADS)
weights_kernel(const float *__restrict__ scores, long long N, float beta,
               const unsigned long long *__restrict__ keys, int nkeys,
               unsigned long long *__restrict__ q, unsigned long long *__restrict__ block_sums,
               const XchgArgs X, MatchDev *match)
{
    __shared__ unsigned long long wsum[PT_THREADS / 32];
    __shared__ unsigned long long kmin_s;
    unsigned long long key = ~0ull;
    for (int k = 0; k < nkeys; ++k) key = keys[k] < key ? keys[k] : key;
    if (X.peers) {
        if (threadIdx.x == 0) kmin_s = ~0ull;
        __syncthreads();
        const unsigned int e1 = *reinterpret_cast<volatile unsigned int *>(&match->epoch);
        if (threadIdx.x < X.nranks) {
            unsigned int lo, hi;
            if (xchg_wait_word(X, e1, threadIdx.x, 0, &lo, &match->error, DEV_ERR_PARTICLES) &&
                xchg_wait_word(X, e1, threadIdx.x, 1, &hi, &match->error, DEV_ERR_PARTICLES))
                atomicMin(&kmin_s, ((unsigned long long)hi << 32) | lo);
        }
        __syncthreads();
        key = kmin_s;
    }
    const float smin = __uint_as_float((unsigned int)(key >> 32));
    const long long base = (long long)blockIdx.x * PT_BLOCK + (long long)threadIdx.x * PT_ITEMS;
    unsigned long long local = 0;
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        const long long i = base + j;
        if (i < N) {
            const float d = __fsub_rn(scores[i], smin);
            const float w = exp_det(-__fmul_rn(beta, d));
            const unsigned long long qi = (unsigned long long)__dmul_rn((double)w, 4294967296.0);
            q[i] = qi;
            local += qi;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
    if (lane == 0) wsum[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < PT_THREADS / 32; ++w) t += wsum[w];
        block_sums[blockIdx.x] = t;
    }
}

// number of slots k in [0, N) with T_k < c  (T_k = U + k Wd + floor(k Wm / N) is non-decreasing in k)
__host__ __device__ inline long long slots_below(unsigned long long c, unsigned long long U, unsigned long long Wd,
                                                 unsigned long long Wm, long long N)
{
    long long lo = 0, hi = N;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        const unsigned long long T = U + (unsigned long long)mid * Wd +
                                     ((unsigned long long)mid * Wm) / (unsigned long long)N;
        if (T < c) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// (Wd * u0) >> 32 without 128-bit arithmetic, exact
__host__ __device__ inline unsigned long long u_offset(unsigned long long Wd, unsigned int u0_q32)
{
    return (Wd >> 32) * u0_q32 + (((Wd & 0xffffffffull) * u0_q32) >> 32);
}

// slots_below for a whole CTA: every thread tests one k per round (the predicate T_k < c is monotone in k), so the
// range shrinks by blockDim.x per round -- two rounds for a million slots, one 64-bit division per thread and round,
// where one thread's binary search is ~40 dependent 64-bit divisions (10 us).  All threads call it and get the
// same answer.
__device__ long long slots_below_cta(unsigned long long c, unsigned long long U, unsigned long long Wd,
                                     unsigned long long Wm, long long N)
{
    long long lo = 0, hi = N;                       // the answer lies in [lo, hi]
    while (hi > lo) {
        const long long step = (hi - lo + blockDim.x - 1) / blockDim.x;
        const long long k = lo + (long long)threadIdx.x * step;
        bool p = false;
        if (k < hi) p = U + (unsigned long long)k * Wd + ((unsigned long long)k * Wm) / (unsigned long long)N < c;
        const int cnt = __syncthreads_count(p);     // the true samples are a prefix
        if (cnt == 0) {
            hi = lo;
        } else {
            const long long last_true = lo + (long long)(cnt - 1) * step;
            lo = last_true + 1;
            hi = last_true + step < hi ? last_true + step : hi;
        }
    }
    return lo;
}

// Exclusive scan of the per-CTA sums (one CTA); total -> out[PF_WLOCAL].
// Sharded resident set (X.peers != nullptr): second exchange of the filter step.  This rank's integer
// weight sum goes to every rank as exchange number match->epoch + 1; when every rank's sum (and the
// particle counts that travelled with the first exchange) are here, the CTA derives everything the
// resampler needs ON THE DEVICE -- global W and N, this rank's offset into the global cumulative weight,
// the systematic-resampling slots whose ancestors live here (b200slam_resample_owned_slots' arithmetic)
// and the first slot every rank holds -- so no sum ever visits the host.
__global__ void __launch_bounds__(1024)
block_sums_scan_kernel(unsigned long long *__restrict__ block_sums, int nb,
                       unsigned long long *__restrict__ out, const XchgArgs X, MatchDev *match, long long N_local,
                       unsigned int u0_q32)
{
    unsigned long long *out_total = out;
    __shared__ unsigned long long wtot[32];
    __shared__ unsigned long long carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const unsigned long long v = i < nb ? block_sums[i] : 0;
        unsigned long long inc = warp_incl_scan(v, lane);
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = wtot[lane];
            t = warp_incl_scan(t, lane);
            wtot[lane] = t;
        }
        __syncthreads();
        const unsigned long long before = carry + (warp ? wtot[warp - 1] : 0);
        if (i < nb) block_sums[i] = before + inc - v;     // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) { out_total[PF_WLOCAL] = carry; out_total[PF_WGLOBAL] = carry; }   // one GPU: global == local
    if (!X.peers) return;
    __shared__ unsigned int words_s[4];
    __shared__ unsigned long long Wr[XCHG_MAX_RANKS];
    __shared__ unsigned int Nr[XCHG_MAX_RANKS];
    const unsigned int e1 = *reinterpret_cast<volatile unsigned int *>(&match->epoch), e2 = e1 + 1;
    const unsigned long long Wl = carry;
    __syncthreads();
    xchg_post_words(X, e2, (unsigned int)Wl, (unsigned int)(Wl >> 32), 0u, 0u, words_s);
    if (threadIdx.x < X.nranks) {
        unsigned int lo = 0, hi = 0, n = 0;
        const bool ok = xchg_wait_word(X, e2, threadIdx.x, 0, &lo, &match->error, DEV_ERR_PARTICLES) &&
                        xchg_wait_word(X, e2, threadIdx.x, 1, &hi, &match->error, DEV_ERR_PARTICLES) &&
                        xchg_wait_word(X, e1, threadIdx.x, 2, &n, &match->error, DEV_ERR_PARTICLES);
        Wr[threadIdx.x] = ok ? ((unsigned long long)hi << 32) | lo : 0ull;
        Nr[threadIdx.x] = ok ? n : 0u;
    }
    __syncthreads();
    __shared__ unsigned long long sWg, sOff, sNg;
    if (threadIdx.x == 0) {
        unsigned long long Wg = 0, off = 0, Ng = 0;
        for (int r = 0; r < X.nranks; ++r) {
            if (r < X.rank) off += Wr[r];
            Wg += Wr[r];
            out[PF_SLOT_BASE + r] = Ng;
            Ng += Nr[r];
        }
        out[PF_SLOT_BASE + X.nranks] = Ng;
        sWg = Wg; sOff = off; sNg = Ng;
    }
    __syncthreads();
    const unsigned long long Wg = sWg, off = sOff, Ng = sNg;
    long long kb = 0, ke = 0;
    if (Wg > 0 && Ng > 0) {                          // uniform across the CTA
        const unsigned long long Wd = Wg / Ng, Wm = Wg % Ng, U = u_offset(Wd, u0_q32);
        kb = slots_below_cta(off, U, Wd, Wm, (long long)Ng);
        ke = slots_below_cta(off + Wl, U, Wd, Wm, (long long)Ng);
    }
    if (threadIdx.x == 0) {
        out[PF_WGLOBAL] = Wg; out[PF_RANK_OFFSET] = off; out[PF_NGLOBAL] = Ng;
        out[PF_KBEGIN] = (unsigned long long)kb; out[PF_KCOUNT] = (unsigned long long)(ke - kb);
        match->epoch = e2; match->posted = e2; match->collected = e2;
        (void)N_local;
    }
}

// In-place inclusive prefix C_i (local to this rank) and normalised weights.
// totals: [0] = local W, [1] = global W (== local on one GPU).
__global__ void __launch_bounds__(PT_THREADS)
prefix_kernel(unsigned long long *__restrict__ q, long long N,
              const unsigned long long *__restrict__ block_offsets,
              const unsigned long long *__restrict__ totals, float *__restrict__ weights)
{
    __shared__ unsigned long long wtot[PT_THREADS / 32];
    const long long base = (long long)blockIdx.x * PT_BLOCK + (long long)threadIdx.x * PT_ITEMS;
    const double W = (double)totals[PF_WGLOBAL];
    unsigned long long v[PT_ITEMS];
    unsigned long long tsum = 0;
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        const long long i = base + j;
        v[j] = i < N ? q[i] : 0;
        if (weights && i < N) weights[i] = __double2float_rn(__ddiv_rn((double)v[j], W));
        tsum += v[j];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long inc = warp_incl_scan(tsum, lane);
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    unsigned long long before = block_offsets[blockIdx.x];
    for (int w = 0; w < warp; ++w) before += wtot[w];
    unsigned long long run = before + inc - tsum;
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        const long long i = base + j;
        run += v[j];
        if (i < N) q[i] = run;
    }
}

// Systematic resampling: slot k -> first local particle whose (rank-offset) inclusive
// prefix exceeds T_k.  Slots [k_begin, k_begin + k_count) are the ones whose ancestor
// lives on this rank.
__global__ void __launch_bounds__(256)
resample_kernel(const unsigned long long *__restrict__ C, long long N_local, long long N_global,
                unsigned long long rank_offset, unsigned long long Wd, unsigned long long Wm,
                unsigned long long U, long long k_begin, long long k_count,
                long long index_base, int *__restrict__ ancestors)
{
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k_count) return;
    const unsigned long long k = (unsigned long long)(k_begin + s);
    const unsigned long long T = U + k * Wd + (k * Wm) / (unsigned long long)N_global;
    const unsigned long long t = T - rank_offset;      // T >= rank_offset for owned slots
    long long lo = 0, hi = N_local - 1;                // first i with C[i] > t
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (C[mid] > t) hi = mid; else lo = mid + 1;
    }
    ancestors[s] = (int)(lo + index_base);
}

// Single-GPU variant for the device-resident particle set: the weight sum is read on the device
// (no host round trip), every slot belongs to this rank, and the offspring is gathered right
// away: particle k of `dst` <- particle ancestors[k] of `src` (SoA x | y | ct | st | theta).
__global__ void __launch_bounds__(256)
resample_gather_kernel(const unsigned long long *__restrict__ C, long long N,
                       const unsigned long long *__restrict__ totals, unsigned int u0_q32,
                       int *__restrict__ ancestors, const float *__restrict__ src, float *__restrict__ dst,
                       size_t cap)
{
    __shared__ unsigned long long sWd, sWm, sU;
    if (threadIdx.x == 0) {
        const unsigned long long W = totals[0];
        sWd = W / (unsigned long long)N;
        sWm = W % (unsigned long long)N;
        sU = u_offset(sWd, u0_q32);
    }
    __syncthreads();
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const unsigned long long T = sU + (unsigned long long)k * sWd + ((unsigned long long)k * sWm) / (unsigned long long)N;
    long long lo = 0, hi = N - 1;                      // first i with C[i] > T
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (C[mid] > T) hi = mid; else lo = mid + 1;
    }
    ancestors[k] = (int)lo;
#pragma unroll
    for (int f = 0; f < 5; ++f) dst[f * cap + k] = src[f * cap + lo];
}

// Sharded resident set: resampling + offspring exchange in ONE kernel.  This rank emits the offspring of ITS
// particles: the systematic slots [kb, kb + kc) whose thresholds fall into its stretch of the global
// cumulative weight (found on the device by the scan kernel).  Slot k lives on the rank d with
// slot_base[d] <= k < slot_base[d + 1]; the ancestor's pose (x | y | ct | st | theta) and its GLOBAL index
// are stored straight into rank d's other pose buffer / ancestor array over NVLink (every rank's pose block
// is mapped through CUDA IPC).  Consecutive slots have consecutive destinations and (mostly) equal or
// adjacent ancestors, so both sides coalesce.  A peer barrier behind this kernel completes the step.
// The step's closing barrier is this kernel's tail: every CTA fences its peer stores (system scope) and takes a
// ticket; the last one posts this rank's barrier flag to every peer and waits for theirs (bounded), so when the
// kernel completes every rank's offspring has landed everywhere -- no separate barrier launch.
__global__ void __launch_bounds__(256)
resample_push_kernel(const unsigned long long *__restrict__ C, long long N_local, unsigned long long *__restrict__ S,
                     unsigned int u0_q32, long long index_base, const float *__restrict__ src, size_t cap,
                     float *const *__restrict__ peer_blocks, int dst_parity, int nranks, const XchgArgs X, MatchDev *match)
{
    __shared__ unsigned long long sbase[XCHG_MAX_RANKS + 1];
    __shared__ int last_s;
    for (int r = threadIdx.x; r <= nranks; r += blockDim.x) sbase[r] = S[PF_SLOT_BASE + r];
    __syncthreads();
    const unsigned long long Wg = S[PF_WGLOBAL], Ng = S[PF_NGLOBAL], roff = S[PF_RANK_OFFSET];
    const long long kb = (long long)S[PF_KBEGIN], kc = Ng ? (long long)S[PF_KCOUNT] : 0;
    const unsigned long long Wd = Ng ? Wg / Ng : 0, Wm = Ng ? Wg % Ng : 0, U = u_offset(Wd, u0_q32);
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < kc; s += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k = (unsigned long long)(kb + s);
        const unsigned long long t = U + k * Wd + (k * Wm) / Ng - roff;     // T_k >= rank offset for owned slots
        long long lo = 0, hi = N_local - 1;                                  // first i with C[i] > t
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (C[mid] > t) hi = mid; else lo = mid + 1;
        }
        int d = 0;
        while (d + 1 < nranks && sbase[d + 1] <= k) ++d;
        const size_t j = (size_t)(k - sbase[d]);
        float *blk = peer_blocks[d];
        float *dst = blk + (size_t)dst_parity * 5 * cap;
#pragma unroll
        for (int f = 0; f < 5; ++f) dst[f * cap + j] = src[f * cap + lo];
        reinterpret_cast<int *>(blk + 10 * cap)[j] = (int)(lo + index_base);
    }
    // ---- closing barrier ----------------------------------------------------------------------------------
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t = atomicAdd(&S[PF_TICKET], 1ull);
        last_s = t == (unsigned long long)gridDim.x - 1;
    }
    __syncthreads();
    if (!last_s) return;
    __threadfence_system();
    const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long *>(&match->bar_epoch) + 1;
    if ((int)threadIdx.x < X.nranks) {
        const int r = threadIdx.x;
        *reinterpret_cast<volatile unsigned long long *>(&X.peers[r]->bar[X.rank]) = epoch;
        const volatile unsigned long long *mine = &X.peers[X.rank]->bar[r];
        const unsigned long long budget = *reinterpret_cast<volatile unsigned int *>(&match->error) ? 0ull : X.timeout_ns;
        if (*mine < epoch) {
            const unsigned long long t0 = global_timer_ns();
            unsigned int spins = 0;
            while (*mine < epoch)
                if ((++spins & 255u) == 0 && global_timer_ns() - t0 > budget) {
                    atomicOr(&match->error, DEV_ERR_BARRIER);
                    break;
                }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        match->bar_epoch = epoch;
        S[PF_TICKET] = 0;
    }
}

}  // namespace

int particles_resample_resident(b200slam_ctx *ctx, int64_t N, float beta, uint32_t u0_q32)
{
    if (N <= 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no scored poses");
    const int nb = (int)((N + PT_BLOCK - 1) / PT_BLOCK);
    XchgArgs X = xchg_args(ctx);
    if (!ctx->pf_sharded) X.peers = nullptr;
    weights_kernel<<<nb, PT_THREADS, 0, ctx->stream>>>(ctx->d_scores, N, beta, &ctx->d_match->key, X.peers ? 0 : 1, ctx->d_q,
                                                       ctx->d_block_sums, X, ctx->d_match);
    LAUNCH_CHECK(ctx);
    block_sums_scan_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_block_sums, nb, ctx->d_wsum, X, ctx->d_match, N, u0_q32);
    LAUNCH_CHECK(ctx);
    prefix_kernel<<<nb, PT_THREADS, 0, ctx->stream>>>(ctx->d_q, N, ctx->d_block_sums, ctx->d_wsum, ctx->d_weights);
    LAUNCH_CHECK(ctx);
    if (ctx->pf_sharded) {
        resample_push_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(
            ctx->d_q, N, ctx->d_wsum, u0_q32, ctx->last_index_base, ctx->d_pose_soa, ctx->pose_cap, ctx->d_pf_peers,
            ctx->pose_parity ^ 1, ctx->nranks, X, ctx->d_match);       // its tail is the step's closing barrier
        LAUNCH_CHECK(ctx);
    } else {
        resample_gather_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->d_q, N, ctx->d_wsum, u0_q32, ctx->d_anc_resident, ctx->d_pose_soa, ctx->d_pose_alt, ctx->pose_cap);
        LAUNCH_CHECK(ctx);
    }
    float *t = ctx->d_pose_soa;          // the offspring is the resident set now
    ctx->d_pose_soa = ctx->d_pose_alt;
    ctx->d_pose_alt = t;
    ctx->pose_parity ^= 1;
    return B200SLAM_OK;
}

void particles_unshare_blocks(b200slam_ctx *ctx)
{
    for (int r = 0; r < XCHG_MAX_RANKS; ++r) {
        if (ctx->pf_peer_block[r] && ctx->pf_peer_block[r] != ctx->pf_shared_block) cudaIpcCloseMemHandle(ctx->pf_peer_block[r]);
        ctx->pf_peer_block[r] = nullptr;
    }
    ctx->pf_shared_block = nullptr;
    ctx->pf_shared_cap = 0;
    ctx->pf_sharded = false;
    cudaGetLastError();
}

// Collective.  Every rank publishes the CUDA IPC handle of its pose block; when any rank's block changed since
// the last exchange (first call, or a capacity growth) all ranks re-map all blocks.  The blocks must have the
// same capacity everywhere (the pusher addresses a peer's buffers with its own layout).
int particles_share_blocks(b200slam_ctx *ctx)
{
    if (!ctx->nccl_comm || ctx->nranks < 2 || !ctx->p2p_ready)
        return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "a sharded particle set needs b200slam_comm_init with NVLink peer memory");
    const int n = ctx->nranks;
    constexpr int REC = 12;                       // 64-byte handle | ok | capacity | changed
    unsigned long long rec[REC] = {0}, all[XCHG_MAX_RANKS * REC];
    cudaIpcMemHandle_t h;
    const bool changed = ctx->pf_shared_block != ctx->d_pose_block || ctx->pf_shared_cap != ctx->pose_cap;
    bool ok = ctx->d_pose_block && cudaIpcGetMemHandle(&h, ctx->d_pose_block) == cudaSuccess;
    if (ok) memcpy(rec, &h, sizeof h);
    rec[8] = ok ? 1 : 0;
    rec[9] = ctx->pose_cap;
    rec[10] = changed ? 1 : 0;
    unsigned long long *d_send = nullptr, *d_recv = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_send, sizeof rec));
    CUDA_TRY(ctx, cudaMalloc(&d_recv, sizeof(rec) * n));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_send, rec, sizeof rec, cudaMemcpyHostToDevice, ctx->stream));
    int rc = comm_allgather_u64(ctx, d_send, d_recv, REC);
    if (rc == B200SLAM_OK) {
        CUDA_TRY(ctx, cudaMemcpyAsync(all, d_recv, sizeof(rec) * n, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(d_send); cudaFree(d_recv);
    if (rc) return rc;
    bool any_changed = false;
    for (int r = 0; r < n; ++r) {
        ok = ok && all[r * REC + 8] == 1 && all[r * REC + 9] == rec[9];
        any_changed = any_changed || all[r * REC + 10] == 1;
    }
    if (!ok) {
        cudaGetLastError();
        return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "particle blocks cannot be peer-shared (no CUDA IPC, or the ranks' "
                                                            "capacities differ: give every rank the same number of particles)");
    }
    if (!any_changed) return B200SLAM_OK;
    for (int r = 0; r < XCHG_MAX_RANKS; ++r) {
        if (ctx->pf_peer_block[r] && ctx->pf_peer_block[r] != ctx->pf_shared_block) cudaIpcCloseMemHandle(ctx->pf_peer_block[r]);
        ctx->pf_peer_block[r] = nullptr;
    }
    ctx->pf_shared_block = ctx->d_pose_block;
    ctx->pf_shared_cap = ctx->pose_cap;
    for (int r = 0; r < n; ++r) {
        if (r == ctx->rank) { ctx->pf_peer_block[r] = ctx->d_pose_block; continue; }
        cudaIpcMemHandle_t ph;
        memcpy(&ph, &all[r * REC], sizeof ph);
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            return b200slam_set_error(ctx, B200SLAM_ERR_CUDA, "cudaIpcOpenMemHandle failed for rank %d's particle block", r);
        }
        ctx->pf_peer_block[r] = static_cast<float *>(p);
    }
    if (!ctx->d_pf_peers) CUDA_TRY(ctx, cudaMalloc(&ctx->d_pf_peers, sizeof(float *) * XCHG_MAX_RANKS));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_pf_peers, ctx->pf_peer_block, sizeof(float *) * n, cudaMemcpyHostToDevice));
    return B200SLAM_OK;
}

int particles_weights_resample(b200slam_ctx *ctx, int64_t N, float beta, uint32_t u0_q32,
                               float *weights, uint64_t *wsum, int32_t *ancestors,
                               int64_t *k_begin_out, int64_t *k_count_out)
{
    if (N <= 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no scored poses");
    const int nb = (int)((N + PT_BLOCK - 1) / PT_BLOCK);
    const bool multi = ctx->nccl_comm != nullptr && ctx->nranks > 1;
    const unsigned long long *keys = &ctx->d_match->key;
    int nkeys = 1;
    if (multi) {
        int rc = comm_allgather_u64(ctx, &ctx->d_match->key, ctx->d_keys, 1);
        if (rc) return rc;
        keys = ctx->d_keys;
        nkeys = ctx->nranks;
    }
    XchgArgs noX = xchg_args(ctx);
    noX.peers = nullptr;
    weights_kernel<<<nb, PT_THREADS, 0, ctx->stream>>>(ctx->d_scores, N, beta, keys, nkeys, ctx->d_q,
                                                       ctx->d_block_sums, noX, ctx->d_match);
    LAUNCH_CHECK(ctx);
    block_sums_scan_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_block_sums, nb, ctx->d_wsum, noX, ctx->d_match, N, u0_q32);
    LAUNCH_CHECK(ctx);

    // totals: local W in d_wsum[0]; gather every rank's W and N for the global picture
    unsigned long long rank_offset = 0, Wglobal = 0;
    long long Nglobal = N, index_base = ctx->last_index_base;
    if (multi) {
        // d_wsum[1] = local N, so one 16-byte all-gather carries both
        unsigned long long nloc = (unsigned long long)N;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_wsum + 1, &nloc, 8, cudaMemcpyHostToDevice, ctx->stream));
        unsigned long long *d_all = ctx->d_keys + 128;              // [nranks][2]
        int rc = comm_allgather_u64(ctx, ctx->d_wsum, d_all, 2);
        if (rc) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_wsum, d_all, 16 * ctx->nranks, cudaMemcpyDeviceToHost,
                                      ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        Nglobal = 0;
        for (int r = 0; r < ctx->nranks; ++r) {
            if (r < ctx->rank) rank_offset += ctx->h_wsum[2 * r];
            Wglobal += ctx->h_wsum[2 * r];
            Nglobal += (long long)ctx->h_wsum[2 * r + 1];
        }
    } else {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_wsum, ctx->d_wsum, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        Wglobal = ctx->h_wsum[0];
    }
    const unsigned long long Wlocal = multi ? ctx->h_wsum[2 * ctx->rank] : Wglobal;
    if (Wglobal == 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "all particle weights are zero");
    // d_wsum[1] <- global W for the normalisation
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_wsum + 1, &Wglobal, 8, cudaMemcpyHostToDevice, ctx->stream));
    prefix_kernel<<<nb, PT_THREADS, 0, ctx->stream>>>(ctx->d_q, N, ctx->d_block_sums, ctx->d_wsum,
                                                      weights ? ctx->d_weights : nullptr);
    LAUNCH_CHECK(ctx);

    const unsigned long long Wd = Wglobal / (unsigned long long)Nglobal;
    const unsigned long long Wm = Wglobal % (unsigned long long)Nglobal;
    const unsigned long long U = (unsigned long long)(((unsigned __int128)Wd * u0_q32) >> 32);
    int64_t kb64 = 0, kc64 = 0;
    b200slam_resample_owned_slots(Wglobal, Nglobal, u0_q32, rank_offset, Wlocal, &kb64, &kc64);
    const long long k_begin = kb64, k_count = kc64;
    if (ancestors && k_count > 0) {
        if ((size_t)k_count > ctx->anc_cap) {
            if (ctx->d_ancestors) cudaFree(ctx->d_ancestors);
            ctx->d_ancestors = nullptr;
            ctx->anc_cap = 0;
            CUDA_TRY(ctx, cudaMalloc(&ctx->d_ancestors, sizeof(int32_t) * (size_t)k_count));
            ctx->anc_cap = (size_t)k_count;
        }
        const unsigned grid = (unsigned)((k_count + 255) / 256);
        resample_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_q, N, Nglobal, rank_offset, Wd, Wm, U,
                                                       k_begin, k_count, index_base, ctx->d_ancestors);
        LAUNCH_CHECK(ctx);
        CUDA_TRY(ctx, cudaMemcpyAsync(ancestors, ctx->d_ancestors, sizeof(int32_t) * (size_t)k_count,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (weights)
        CUDA_TRY(ctx, cudaMemcpyAsync(weights, ctx->d_weights, sizeof(float) * (size_t)N,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (wsum) *wsum = Wglobal;
    if (k_begin_out) *k_begin_out = k_begin;
    if (k_count_out) *k_count_out = k_count;
    return B200SLAM_OK;
}

// Pure host arithmetic (no GPU): the slots k in [0, N) whose threshold T_k falls inside this
// rank's stretch [rank_offset, rank_offset + w_local) of the global cumulative weight.
extern "C" void b200slam_resample_owned_slots(uint64_t w_global, int64_t n_global, uint32_t u0_q32,
                                              uint64_t rank_offset, uint64_t w_local, int64_t *k_begin,
                                              int64_t *k_count)
{
    long long kb = 0, ke = 0;
    if (n_global > 0 && w_global > 0) {
        const unsigned long long Wd = w_global / (unsigned long long)n_global;
        const unsigned long long Wm = w_global % (unsigned long long)n_global;
        const unsigned long long U = (unsigned long long)(((unsigned __int128)Wd * u0_q32) >> 32);
        kb = slots_below(rank_offset, U, Wd, Wm, n_global);
        ke = slots_below(rank_offset + w_local, U, Wd, Wm, n_global);
    }
    if (k_begin) *k_begin = kb;
    if (k_count) *k_count = ke - kb;
}
