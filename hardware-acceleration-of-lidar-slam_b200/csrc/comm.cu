// comm.cu -- multi-GPU exchange: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The path shards by candidate rows / particles with the map replicated (SURVEY.md
// section 8e), so the only data-path collectives are all-gathers of a few bytes per rank:
// the packed (score, index) arg-min key and the integer weight sums.  NCCL is resolved
// with dlopen at b200slam_comm_init time, so a single-GPU user of libb200slam.so needs no
// NCCL at all, and a process that already loaded a libnccl.so.2 (e.g. torch's bundled one)
// shares it instead of loading a second copy.
#include <dlfcn.h>
#include <stdlib.h>
#include <nccl.h>

#include "common.cuh"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    char why[256] = "";
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        snprintf(api.why, sizeof api.why, "dlopen(libnccl.so.2) failed: %s", dlerror());
        return &api;
    }
#define SYM(field, name)                                                             \
    do {                                                                             \
        api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));  \
        if (!api.field) {                                                            \
            snprintf(api.why, sizeof api.why, "dlsym(%s) failed", name);             \
            api.handle = nullptr;                                                    \
            return &api;                                                             \
        }                                                                            \
    } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather");
    SYM(Broadcast, "ncclBroadcast");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return &api;
}

}  // namespace

int comm_allgather_u64(b200slam_ctx *ctx, const unsigned long long *d_send, unsigned long long *d_recv,
                       int count_per_rank)
{
    NcclApi *api = nccl_api();
    if (!api->handle || !ctx->nccl_comm)
        return b200slam_set_error(ctx, B200SLAM_ERR_NCCL, "no communicator (%s)", api->why);
    ncclResult_t r = api->AllGather(d_send, d_recv, (size_t)count_per_rank, ncclUint64,
                                    (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != ncclSuccess)
        return b200slam_set_error(ctx, B200SLAM_ERR_NCCL, "ncclAllGather -> %s", api->GetErrorString(r));
    return B200SLAM_OK;
}

int comm_gather_row_blocks(b200slam_ctx *ctx, float *d_field, int pitch, int rows)
{
    NcclApi *api = nccl_api();
    if (!api->handle || !ctx->nccl_comm)
        return b200slam_set_error(ctx, B200SLAM_ERR_NCCL, "no communicator (%s)", api->why);
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const int n = ctx->nranks;
    ncclResult_t r = ncclSuccess;
    if (rows % n == 0) {
        const size_t blk = (size_t)(rows / n) * pitch;
        r = api->AllGather(d_field + blk * ctx->rank, d_field, blk, ncclFloat, comm, ctx->stream);
    } else {
        r = api->GroupStart();
        for (int k = 0; k < n && r == ncclSuccess; ++k) {
            int64_t b, e;
            b200slam_shard_range(rows, n, k, &b, &e);
            if (e > b) {
                float *blk = d_field + (size_t)b * pitch;
                r = api->Broadcast(blk, blk, (size_t)(e - b) * pitch, ncclFloat, k, comm, ctx->stream);
            }
        }
        ncclResult_t r2 = api->GroupEnd();
        if (r == ncclSuccess) r = r2;
    }
    if (r != ncclSuccess)
        return b200slam_set_error(ctx, B200SLAM_ERR_NCCL, "field row-block exchange -> %s", api->GetErrorString(r));
    return B200SLAM_OK;
}

namespace {

// Device-side barrier over the ranks: everybody stores the barrier's number into its slot of
// every peer's buffer and waits until all the slots of its own buffer have reached it.  The
// system-scope fences order the peer-memory stores of the kernels queued before the barrier
// (the row-sharded EDT's remote rows) against the flag, and the loads of the kernels behind it.
// The barrier's number lives on the DEVICE (MatchDev::bar_epoch, advanced by this kernel alone, in stream
// order), not in a launch argument: a barrier captured into a CUDA graph must count on every replay.
__global__ void __launch_bounds__(64) peer_barrier_kernel(const XchgArgs X, MatchDev *match)
{
    __shared__ unsigned long long epoch_s;
    if (threadIdx.x == 0) epoch_s = *reinterpret_cast<volatile unsigned long long *>(&match->bar_epoch) + 1;
    __syncthreads();
    const unsigned long long epoch = epoch_s;
    unsigned int *error = &match->error;
    const int r = threadIdx.x;
    __threadfence_system();
    if (r < X.nranks) {
        *reinterpret_cast<volatile unsigned long long *>(&X.peers[r]->bar[X.rank]) = epoch;
        const volatile unsigned long long *mine = &X.peers[X.rank]->bar[r];
        // bounded: a peer that never arrives sets a sticky error bit instead of hanging the GPU
        const unsigned long long budget = *reinterpret_cast<volatile unsigned int *>(error) ? 0ull : X.timeout_ns;
        if (*mine < epoch) {
            const unsigned long long t0 = global_timer_ns();
            unsigned int spins = 0;
            while (*mine < epoch)
                if ((++spins & 255u) == 0 && global_timer_ns() - t0 > budget) {
                    atomicOr(error, DEV_ERR_BARRIER);
                    break;
                }
        }
    }
    __threadfence_system();
    if (threadIdx.x == 0) match->bar_epoch = epoch;
}

}  // namespace

int comm_peer_barrier(b200slam_ctx *ctx)
{
    if (!ctx->p2p_ready) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "peer memory exchange is not set up");
    peer_barrier_kernel<<<1, 64, 0, ctx->stream>>>(xchg_args(ctx), ctx->d_match);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}

void comm_unshare_map(b200slam_ctx *ctx, b200slam_map *map)
{
    for (int r = 0; r < map->shared_nranks; ++r) {
        if (map->peer_alloc[r] && map->peer_alloc[r] != map->d_field_alloc) cudaIpcCloseMemHandle(map->peer_alloc[r]);
        map->peer_alloc[r] = nullptr;
    }
    map->shared_nranks = 0;
    (void)ctx;
}

namespace {

void teardown_peer_exchange(b200slam_ctx *ctx)
{
    for (int r = 0; r < XCHG_MAX_RANKS; ++r) {
        if (ctx->peer_ptrs[r] && r != ctx->rank) cudaIpcCloseMemHandle(ctx->peer_ptrs[r]);
        ctx->peer_ptrs[r] = nullptr;
    }
    cudaFree(ctx->d_xchg); ctx->d_xchg = nullptr;
    cudaFree(ctx->d_peers); ctx->d_peers = nullptr;
    ctx->p2p_ready = false;
    ctx->posted_uncollected = 0;
}

// Allocates this rank's exchange buffer, trades CUDA IPC handles with the other ranks (one
// NCCL all-gather at init time) and maps every peer's buffer.  All ranks must agree on the
// outcome, so the per-rank success flags are all-gathered too.
void setup_peer_exchange(b200slam_ctx *ctx)
{
    const int n = ctx->nranks;
    if (n > XCHG_MAX_RANKS) return;
    ctx->posted_uncollected = 0;
    bool ok = cudaMalloc(&ctx->d_xchg, sizeof(XchgBuf)) == cudaSuccess &&
              cudaMemset(ctx->d_xchg, 0, sizeof(XchgBuf)) == cudaSuccess &&
              cudaMalloc(&ctx->d_peers, sizeof(XchgBuf *) * n) == cudaSuccess &&
              cudaMemset(&ctx->d_match->epoch, 0, sizeof(unsigned int)) == cudaSuccess &&
              cudaMemset(&ctx->d_match->collected, 0, sizeof(unsigned int)) == cudaSuccess &&
              cudaMemset(&ctx->d_match->posted, 0, sizeof(unsigned int)) == cudaSuccess &&
              cudaMemset(&ctx->d_match->error, 0, sizeof(unsigned int)) == cudaSuccess &&
              cudaMemset(&ctx->d_match->bar_epoch, 0, sizeof(unsigned long long)) == cudaSuccess;   // fresh XchgBuf: zeroed barrier slots
    // handle record: 64-byte IPC handle + 8-byte ok flag, padded to 80 bytes (10 x u64)
    constexpr int REC = 10;
    unsigned long long rec[REC] = {0};
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "IPC handle size");
    if (ok) ok = cudaIpcGetMemHandle(&h, ctx->d_xchg) == cudaSuccess;
    if (ok) memcpy(rec, &h, sizeof h);
    rec[8] = ok ? 1 : 0;
    unsigned long long *d_send = nullptr, *d_recv = nullptr;
    unsigned long long all[XCHG_MAX_RANKS * REC];
    bool gathered = cudaMalloc(&d_send, sizeof rec) == cudaSuccess &&
                    cudaMalloc(&d_recv, sizeof(rec) * n) == cudaSuccess &&
                    cudaMemcpyAsync(d_send, rec, sizeof rec, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
                    comm_allgather_u64(ctx, d_send, d_recv, REC) == B200SLAM_OK &&
                    cudaMemcpyAsync(all, d_recv, sizeof(rec) * n, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
                    cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    cudaFree(d_send); cudaFree(d_recv);
    if (gathered)
        for (int r = 0; r < n; ++r) ok = ok && all[r * REC + 8] == 1;
    else
        ok = false;
    if (ok) {
        for (int r = 0; r < n && ok; ++r) {
            if (r == ctx->rank) { ctx->peer_ptrs[r] = ctx->d_xchg; continue; }
            cudaIpcMemHandle_t ph;
            memcpy(&ph, &all[r * REC], sizeof ph);
            void *p = nullptr;
            ok = cudaIpcOpenMemHandle(&p, ph, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            ctx->peer_ptrs[r] = static_cast<XchgBuf *>(p);
        }
        if (ok) ok = cudaMemcpy(ctx->d_peers, ctx->peer_ptrs, sizeof(XchgBuf *) * n, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    cudaGetLastError();               // failures here only mean "use the NCCL path"
    // second agreement round: mapping may have failed on one rank only
    unsigned long long mine = ok ? 1 : 0, *d_f = nullptr, *d_all = nullptr, flags[XCHG_MAX_RANKS] = {0};
    bool agree = cudaMalloc(&d_f, 8) == cudaSuccess && cudaMalloc(&d_all, 8 * n) == cudaSuccess &&
                 cudaMemcpyAsync(d_f, &mine, 8, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
                 comm_allgather_u64(ctx, d_f, d_all, 1) == B200SLAM_OK &&
                 cudaMemcpyAsync(flags, d_all, 8 * n, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
                 cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    cudaFree(d_f); cudaFree(d_all);
    for (int r = 0; r < n && agree; ++r) agree = flags[r] == 1;
    if (agree) ctx->p2p_ready = true;
    else teardown_peer_exchange(ctx);
    cudaGetLastError();
}

}  // namespace

extern "C" {

int b200slam_comm_unique_id(void *id_out)
{
    static_assert(sizeof(ncclUniqueId) == B200SLAM_UNIQUE_ID_BYTES, "unique id size");
    if (!id_out) return B200SLAM_ERR_ARG;
    NcclApi *api = nccl_api();
    if (!api->handle) return b200slam_set_error(nullptr, B200SLAM_ERR_NCCL, "%s", api->why);
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess)
        return b200slam_set_error(nullptr, B200SLAM_ERR_NCCL, "ncclGetUniqueId -> %s", api->GetErrorString(r));
    memcpy(id_out, &id, sizeof id);
    return B200SLAM_OK;
}

int b200slam_comm_init(b200slam_ctx *ctx, int nranks, int rank, const void *id_in)
{
    if (!ctx || !id_in || nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) return B200SLAM_ERR_ARG;
    NcclApi *api = nccl_api();
    if (!api->handle) return b200slam_set_error(ctx, B200SLAM_ERR_NCCL, "%s", api->why);
    if (ctx->nccl_comm) b200slam_comm_destroy(ctx);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_in, sizeof id);
    ncclComm_t comm = nullptr;
    ncclResult_t r = api->CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess)
        return b200slam_set_error(ctx, B200SLAM_ERR_NCCL, "ncclCommInitRank -> %s", api->GetErrorString(r));
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    if (const char *e = getenv("B200SLAM_SPIN_TIMEOUT_MS"))
        if (atof(e) > 0) ctx->spin_timeout_ns = (unsigned long long)(atof(e) * 1e6);
    // NVLink peer exchange: best effort.  When CUDA IPC / P2P is unavailable the match results
    // are all-gathered with NCCL instead (same answers, ~20 us more latency per match).
    if (nranks > 1 && !getenv("B200SLAM_NO_P2P")) setup_peer_exchange(ctx);
    return B200SLAM_OK;
}

/* Collective, asynchronous: everything this rank queues behind it runs only once every rank has reached its
 * own call (one 64-thread kernel exchanging flags through NVLink peer memory; a 16-byte ncclAllGather where
 * CUDA IPC is unavailable).  No-op without a communicator. */
int b200slam_comm_barrier_async(b200slam_ctx *ctx)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (!ctx->nccl_comm || ctx->nranks < 2) return B200SLAM_OK;
    if (ctx->p2p_ready) return comm_peer_barrier(ctx);
    return comm_allgather_u64(ctx, ctx->d_wsum, ctx->d_keys + 128, 1);
}

/* Collective: every rank calls it with its own copy of the (identically sized) map. */
int b200slam_map_share(b200slam_ctx *ctx, b200slam_map *map)
{
    if (!ctx || !map) return B200SLAM_ERR_ARG;
    if (!ctx->nccl_comm || ctx->nranks < 2) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_comm_init first");
    if (ctx->nranks > 8) return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "peer-shared maps support up to 8 ranks");
    if (map->shared_nranks) return B200SLAM_OK;
    const int n = ctx->nranks;
    constexpr int REC = 10;                       // 64-byte handle + ok flag + size check
    unsigned long long rec[REC] = {0}, all[8 * REC] = {0};
    cudaIpcMemHandle_t h;
    bool ok = ctx->p2p_ready && cudaIpcGetMemHandle(&h, map->d_field_alloc) == cudaSuccess;
    if (ok) memcpy(rec, &h, sizeof h);
    rec[8] = ok ? 1 : 0;
    rec[9] = ((unsigned long long)map->cap_rows << 32) | (unsigned int)map->field_pitch;
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
    for (int r = 0; r < n; ++r) ok = ok && all[r * REC + 8] == 1 && all[r * REC + 9] == rec[9];
    if (!ok) {
        cudaGetLastError();
        return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "maps cannot be peer-shared (no CUDA IPC / P2P, or sizes differ)");
    }
    map->shared_nranks = n;
    for (int r = 0; r < n; ++r) {
        if (r == ctx->rank) { map->peer_alloc[r] = map->d_field_alloc; continue; }
        cudaIpcMemHandle_t ph;
        memcpy(&ph, &all[r * REC], sizeof ph);
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            comm_unshare_map(ctx, map);
            return b200slam_set_error(ctx, B200SLAM_ERR_CUDA, "cudaIpcOpenMemHandle failed for rank %d's field", r);
        }
        map->peer_alloc[r] = static_cast<float *>(p);
    }
    return B200SLAM_OK;
}

int b200slam_map_edt_rows(b200slam_ctx *ctx, b200slam_map *map, float max_dist, int row_begin, int row_end)
{
    if (!ctx || !map) return B200SLAM_ERR_ARG;
    return edt_launch_rows(ctx, map->d_occ, map->occ_pitch, map->d_field, map->field_pitch, map->rows, map->cols,
                           max_dist, row_begin, row_end, nullptr, 0);
}

int b200slam_map_edt_sharded(b200slam_ctx *ctx, b200slam_map *map, float max_dist, int mode)
{
    if (!ctx || !map) return B200SLAM_ERR_ARG;
    if (!ctx->nccl_comm || ctx->nranks < 2) return b200slam_map_edt(ctx, map, max_dist);
    int64_t rb, re;
    b200slam_shard_range(map->rows, ctx->nranks, ctx->rank, &rb, &re);
    if (mode == B200SLAM_EDT_GATHER_NCCL) {
        int rc = edt_launch_rows(ctx, map->d_occ, map->occ_pitch, map->d_field, map->field_pitch, map->rows,
                                 map->cols, max_dist, (int)rb, (int)re, nullptr, 0);
        if (rc) return rc;
        return comm_gather_row_blocks(ctx, map->d_field, map->field_pitch, map->rows);
    }
    if (mode != B200SLAM_EDT_GATHER_P2P) return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "unknown gather mode %d", mode);
    if (map->shared_nranks != ctx->nranks)
        return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_map_share must be called first");
    float *peers[7];
    int np = 0;
    for (int r = 0; r < ctx->nranks; ++r)
        if (r != ctx->rank) peers[np++] = map->peer_alloc[r] + field_pad_floats(map->field_pitch);
    // nobody may still be reading the old field of a peer when its new rows arrive
    int rc = comm_peer_barrier(ctx);
    if (rc) return rc;
    rc = edt_launch_rows(ctx, map->d_occ, map->occ_pitch, map->d_field, map->field_pitch, map->rows, map->cols,
                         max_dist, (int)rb, (int)re, peers, np);
    if (rc) return rc;
    return comm_peer_barrier(ctx);           // ... and every rank's rows have landed everywhere
}

int b200slam_comm_destroy(b200slam_ctx *ctx)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (ctx->nccl_comm) {
        NcclApi *api = nccl_api();
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        particles_unshare_blocks(ctx);
        teardown_peer_exchange(ctx);
        if (api->handle) api->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->nranks = 1;
    ctx->rank = 0;
    return B200SLAM_OK;
}

}  // extern "C"
