// common.cuh -- internals shared by the translation units of libb200slam.so.
// Nothing here is visible through the C ABI (include/b200slam.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200slam.h"

// Rows of zeros in front of every distance field.  field[-1] is a zero the scoring kernels read for
// out-of-bounds beams (adding +0.0f leaves a score unchanged); the row-reuse matcher reads K <= 9 consecutive
// rows of one column per beam and sends an invalid beam to index -(8 * pitch + 1), from where all of them
// land in this region -- no select on the row stride, no branch.
#define B200SLAM_FIELD_PAD_ROWS 9
inline size_t field_pad_floats(int pitch) { return (size_t)B200SLAM_FIELD_PAD_ROWS * (size_t)pitch; }
#define LAT_SLOTS 4

struct b200slam_map {
    int rows = 0, cols = 0;          // current size (<= capacity; changed by b200slam_map_rasterise)
    int cap_rows = 0, cap_cols = 0;  // allocated size
    int occ_pitch = 0;       // elements
    int field_pitch = 0;     // elements
    int32_t *d_occ = nullptr;
    float *d_field_alloc = nullptr;   // allocation start (pad in front)
    float *d_field = nullptr;         // [0][0]
    float pixel_size = 1.0f, top_left_x = 0.0f, top_left_y = 0.0f;
    bool has_geometry = false;
    // Rasterisation keeps the occupancy "zero everywhere except the cells the last rasterisation set" and
    // remembers those cells, so the next one erases O(points) cells instead of clearing O(cells) bytes:
    // raster_cells_n < 0: the allocation's contents are unknown (fresh, uploaded by the host) -> clear it all once;
    // occ_exposed: the raw pointer was handed out (b200slam_map_device_ptrs) -> always clear the region in use.
    // Byte shadow of the occupancy (1 = occupied), same pitch, kept by every writer of d_occ inside the library
    // (upload: a pack kernel; rasterisation: both stores) for maps of >= 2^20 cells: the transform is HBM bound
    // there and reads 1 byte per cell instead of 4.  occ8_valid is false while d_occ may have been written behind
    // the library's back (occ_exposed) -- the transform then reads the int32 grid.
    uint8_t *d_occ8 = nullptr;
    bool occ8_valid = false;
    int occ8_rows = 0, occ8_cols = 0;     // region the shadow covers (an upload packs the grid in use; a rasterisation: the capacity)
    int32_t *d_raster_cells = nullptr;
    size_t raster_cells_cap = 0;
    int raster_cells_n = -1;
    bool occ_exposed = false;
    // b200slam_map_share: every rank's field allocation mapped here through CUDA IPC
    // (peer_alloc[own rank] == d_field_alloc); nullptr until shared
    float *peer_alloc[64] = {};
    int shared_nranks = 0;
};

// Device-side state of a match.  work_key / tickets are the in-flight arg-min cell and the
// finished-CTA counter; the last CTA of a scoring kernel publishes {key, best_hits,
// last_hits} (what b200slam_match_fetch and the all-gather read) and resets the first two,
// so between launches work_key == ~0 and tickets == 0.
struct MatchDev {
    unsigned long long work_key;
    unsigned int tickets;
    unsigned int epoch;         // peer exchanges this context has POSTED
    unsigned long long key;     // (score bits << 32) | global linear index; ~0 = nothing scored
    int best_hits;
    int last_hits;
    unsigned int collected;     // peer exchanges this context has COLLECTED (merged)
    unsigned int error;         // sticky DEV_ERR_* bits: a bounded device-side wait gave up (b200slam_sync /
                                // b200slam_match_fetch then return B200SLAM_ERR_STATE)
    unsigned long long gkey;    // GLOBAL result of the last collected exchange (all ranks merged)
    int gbest_hits;
    int glast_hits;
    // FastMatch-sized lattices (<= MATCH_SMALL candidates): per-candidate hit counts, and how many
    // leading entries of the hit-value buffer this match (re)wrote.  The reference lets EVERY candidate
    // overwrite bestHits[] from index 0 (main.c:515), so behind the last candidate's hits the buffer
    // holds those of the most recent candidate that had more; the kernel's tail reproduces that.
    int written_hits;
    unsigned int posted;        // peer exchanges whose result has actually been stored into the peers' buffers
    int cand_hits[64];
    // deferred posts (allreduce = 3): a burst of matches records each result here; the collect
    // kernel at the end of the burst sends them all to the peers at once, so no scoring kernel has
    // NVLink stores in flight when it completes (measured: ~8 us per kernel on a 2-GPU box)
    struct Outbox { unsigned long long key; int best_hits, last_hits; } outbox[64];
    unsigned long long bar_epoch;   // device-side peer barriers this context has passed (same on every rank)
    unsigned long long seed_key;    // b200slam_fastmatch_pair_async: the FIRST match's key, saved by the second (seeded) one
};
static_assert(sizeof(MatchDev::outbox) / sizeof(MatchDev::Outbox) == 64, "outbox ring == XCHG_EPOCHS");
// Result block in MAPPED pinned host memory: the tail of the second kernel of a b200slam_fastmatch_pair_async
// writes it straight over PCIe (seq last, behind a system-scope fence), so the host neither queues a copy nor
// calls into the driver to learn the result -- it watches seq.
struct MatchHost {
    unsigned long long key, seed_key;
    int best_hits, last_hits, written_hits, scan_n, mp_n;
    unsigned int error;
    unsigned long long seq;
    long long trace[16];        // SM cycle counter at the fused kernel's phase boundaries (B200SLAM_FM_TRACE diagnostics)
};
// Device-resident per-scan loop (b200slam_scan_chain_*): the poses the motion model and the mini-update test of the
// reference's main() need (main.c:875-898, 928-940) live here, so a scan's kernel can be queued before the scan in
// front of it has finished.  A kernel of scan k runs only while next_scan == k and stop == 0; its tail commits the
// new pose, advances next_scan and raises stop when the mini-update test fires (the kernels queued behind it then
// return at once and leave the scan, the match state and bestHits[] of scan k in place for the host's map update).
struct ChainDev {
    float pose[3], prev[3], map_pose[3];
    int have_prev, next_scan, stop;
};
// One slot per scan in a ring of mapped host memory; seq = scan index + 1, written last behind a system fence.
struct ChainSlot {
    float pose_a[3], pose_b[3];
    int scan_n, best_hits, mp_n, stopped;
    unsigned int error, pad;
    unsigned long long seq;
};
constexpr int CHAIN_RING = 64;
constexpr int MATCH_SMALL = 64;
// Every device-side wait is bounded (%globaltimer against b200slam_ctx::spin_timeout_ns, or an iteration
// cap for the TMA barrier): a dead or mis-ordered peer costs a timeout and an error code, never a hung GPU.
constexpr unsigned int DEV_ERR_EXCHANGE = 1u;   // exchange_collect: a peer's post never arrived
constexpr unsigned int DEV_ERR_BARRIER = 2u;    // peer_barrier_kernel: a peer never reached the barrier
constexpr unsigned int DEV_ERR_TMA = 4u;        // edt_tma_kernel: a TMA load never completed
constexpr unsigned int DEV_ERR_PARTICLES = 8u;  // sharded particle filter: a peer's sums / offspring never arrived

// Peer-memory exchange of per-rank match results (multi-GPU): every context owns one
// XchgBuf; the last CTA of a scoring kernel stores its {key, best_hits, last_hits} into
// EVERY peer's buffer over NVLink (P2P mapped through CUDA IPC), waits until all ranks'
// words of that epoch have landed in its own buffer, and merges.  The four 32-bit payload
// words travel as 8-byte stores {data, epoch} (each store is atomic and carries its own
// validity flag, like NCCL's LL protocol), so no fence and no separate flag are needed.
// 64 epochs of slots.  With a merge in every kernel tail a rank runs at most two posts ahead of
// the slowest reader; with post-only bursts (allreduce = 3) a rank posts a burst only after its
// blocking collect of the previous burst, which needed every peer's posts of that burst, which
// the peers issued after collecting the burst before -- so slots at distance >= two bursts
// (2 x XCHG_MAX_POSTED <= 64) have been consumed everywhere.
constexpr int XCHG_MAX_RANKS = 64;
constexpr int XCHG_EPOCHS = 64;
constexpr int XCHG_MAX_POSTED = 31;   // posts a rank may issue between two blocking collects (two bursts fit the ring)
struct XchgSlot {
    unsigned long long w[4];          // {key lo, key hi, best_hits, last_hits}, each | (epoch << 32)
};
struct XchgBuf {
    XchgSlot slot[XCHG_EPOCHS][XCHG_MAX_RANKS];
    unsigned long long bar[XCHG_MAX_RANKS];   // peer barrier: bar[r] = last barrier epoch rank r has reached
};
struct XchgArgs {                     // kernel-side view; peers == nullptr: no exchange
    XchgBuf *const *peers;            // [nranks] device pointers (own buffer at [rank])
    int nranks, rank;
    unsigned long long timeout_ns;    // bound of every wait on a peer
};

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct b200slam_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    char err[512] = {0};
    uint64_t launches = 0;
    bool use_pdl = true;         // programmatic dependent launch between consecutive scan-matching kernels
    bool prev_launch_was_lattice = false;   // the last kernel queued on the stream was a scan-matching kernel
    bool prev_launch_was_edt = false;       // ... was a distance transform (the next transform may start under its tail)
    const float *prev_lattice_field = nullptr;   // field the last scan-matching kernel reads (valid while prev_launch_was_lattice)
    int match_mode = B200SLAM_MATCH_LATENCY; // tile-shape policy of the lattice kernel (b200slam_set_match_mode)

    // scan (sensor frame), device resident
    float *d_scan_x = nullptr, *d_scan_y = nullptr;   // one allocation: x[scan_cap] | y[scan_cap]
    float *h_scan = nullptr;                          // pinned staging, same layout
    float *h_hit_values = nullptr;                    // pinned landing zone for the last candidate's hits
    cudaEvent_t scan_event = nullptr;                 // staging buffer free again
    int nbeams = 0, scan_cap = 0;

    // lattice axis tables: host-pinned staging + device copy
    float *h_lat = nullptr, *d_lat = nullptr;
    size_t lat_cap = 0;          // floats per slot
    cudaEvent_t lat_event[LAT_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    int lat_next = 0, lat_cur = 0;
    uint64_t graph_launch_mark = 0;

    // match result
    MatchDev *d_match = nullptr;
    MatchDev *h_match = nullptr;            // pinned
    unsigned long long *d_keys = nullptr;   // all-gather landing zones: [0,128) {key, hits} per rank,
                                            // [128,256) {weight sum, count} per rank
    unsigned long long *h_keys = nullptr;   // pinned [128]
    float *d_hit_values = nullptr;          // [2][scan_cap]: best / last candidate
    // bookkeeping of the last queued lattice match (host side)
    struct {
        bool valid = false;
        int n[3] = {0, 0, 0};
        float pose0[3] = {0, 0, 0};
        float step[3] = {0, 0, 0};
        bool is_poses = false;
        bool gathered = false;   // per-rank results were all-gathered into d_keys
        bool exchanged = false;  // per-rank results were exchanged through peer memory: read MatchDev::g*
    } last;

    // optional full score table
    float *d_scores = nullptr;
    size_t scores_cap = 0;       // floats

    // pose-list scoring (particles)
    // one allocation d_pose_block = [2][5 * pose_cap] floats | [pose_cap] int32 ancestors of the resident set; the
    // two pose buffers swap at every resampling (pose_parity = which half d_pose_soa is).  One block, so that
    // the sharded particle filter can map every peer's block through ONE CUDA IPC handle (pf_peer_block).
    float *d_pose_block = nullptr;
    float *d_pose_soa = nullptr;  // x | y | ct | st | theta, each [pose_cap]
    float *d_pose_alt = nullptr;  // same layout: target of the resampling gather (buffers swap)
    int32_t *d_anc_resident = nullptr;
    int pose_parity = 0;
    size_t pose_cap = 0;
    // sharded resident particle set (b200slam_particles_shard): this rank's slice of N_global particles
    bool pf_sharded = false;
    long long pf_nglobal = 0;
    float *pf_peer_block[XCHG_MAX_RANKS] = {};   // every rank's d_pose_block (own at [rank]), IPC mapped
    float **d_pf_peers = nullptr;                // device copy [nranks]
    float *pf_shared_block = nullptr;            // the d_pose_block the mappings were exchanged for
    size_t pf_shared_cap = 0;
    int32_t *d_hits = nullptr;
    float *h_pose_stage = nullptr;   // pinned staging [4][pose_cap]
    int64_t last_P = 0;
    int64_t last_index_base = 0;
    const float *last_poses_host = nullptr;

    // particle weights / resampling scratch
    unsigned long long *d_q = nullptr;        // fixed-point weights / inclusive prefix [pose_cap]
    unsigned long long *d_block_sums = nullptr;
    float *d_weights = nullptr;
    int32_t *d_ancestors = nullptr;
    size_t anc_cap = 0;
    unsigned long long *d_wsum = nullptr;     // [PF_SCALARS] scratch scalars (layout: PfScalar)
    unsigned long long *h_wsum = nullptr;     // pinned [4]

    // map points staged for rasterisation: x | y
    float *d_points = nullptr, *h_points = nullptr;   // device / pinned
    size_t points_cap = 0;

    // scan front end (frontend.cu): lidar tables, raw ranges, world-frame scan, map points.
    // The local map extracted from the map points lives in d_points (what rasterisation reads).
    float *d_lidar = nullptr;            // cos | sin of the beam angles, each [lidar_n]
    int lidar_n = 0;
    float lidar_range_min = 0.0f;
    float *d_ranges = nullptr, *h_ranges = nullptr;   // device / pinned, [lidar_n]
    float *d_scan_t = nullptr;           // tx | ty, each [scan_t_cap]
    int scan_t_cap = 0;
    bool scan_t_valid = false;
    float *d_mp = nullptr;               // map points x | y, each [mp_cap]
    int mp_cap = 0, mp_size = 0;
    int local_n = -1;                    // points of the resident local map (-1: none)
    float local_bbox[4] = {0, 0, 0, 0};  // min x, min y, max x, max y of the local map
    // count / bbox: result of the last front-end kernel; scan_n / mp_n: the scan's and the map points' sizes as the
    // DEVICE knows them -- authoritative while scan_n_dev / mp_n_dev is set (asynchronous scan loop: the host then
    // only holds upper bounds in nbeams / mp_size until the next call that reads them back)
    struct FrontOut { int count; int scan_n; int mp_n; int pad; float bbox[4]; } *d_front = nullptr, *h_front = nullptr;
    bool scan_n_dev = false, mp_n_dev = false;
    // CSV ingest (csv.cu): raw text, parsed values (resident), scratch
    char *d_csv_text = nullptr;
    float *d_csv_values = nullptr;
    unsigned long long *d_csv_scratch = nullptr, *h_csv_scratch = nullptr;
    size_t csv_text_cap = 0, csv_values_cap = 0;
    int64_t csv_count = 0;
    // b200slam_fastmatch_pair_async bookkeeping (host side)
    struct { bool valid = false; float guess[3], step_a[3], step_b[3]; } pair;
    MatchHost *h_result = nullptr;        // mapped pinned (host pointer == device pointer under UVA)
    unsigned long long result_seq = 0;
    // b200slam_scan_chain_*: device pose state, result ring (mapped pinned), whether the device's cosf / sinf have
    // been checked against the running libm (0 = not yet, 1 = identical, -1 = differ: the chain is refused)
    ChainDev *d_chain = nullptr;
    ChainSlot *h_chain_ring = nullptr;
    int chain_trig_ok = 0;
    float chain_mini_dt = 0.0f, chain_mini_dr = 0.0f;

    // generic EDT scratch (u16 column distances)
    uint16_t *d_edt_scratch = nullptr;
    size_t edt_scratch_cap = 0;
    // one-shot host EDT: cached map
    b200slam_map *edt_map = nullptr;

    // timing events (lazily created)
    cudaEvent_t timing[4096] = {};

    // NCCL (dlopen'ed lazily; see comm.cu)
    void *nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    // NVLink peer exchange (set up by b200slam_comm_init when P2P + CUDA IPC work)
    XchgBuf *d_xchg = nullptr;            // this rank's buffer
    XchgBuf **d_peers = nullptr;          // device array [nranks]
    XchgBuf *peer_ptrs[XCHG_MAX_RANKS] = {};   // host copy (for closing the IPC mappings)
    bool p2p_ready = false;
    int posted_uncollected = 0;           // post-only matches queued since the last blocking collect
    unsigned long long spin_timeout_ns = 30ull * 1000 * 1000 * 1000;   // B200SLAM_SPIN_TIMEOUT_MS
    float h_param_tab[960] = {};          // host copy of a by-parameter lattice's axis tables (no DMA reads it)
};

int b200slam_set_error(b200slam_ctx *ctx, int code, const char *fmt, ...);

#define CUDA_TRY(ctx, expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return b200slam_set_error((ctx), B200SLAM_ERR_CUDA, "%s:%d %s -> %s", __FILE__,  \
                                      __LINE__, #expr, cudaGetErrorString(_e));              \
    } while (0)

#define LAUNCH_CHECK(ctx)                                                                    \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        (ctx)->prev_launch_was_lattice = false;                                              \
        (ctx)->prev_launch_was_edt = false;                                                  \
        CUDA_TRY((ctx), cudaGetLastError());                                                 \
    } while (0)

// ---- internal entry points between translation units ---------------------------------
int edt_launch(b200slam_ctx *ctx, const int32_t *d_occ, int occ_pitch, float *d_field,
               int field_pitch, int rows, int cols, float max_dist);
// The same reading the byte shadow of the occupancy (b200slam_map::d_occ8, pitch in bytes a multiple of 16).
int edt_launch_bytes(b200slam_ctx *ctx, const uint8_t *d_occ8, int occ8_pitch, float *d_field,
                     int field_pitch, int rows, int cols, float max_dist);
// int32 occupancy -> byte shadow (rows x cols in use)
int occ_pack_launch(b200slam_ctx *ctx, const b200slam_map *map);

// Output rows [row_begin, row_end) only, each row also stored into `npeers` (<= 7) other fields
// of identical layout (peer GPUs' copies of the map, mapped through CUDA IPC).
int edt_launch_rows(b200slam_ctx *ctx, const int32_t *d_occ, int occ_pitch, float *d_field,
                    int field_pitch, int rows, int cols, float max_dist, int row_begin, int row_end,
                    float *const *peer_fields, int npeers);

struct LatticeLaunch {
    const b200slam_map *map;
    int nth, ntx, nty;
    // axis tables ct[nth_tab] | st[nth_tab] | sxt[ntx] | syt[nty], theta entries only for the
    // nth_tab angles [th_first, th_first + nth_tab) this launch's rows touch: host copy (passed
    // as kernel parameters when small enough) or, for large lattices, a device copy
    const float *h_tables;
    const float *d_tables;
    int th_first, nth_tab;
    int64_t row_begin, row_end;
    float *d_scores;   // optional
    bool exchange;     // post the result to the other ranks through peer memory from the kernel's tail
    bool collect_prev; // ... and merge the previous, deferred exchange in the same tail
    bool post_deferred; // record the result in the outbox only; the next collect kernel posts it
    bool host_result = false;   // the kernel's tail also writes the result into ctx->h_result (mapped host memory)
    bool seeded = false;   // 3 x 3 x 3 lattice centred on the winner of the match in front: h_tables holds all three
                           // candidate table sets (36 floats), the kernel picks by the previous key
};
constexpr size_t LATTICE_PARAM_FLOATS = 960;
int lattice_launch(b200slam_ctx *ctx, const LatticeLaunch &L);
int exchange_collect_launch(b200slam_ctx *ctx);
struct ChainLaunch { int scan_index, count; float step_a[3], step_b[3]; };
int scan_chain_launch(b200slam_ctx *ctx, const b200slam_map *ma, const b200slam_map *mb, const ChainLaunch &C,
                      const float *d_ranges, int max_range);
int chain_trig_selftest(b200slam_ctx *ctx);
int scan_step_launch(b200slam_ctx *ctx, const b200slam_map *ma, const b200slam_map *mb, const float *tables12,
                     const float *tables36, const float *d_ranges, int max_range);
// B200SLAM_ERR_STATE (with a message) when a bounded device-side wait has given up since the last comm_init.
int device_error_check(b200slam_ctx *ctx, unsigned int error_bits);
inline XchgArgs xchg_args(const b200slam_ctx *ctx)
{
    XchgArgs X;
    X.peers = ctx->d_peers; X.nranks = ctx->nranks; X.rank = ctx->rank; X.timeout_ns = ctx->spin_timeout_ns;
    return X;
}
// post_sharded: the kernel's tail posts {arg-min key, P} to every peer (first exchange of a sharded filter step)
int poses_launch(b200slam_ctx *ctx, const b200slam_map *map, int64_t P, int64_t index_base,
                 float *d_scores, int32_t *d_hits, bool post_sharded = false);

int particles_resample_resident(b200slam_ctx *ctx, int64_t N, float beta, uint32_t u0_q32);
// Sharded resident set: collective set-up (maps every rank's pose block through CUDA IPC) and tear-down.
int particles_share_blocks(b200slam_ctx *ctx);
void particles_unshare_blocks(b200slam_ctx *ctx);
// d_wsum layout
enum PfScalar { PF_WLOCAL = 0, PF_WGLOBAL = 1, PF_RANK_OFFSET = 2, PF_KBEGIN = 3, PF_KCOUNT = 4, PF_NGLOBAL = 5,
                PF_TICKET = 6 /* finished CTAs of the push kernel (reset by its last CTA) */, PF_SLOT_BASE = 8 /* [nranks + 1]: first resampling slot held by each rank */, PF_SCALARS = 8 + XCHG_MAX_RANKS + 1 };
int particles_weights_resample(b200slam_ctx *ctx, int64_t N, float beta, uint32_t u0_q32,
                               float *weights, uint64_t *wsum, int32_t *ancestors,
                               int64_t *k_begin, int64_t *k_count);

int rasterise_launch(b200slam_ctx *ctx, b200slam_map *map, int npoints, float min_x, float min_y,
                     float pixel_size);
// One level of OccupationalGrid from points already in ctx->d_points whose bounding box is known.
int rasterise_from_bbox(b200slam_ctx *ctx, b200slam_map *map, int npoints, const float bbox[4], float pixel_size,
                        int *rows_out, int *cols_out, float top_left_out[2]);
int ensure_points_capacity(b200slam_ctx *ctx, size_t npoints);
int ensure_scan_capacity(b200slam_ctx *ctx, int nbeams);
void frontend_free(b200slam_ctx *ctx);
void csv_free(b200slam_ctx *ctx);
int ensure_front(b200slam_ctx *ctx);
int fetch_front(b200slam_ctx *ctx);

int comm_allgather_u64(b200slam_ctx *ctx, const unsigned long long *d_send,
                       unsigned long long *d_recv, int count_per_rank);
// In-place exchange of row blocks of a field of `rows` rows x `pitch` floats: rank r owns the rows
// b200slam_shard_range(rows, nranks, r) (one ncclAllGather when the blocks are equal, grouped
// ncclBroadcasts otherwise).
int comm_gather_row_blocks(b200slam_ctx *ctx, float *d_field, int pitch, int rows);
// All ranks wait for each other on the device (flags in peer memory); needs p2p_ready.
int comm_peer_barrier(b200slam_ctx *ctx);
void comm_unshare_map(b200slam_ctx *ctx, b200slam_map *map);

// ---- NVLink peer exchange primitives (device) ------------------------------------------------
// POST: four 32-bit payload words of exchange number `epoch` are stored into EVERY peer's buffer as 8-byte
// words {data, epoch} (each store atomic and self-validating).  Fire and forget.  Whole CTA; words_smem[4].
__device__ __forceinline__ void xchg_post_words(const XchgArgs &X, unsigned int epoch, unsigned int w0, unsigned int w1,
                                                unsigned int w2, unsigned int w3, unsigned int *words_smem)
{
    const int tid = threadIdx.x;
    const int ring = (int)(epoch % XCHG_EPOCHS);
    if (tid == 0) { words_smem[0] = w0; words_smem[1] = w1; words_smem[2] = w2; words_smem[3] = w3; }
    __syncthreads();
    const unsigned long long tag = (unsigned long long)epoch << 32;
    for (int i = tid; i < 4 * X.nranks; i += blockDim.x) {
        const int r = i >> 2, w = i & 3;
        *reinterpret_cast<volatile unsigned long long *>(&X.peers[r]->slot[ring][X.rank].w[w]) = tag | words_smem[w];
    }
}
// WAIT for word w of rank r's post number `epoch` in OUR buffer.  Bounded by X.timeout_ns (0 once *error is
// set): false + error bit when the word never arrives.
__device__ __forceinline__ bool xchg_wait_word(const XchgArgs &X, unsigned int epoch, int r, int w, unsigned int *out,
                                               unsigned int *error, unsigned int error_bit)
{
    const volatile unsigned long long *src = &X.peers[X.rank]->slot[epoch % XCHG_EPOCHS][r].w[w];
    unsigned long long v = *src;
    if ((unsigned int)(v >> 32) != epoch) {
        const unsigned long long budget = *reinterpret_cast<volatile unsigned int *>(error) ? 0ull : X.timeout_ns;
        const unsigned long long t0 = global_timer_ns();
        unsigned int spins = 0;
        for (;;) {
            v = *src;
            if ((unsigned int)(v >> 32) == epoch) break;
            if ((++spins & 255u) == 0 && global_timer_ns() - t0 > budget) break;
        }
    }
    if ((unsigned int)(v >> 32) != epoch) {
        atomicOr(error, error_bit);
        return false;
    }
    *out = (unsigned int)v;
    return true;
}

// Programmatic dependent launch (PDL) between consecutive scan-matching kernels: a kernel lets
// the NEXT one in the stream start (launch latency, table construction, its whole gather loop
// -- none of which depends on this kernel) while its own last CTA is still reducing and
// tracing, and the next one waits for this one right before it touches the shared match state.
// Both are no-ops for a kernel launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_prior_grids()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Packed f32x2 arithmetic (sm_100: FMUL2 / FADD2 process two independent IEEE fp32 operations per
// instruction).  Each half is rounded exactly like the scalar op (no contraction), so using them
// changes the instruction count, never a bit of the result.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi)
{
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t f2_mul_rn(f32x2_t a, f32x2_t b)
{
    f32x2_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2_t f2_add_rn(f32x2_t a, f32x2_t b)
{
    f32x2_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2_t f2_add_rz(f32x2_t a, f32x2_t b)
{
    f32x2_t d;
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// packed arg-min key helpers (scores are sums of non-negative floats, so the IEEE bit
// pattern orders like the value and uint64 min == (lowest score, then lowest index))
__host__ __device__ inline unsigned long long pack_key(float score, unsigned int index)
{
#ifdef __CUDA_ARCH__
    return ((unsigned long long)__float_as_uint(score) << 32) | index;
#else
    uint32_t b;
    memcpy(&b, &score, 4);
    return ((unsigned long long)b << 32) | index;
#endif
}
