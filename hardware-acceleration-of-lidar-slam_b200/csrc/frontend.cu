// frontend.cu -- the steps either side of the hot path, device resident:
//   readAScan        Subsystem_1/main.c:71-95    range filter, polar -> cartesian, compaction
//   Transform        Subsystem_1/main.c:97-118   sensor frame -> world frame for a pose
//   Initialise       Subsystem_1/main.c:136-145  map points <- world-frame scan
//   ExtractLocalMap  Subsystem_1/main.c:155-198  bounding box of the scan + border, strict-compare
//                                                filter of the map points, order kept
//   map growth       Subsystem_1/main.c:942-948  append scan.tx/ty[j] for bestHits[j] > 1.5
// With these the per-scan loop of main() (main.c:859-970) keeps the scan, the map points, the
// local map, both grids and both distance fields on the device: 4 bytes per beam go up per scan
// and a pose comes back.  Every arithmetic step is the reference's: products and sums rounded
// separately, cos / sin of the beam angles and of the pose supplied by the host libm, strict
// compares, compaction in input order (so indices j mean what they mean in the reference).
// The inputs are a few thousand elements, so each step is ONE CTA: ballots + a block-wide scan
// keep the order without any inter-CTA protocol, and the cost is the launch.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int FE_THREADS = 1024;

// Order-preserving compaction helper: returns the output slot of this thread's element (valid
// only where keep is true) and advances `base` (identical in all threads) by the chunk's count.
__device__ __forceinline__ int compact_slot(bool keep, int &base, int *warp_counts)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    __syncthreads();                                   // warp_counts free again
    if (lane == 0) warp_counts[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll 8
    for (int w = 0; w < FE_THREADS / 32; ++w) {
        const int c = warp_counts[w];
        before += w < warp ? c : 0;
        total += c;
    }
    const int slot = base + before + __popc(m & ((1u << lane) - 1u));
    base += total;
    return slot;
}

// main.c:71-95.  x, y: the context's scan (what the matcher reads).
__global__ void __launch_bounds__(FE_THREADS)
scan_read_kernel(const float *__restrict__ ranges, const float *__restrict__ cos_a, const float *__restrict__ sin_a,
                 int n, float range_min, int max_range, float *__restrict__ x, float *__restrict__ y,
                 b200slam_ctx::FrontOut *out)
{
    __shared__ int warp_counts[FE_THREADS / 32];
    int base = 0;
    const float maxr = (float)max_range;               // the int is converted for the compare (main.c:78)
    for (int i0 = 0; i0 < n; i0 += FE_THREADS) {
        const int i = i0 + threadIdx.x;
        float r = 0.0f;
        bool keep = false;
        if (i < n) {
            r = ranges[i];
            keep = !((r < range_min) | (r > maxr));    // main.c:78: skip if range is bad
        }
        const int slot = compact_slot(keep, base, warp_counts);
        if (keep) {
            x[slot] = __fmul_rn(r, cos_a[i]);          // main.c:90
            y[slot] = __fmul_rn(r, sin_a[i]);          // main.c:91
        }
    }
    if (threadIdx.x == 0) { out->count = base; out->scan_n = base; }           // scan.size, main.c:94
}

__global__ void set_int_kernel(int *p, int v) { *p = v; }
__global__ void copy_int_kernel(int *dst, const int *src) { *dst = *src; }

// main.c:97-118
__global__ void __launch_bounds__(256)
scan_transform_kernel(const float *__restrict__ x, const float *__restrict__ y, int n, float ct, float st, float px,
                      float py, float *__restrict__ tx, float *__restrict__ ty, const int *__restrict__ n_dev)
{
    if (n_dev) n = *n_dev;                             // the device's own count (asynchronous scan loop)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sx = x[i], sy = y[i];
    tx[i] = __fadd_rn(__fadd_rn(__fmul_rn(ct, sx), __fmul_rn(st, sy)), px);      // main.c:115
    ty[i] = __fadd_rn(__fadd_rn(__fmul_rn(-st, sx), __fmul_rn(ct, sy)), py);     // main.c:116
}

__device__ __forceinline__ float block_reduce(float v, bool want_min, float *red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, v, s);
        v = want_min ? fminf(v, o) : fmaxf(v, o);
    }
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int w = 1; w < FE_THREADS / 32; ++w) r = want_min ? fminf(r, red[w]) : fmaxf(r, red[w]);
    return r;
}

// main.c:155-198: bounding box of the world-frame scan (strict compares seeded with point 0 ==
// min / max), +- BORDERSIZE, then the map points strictly inside, in order.  Also leaves the
// bounding box of the SELECTED points behind: OccupationalGrid starts with it (main.c:272-289).
__global__ void __launch_bounds__(FE_THREADS)
local_map_kernel(const float *__restrict__ tx, const float *__restrict__ ty, int nscan, float border,
                 const float *__restrict__ mx, const float *__restrict__ my, int nmap, float *__restrict__ lx,
                 float *__restrict__ ly, b200slam_ctx::FrontOut *out, bool nscan_dev, bool nmap_dev)
{
    __shared__ int warp_counts[FE_THREADS / 32];
    __shared__ float red[FE_THREADS / 32];
    if (nscan_dev) nscan = out->scan_n;
    if (nmap_dev) nmap = out->mp_n;
    float lo_x = INFINITY, lo_y = INFINITY, hi_x = -INFINITY, hi_y = -INFINITY;
    for (int i = threadIdx.x; i < nscan; i += FE_THREADS) {
        const float a = tx[i], b = ty[i];
        lo_x = fminf(lo_x, a); hi_x = fmaxf(hi_x, a);
        lo_y = fminf(lo_y, b); hi_y = fmaxf(hi_y, b);
    }
    const float minX = __fsub_rn(block_reduce(lo_x, true, red), border);         // main.c:179-182
    const float minY = __fsub_rn(block_reduce(lo_y, true, red), border);
    const float maxX = __fadd_rn(block_reduce(hi_x, false, red), border);
    const float maxY = __fadd_rn(block_reduce(hi_y, false, red), border);
    int base = 0;
    float sel_lo_x = INFINITY, sel_lo_y = INFINITY, sel_hi_x = -INFINITY, sel_hi_y = -INFINITY;
    for (int i0 = 0; i0 < nmap; i0 += FE_THREADS) {
        const int i = i0 + threadIdx.x;
        float a = 0.0f, b = 0.0f;
        bool keep = false;
        if (i < nmap) {
            a = mx[i]; b = my[i];
            keep = (a > minX) && (a < maxX) && (b > minY) && (b < maxY);         // main.c:189-190
        }
        const int slot = compact_slot(keep, base, warp_counts);
        if (keep) {
            lx[slot] = a; ly[slot] = b;
            sel_lo_x = fminf(sel_lo_x, a); sel_hi_x = fmaxf(sel_hi_x, a);
            sel_lo_y = fminf(sel_lo_y, b); sel_hi_y = fmaxf(sel_hi_y, b);
        }
    }
    const float b0 = block_reduce(sel_lo_x, true, red), b1 = block_reduce(sel_lo_y, true, red);
    const float b2 = block_reduce(sel_hi_x, false, red), b3 = block_reduce(sel_hi_y, false, red);
    if (threadIdx.x == 0) {
        out->count = base;                                                       // local_map.size
        out->bbox[0] = b0; out->bbox[1] = b1; out->bbox[2] = b2; out->bbox[3] = b3;
    }
}

// main.c:942-948: for j < bestHits_size (of the WINNER): bestHits[j] (hits of the LAST candidate,
// stale beyond its count) > 1.5 -> append scan.tx[j], scan.ty[j].
__global__ void __launch_bounds__(FE_THREADS)
map_grow_kernel(const float *__restrict__ hit_values, const MatchDev *__restrict__ match, const float *__restrict__ tx,
                const float *__restrict__ ty, float threshold, float *__restrict__ mx, float *__restrict__ my,
                int map_size, int map_cap, b200slam_ctx::FrontOut *out, bool size_dev)
{
    __shared__ int warp_counts[FE_THREADS / 32];
    const int n = match->key == ~0ull ? 0 : match->best_hits;
    if (size_dev) map_size = out->mp_n;
    __syncthreads();                                   // every thread has read mp_n before thread 0 rewrites it
    int base = map_size;
    for (int j0 = 0; j0 < n; j0 += FE_THREADS) {
        const int j = j0 + threadIdx.x;
        const bool keep = j < n && hit_values[j] > threshold;
        const int slot = compact_slot(keep, base, warp_counts);
        if (keep && slot < map_cap) { mx[slot] = tx[j]; my[slot] = ty[j]; }
    }
    if (threadIdx.x == 0) {
        out->count = base - map_size;                                            // newPointSize
        out->mp_n = base < map_cap ? base : map_cap;
    }
}

}  // namespace

int ensure_front(b200slam_ctx *ctx)
{
    if (ctx->d_front) return B200SLAM_OK;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_front, sizeof(*ctx->d_front)));
    CUDA_TRY(ctx, cudaMemset(ctx->d_front, 0, sizeof(*ctx->d_front)));
    CUDA_TRY(ctx, cudaHostAlloc(&ctx->h_front, sizeof(*ctx->h_front), cudaHostAllocDefault));
    return B200SLAM_OK;
}

// Reads the front-end block back; the device's scan / map-point counts replace the host's upper bounds.
int fetch_front(b200slam_ctx *ctx)
{
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_front, ctx->d_front, sizeof(*ctx->d_front), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->scan_n_dev) ctx->nbeams = ctx->h_front->scan_n;
    if (ctx->mp_n_dev) ctx->mp_size = ctx->h_front->mp_n;
    return B200SLAM_OK;
}

namespace {

int ensure_scan_t(b200slam_ctx *ctx)
{
    if (ctx->scan_t_cap >= ctx->scan_cap && ctx->d_scan_t) return B200SLAM_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_scan_t);
    ctx->d_scan_t = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_scan_t, sizeof(float) * 2 * (size_t)ctx->scan_cap));
    ctx->scan_t_cap = ctx->scan_cap;
    ctx->scan_t_valid = false;
    return B200SLAM_OK;
}

int ensure_map_points(b200slam_ctx *ctx, int need)
{
    if (need <= ctx->mp_cap) return B200SLAM_OK;
    int cap = ctx->mp_cap ? ctx->mp_cap : 20480;          // MapPoints holds 20 000 (main.c:122-123)
    while (cap < need) cap *= 2;
    float *d = nullptr;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaMalloc(&d, sizeof(float) * 2 * (size_t)cap));
    if (ctx->d_mp && ctx->mp_size > 0) {
        CUDA_TRY(ctx, cudaMemcpy(d, ctx->d_mp, sizeof(float) * ctx->mp_size, cudaMemcpyDeviceToDevice));
        CUDA_TRY(ctx, cudaMemcpy(d + cap, ctx->d_mp + ctx->mp_cap, sizeof(float) * ctx->mp_size, cudaMemcpyDeviceToDevice));
    }
    cudaFree(ctx->d_mp);
    ctx->d_mp = d;
    ctx->mp_cap = cap;
    return B200SLAM_OK;
}

}  // namespace

void frontend_free(b200slam_ctx *ctx)
{
    csv_free(ctx);
    cudaFree(ctx->d_lidar); cudaFree(ctx->d_ranges); cudaFreeHost(ctx->h_ranges);
    cudaFree(ctx->d_scan_t); cudaFree(ctx->d_mp); cudaFree(ctx->d_front); cudaFreeHost(ctx->h_front);
}

extern "C" {

int b200slam_lidar_set(b200slam_ctx *ctx, const float *cos_a, const float *sin_a, int nbeams, float range_min)
{
    if (!ctx || !cos_a || !sin_a || nbeams <= 0) return B200SLAM_ERR_ARG;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_lidar); cudaFree(ctx->d_ranges); cudaFreeHost(ctx->h_ranges);
    ctx->d_lidar = ctx->d_ranges = ctx->h_ranges = nullptr;
    ctx->lidar_n = 0;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_lidar, sizeof(float) * 2 * (size_t)nbeams));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_ranges, sizeof(float) * (size_t)nbeams));
    CUDA_TRY(ctx, cudaHostAlloc(&ctx->h_ranges, sizeof(float) * (size_t)nbeams, cudaHostAllocDefault));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_lidar, cos_a, sizeof(float) * nbeams, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_lidar + nbeams, sin_a, sizeof(float) * nbeams, cudaMemcpyHostToDevice));
    ctx->lidar_n = nbeams;
    ctx->lidar_range_min = range_min;
    return B200SLAM_OK;
}

// d_src: ranges already on the device (CSV ingest), else `ranges` (host) go through the pinned staging buffer
static int scan_read_queue(b200slam_ctx *ctx, const float *ranges, const float *d_src, int max_range)
{
    if (ctx->lidar_n <= 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_lidar_set first");
    int rc = ensure_scan_capacity(ctx, ctx->lidar_n);
    if (!rc) rc = ensure_front(ctx);
    if (rc) return rc;
    const int n = ctx->lidar_n;
    if (!d_src) {
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->scan_event));              // pinned staging free again
        memcpy(ctx->h_ranges, ranges, sizeof(float) * n);
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_ranges, ctx->h_ranges, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->scan_event, ctx->stream));
        d_src = ctx->d_ranges;
    }
    scan_read_kernel<<<1, FE_THREADS, 0, ctx->stream>>>(d_src, ctx->d_lidar, ctx->d_lidar + n, n, ctx->lidar_range_min,
                                                        max_range, ctx->d_scan_x, ctx->d_scan_y, ctx->d_front);
    LAUNCH_CHECK(ctx);
    ctx->nbeams = n;                       // upper bound until somebody reads the count back
    ctx->scan_n_dev = true;
    ctx->scan_t_valid = false;
    return B200SLAM_OK;
}

int b200slam_scan_read_async(b200slam_ctx *ctx, const float *ranges, int max_range)
{
    if (!ctx || !ranges) return B200SLAM_ERR_ARG;
    return scan_read_queue(ctx, ranges, nullptr, max_range);
}

int b200slam_scan_read_resident_async(b200slam_ctx *ctx, int64_t first_value, int max_range)
{
    if (!ctx || first_value < 0) return B200SLAM_ERR_ARG;
    if (ctx->lidar_n <= 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_lidar_set first");
    if (!ctx->d_csv_values || first_value + ctx->lidar_n > ctx->csv_count)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "values [%lld, %lld) are not in the ingested CSV (%lld values)",
                                  (long long)first_value, (long long)(first_value + ctx->lidar_n), (long long)ctx->csv_count);
    return scan_read_queue(ctx, nullptr, ctx->d_csv_values + first_value, max_range);
}

int b200slam_scan_read(b200slam_ctx *ctx, const float *ranges, int max_range, int *size)
{
    if (!ctx || !ranges) return B200SLAM_ERR_ARG;
    int rc = scan_read_queue(ctx, ranges, nullptr, max_range);
    if (rc) return rc;
    rc = fetch_front(ctx);                                              // ctx->nbeams <- the device's count
    if (rc) return rc;
    if (size) *size = ctx->nbeams;
    return B200SLAM_OK;
}

int b200slam_scan_transform(b200slam_ctx *ctx, const float pose[3])
{
    if (!ctx || !pose) return B200SLAM_ERR_ARG;
    if (!ctx->d_scan_x || ctx->nbeams < 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no scan on the device");
    int rc = ensure_scan_t(ctx);
    if (rc) return rc;
    const float ct = cosf(pose[2]), st = sinf(pose[2]);                 // main.c:101-102 (host libm)
    if (ctx->nbeams > 0) {
        scan_transform_kernel<<<(ctx->nbeams + 255) / 256, 256, 0, ctx->stream>>>(
            ctx->d_scan_x, ctx->d_scan_y, ctx->nbeams, ct, st, pose[0], pose[1], ctx->d_scan_t,
            ctx->d_scan_t + ctx->scan_t_cap, ctx->scan_n_dev ? &ctx->d_front->scan_n : nullptr);
        LAUNCH_CHECK(ctx);
    }
    ctx->scan_t_valid = true;
    return B200SLAM_OK;
}

int b200slam_scan_download(b200slam_ctx *ctx, float *x, float *y, float *tx, float *ty, int *size)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (!ctx->d_scan_x || ctx->nbeams < 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no scan on the device");
    if (ctx->scan_n_dev) {
        int rc = fetch_front(ctx);
        if (rc) return rc;
    }
    const size_t b = sizeof(float) * (size_t)ctx->nbeams;
    if ((tx || ty) && !ctx->scan_t_valid) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_scan_transform first");
    if (b) {
        if (x) CUDA_TRY(ctx, cudaMemcpyAsync(x, ctx->d_scan_x, b, cudaMemcpyDeviceToHost, ctx->stream));
        if (y) CUDA_TRY(ctx, cudaMemcpyAsync(y, ctx->d_scan_y, b, cudaMemcpyDeviceToHost, ctx->stream));
        if (tx) CUDA_TRY(ctx, cudaMemcpyAsync(tx, ctx->d_scan_t, b, cudaMemcpyDeviceToHost, ctx->stream));
        if (ty) CUDA_TRY(ctx, cudaMemcpyAsync(ty, ctx->d_scan_t + ctx->scan_t_cap, b, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (size) *size = ctx->nbeams;
    return B200SLAM_OK;
}

int b200slam_mappoints_upload(b200slam_ctx *ctx, const float *x, const float *y, int n, int offset)
{
    if (!ctx || n < 0 || offset < 0 || offset > ctx->mp_size || (n > 0 && (!x || !y))) return B200SLAM_ERR_ARG;
    int rc = ensure_map_points(ctx, offset + n);
    if (rc) return rc;
    if (n > 0) {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_mp + offset, x, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_mp + ctx->mp_cap + offset, y, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));             // x / y may be pageable: done with them
    }
    ctx->mp_size = offset + n;
    ctx->mp_n_dev = false;
    return B200SLAM_OK;
}

int b200slam_mappoints_from_scan(b200slam_ctx *ctx)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (!ctx->scan_t_valid) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_scan_transform first");
    int rc = ensure_map_points(ctx, ctx->nbeams);
    if (rc) return rc;
    if (ctx->nbeams > 0) {                                              // main.c:137-141
        const size_t b = sizeof(float) * (size_t)ctx->nbeams;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_mp, ctx->d_scan_t, b, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_mp + ctx->mp_cap, ctx->d_scan_t + ctx->scan_t_cap, b, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ctx->mp_size = ctx->nbeams;
    ctx->mp_n_dev = ctx->scan_n_dev;       // the device's scan count is the map's size now
    if (ctx->mp_n_dev) {
        copy_int_kernel<<<1, 1, 0, ctx->stream>>>(&ctx->d_front->mp_n, &ctx->d_front->scan_n);
        LAUNCH_CHECK(ctx);
    }
    return B200SLAM_OK;
}

int b200slam_mappoints_download(b200slam_ctx *ctx, float *x, float *y, int *size)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (ctx->mp_n_dev) {
        int rc = fetch_front(ctx);
        if (rc) return rc;
    }
    if (ctx->mp_size > 0) {
        const size_t b = sizeof(float) * (size_t)ctx->mp_size;
        if (x) CUDA_TRY(ctx, cudaMemcpyAsync(x, ctx->d_mp, b, cudaMemcpyDeviceToHost, ctx->stream));
        if (y) CUDA_TRY(ctx, cudaMemcpyAsync(y, ctx->d_mp + ctx->mp_cap, b, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (size) *size = ctx->mp_size;
    return B200SLAM_OK;
}

static int mappoints_grow_queue(b200slam_ctx *ctx, float threshold)
{
    if (!ctx->scan_t_valid) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_scan_transform first");
    if (!ctx->last.valid || ctx->last.is_poses) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no lattice match to grow from");
    // main.c:942-948 reads the winner's count and the LAST candidate's hit values of the WHOLE lattice; after a
    // sharded match both may live on another rank
    if (ctx->last.exchanged || ctx->last.gathered)
        return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "map growth needs a match scored entirely on this GPU");
    int rc = ensure_map_points(ctx, ctx->mp_size + ctx->nbeams);
    if (!rc) rc = ensure_front(ctx);
    if (rc) return rc;
    map_grow_kernel<<<1, FE_THREADS, 0, ctx->stream>>>(ctx->d_hit_values + ctx->scan_cap, ctx->d_match, ctx->d_scan_t,
                                                       ctx->d_scan_t + ctx->scan_t_cap, threshold, ctx->d_mp,
                                                       ctx->d_mp + ctx->mp_cap, ctx->mp_size, ctx->mp_cap, ctx->d_front,
                                                       ctx->mp_n_dev);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}

int b200slam_mappoints_grow(b200slam_ctx *ctx, float threshold, int *added)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    const int before = ctx->mp_size;
    const bool dev = ctx->mp_n_dev;
    int rc = mappoints_grow_queue(ctx, threshold);
    if (rc) return rc;
    rc = fetch_front(ctx);                  // device-side size: ctx->mp_size <- exact
    if (rc) return rc;
    if (!dev) ctx->mp_size = before + ctx->h_front->count;
    if (added) *added = ctx->h_front->count;
    return B200SLAM_OK;
}

int b200slam_mappoints_grow_async(b200slam_ctx *ctx, float threshold)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    int rc = ensure_front(ctx);
    if (rc) return rc;
    if (!ctx->mp_n_dev) {                   // from here on the device keeps the size
        set_int_kernel<<<1, 1, 0, ctx->stream>>>(&ctx->d_front->mp_n, ctx->mp_size);
        LAUNCH_CHECK(ctx);
        ctx->mp_n_dev = true;
    }
    rc = mappoints_grow_queue(ctx, threshold);
    if (rc) return rc;
    ctx->mp_size += ctx->nbeams;            // upper bound until the size is read back
    return B200SLAM_OK;
}

int b200slam_local_map_extract(b200slam_ctx *ctx, float border, int *size)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (!ctx->scan_t_valid || ctx->nbeams <= 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "needs a transformed, non-empty scan");
    ctx->local_n = -1;
    int rc = ensure_points_capacity(ctx, (size_t)(ctx->mp_size > 0 ? ctx->mp_size : 1));
    if (!rc) rc = ensure_front(ctx);
    if (rc) return rc;
    local_map_kernel<<<1, FE_THREADS, 0, ctx->stream>>>(ctx->d_scan_t, ctx->d_scan_t + ctx->scan_t_cap, ctx->nbeams, border,
                                                        ctx->d_mp, ctx->d_mp + ctx->mp_cap, ctx->mp_size, ctx->d_points,
                                                        ctx->d_points + ctx->points_cap, ctx->d_front, ctx->scan_n_dev,
                                                        ctx->mp_n_dev);
    LAUNCH_CHECK(ctx);
    rc = fetch_front(ctx);                  // also replaces the host's upper bounds of the scan / map sizes
    if (rc) return rc;
    ctx->local_n = ctx->h_front->count;
    for (int i = 0; i < 4; ++i) ctx->local_bbox[i] = ctx->h_front->bbox[i];
    if (size) *size = ctx->local_n;
    return B200SLAM_OK;
}

int b200slam_local_map_download(b200slam_ctx *ctx, float *x, float *y, int *size)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (ctx->local_n < 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no local map on the device");
    if (ctx->local_n > 0) {
        const size_t b = sizeof(float) * (size_t)ctx->local_n;
        if (x) CUDA_TRY(ctx, cudaMemcpyAsync(x, ctx->d_points, b, cudaMemcpyDeviceToHost, ctx->stream));
        if (y) CUDA_TRY(ctx, cudaMemcpyAsync(y, ctx->d_points + ctx->points_cap, b, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (size) *size = ctx->local_n;
    return B200SLAM_OK;
}

int b200slam_map_rasterise_local(b200slam_ctx *ctx, b200slam_map *map, float pixel_size, int *rows, int *cols,
                                 float top_left[2])
{
    if (!ctx || !map || !(pixel_size > 0.0f)) return B200SLAM_ERR_ARG;
    if (ctx->local_n <= 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no (non-empty) local map on the device");
    return rasterise_from_bbox(ctx, map, ctx->local_n, ctx->local_bbox, pixel_size, rows, cols, top_left);
}

}  // extern "C"
