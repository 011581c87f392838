// api.cu -- the C ABI of libb200slam.so (include/b200slam.h): context, device-resident
// maps, and the host-side orchestration of the kernels in edt.cu / score.cu /
// particles.cu.  Host-side arithmetic that the reference does on the CPU and whose bits
// matter (lattice axis values, cosf/sinf, (t - min) * ipixel; Subsystem_1/main.c:424-438)
// is done here with the same float operations and the host libm.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <time.h>
#if defined(__SSE__) || defined(__x86_64__)
#include <xmmintrin.h>
#define B200SLAM_HAVE_SSE 1
#endif

#include <new>

#include "common.cuh"

static char g_create_error[512] = "";

int b200slam_set_error(b200slam_ctx *ctx, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx ? ctx->err : g_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

int device_error_check(b200slam_ctx *ctx, unsigned int bits)
{
    if (!bits) return B200SLAM_OK;
    return b200slam_set_error(ctx, B200SLAM_ERR_STATE,
                              "a bounded device-side wait gave up after %.0f ms:%s%s%s%s (sticky until b200slam_comm_init)",
                              ctx->spin_timeout_ns * 1e-6,
                              bits & DEV_ERR_EXCHANGE ? " a peer's match result never arrived;" : "",
                              bits & DEV_ERR_BARRIER ? " a peer never reached a device barrier;" : "",
                              bits & DEV_ERR_TMA ? " a TMA load of the EDT never completed;" : "",
                              bits & DEV_ERR_PARTICLES ? " a peer's particle sums / offspring never arrived;" : "");
}

extern "C" {

int b200slam_abi_version(void) { return B200SLAM_ABI_VERSION; }

const char *b200slam_last_error(const b200slam_ctx *ctx) { return ctx ? ctx->err : g_create_error; }

int b200slam_create(b200slam_ctx **out, int device)
{
    if (!out) return B200SLAM_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return b200slam_set_error(nullptr, B200SLAM_ERR_CUDA,
                                  "no CUDA device (%s); libb200slam has no CPU fallback",
                                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev)
        return b200slam_set_error(nullptr, B200SLAM_ERR_ARG, "device %d out of range [0,%d)", device, ndev);
    b200slam_ctx *ctx = new (std::nothrow) b200slam_ctx();
    if (!ctx) return B200SLAM_ERR_NOMEM;
    ctx->device = device;
    ctx->use_pdl = getenv("B200SLAM_NO_PDL") == nullptr;
    if (const char *e = getenv("B200SLAM_SPIN_TIMEOUT_MS"))
        if (atof(e) > 0) ctx->spin_timeout_ns = (unsigned long long)(atof(e) * 1e6);
#define CREATE_TRY(expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            b200slam_set_error(nullptr, B200SLAM_ERR_CUDA, "%s -> %s", #expr,                 \
                               cudaGetErrorString(_e));                                       \
            b200slam_destroy(ctx);                                                            \
            return B200SLAM_ERR_CUDA;                                                         \
        }                                                                                     \
    } while (0)
    CREATE_TRY(cudaSetDevice(device));
    CREATE_TRY(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaMalloc(&ctx->d_match, sizeof(MatchDev)));
    CREATE_TRY(cudaHostAlloc(&ctx->h_match, sizeof(MatchDev), cudaHostAllocDefault));
    {
        MatchDev init;
        init.work_key = ~0ull; init.tickets = 0; init.epoch = 0; init.key = ~0ull;
        init.best_hits = 0; init.last_hits = 0; init.collected = 0; init.error = 0; init.written_hits = 0; init.posted = 0;
        memset(init.cand_hits, 0, sizeof init.cand_hits); memset(init.outbox, 0, sizeof init.outbox);
        init.bar_epoch = 0; init.seed_key = ~0ull; init.gkey = ~0ull; init.gbest_hits = 0; init.glast_hits = 0;
        CREATE_TRY(cudaMemcpy(ctx->d_match, &init, sizeof init, cudaMemcpyHostToDevice));
    }
    CREATE_TRY(cudaMalloc(&ctx->d_keys, sizeof(unsigned long long) * 256));
    CREATE_TRY(cudaHostAlloc(&ctx->h_keys, sizeof(unsigned long long) * 128, cudaHostAllocDefault));
    CREATE_TRY(cudaMalloc(&ctx->d_wsum, sizeof(unsigned long long) * PF_SCALARS));
    CREATE_TRY(cudaMemset(ctx->d_wsum, 0, sizeof(unsigned long long) * PF_SCALARS));
    CREATE_TRY(cudaHostAlloc(&ctx->h_wsum, sizeof(unsigned long long) * 2 * 64, cudaHostAllocDefault));
#undef CREATE_TRY
    *out = ctx;
    return B200SLAM_OK;
}

void b200slam_destroy(b200slam_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    b200slam_comm_destroy(ctx);
    if (ctx->edt_map) b200slam_map_destroy(ctx, ctx->edt_map);
    cudaFree(ctx->d_scan_x); cudaFreeHost(ctx->h_scan); cudaFreeHost(ctx->h_hit_values);
    if (ctx->scan_event) cudaEventDestroy(ctx->scan_event);
    cudaFreeHost(ctx->h_lat); cudaFree(ctx->d_lat);
    cudaFree(ctx->d_match); cudaFreeHost(ctx->h_match); cudaFreeHost(ctx->h_result);
    cudaFree(ctx->d_chain); cudaFreeHost(ctx->h_chain_ring);
    cudaFree(ctx->d_keys); cudaFreeHost(ctx->h_keys); cudaFree(ctx->d_hit_values);
    cudaFree(ctx->d_scores);
    particles_unshare_blocks(ctx);
    cudaFree(ctx->d_pf_peers);
    cudaFree(ctx->d_pose_block); cudaFree(ctx->d_hits); cudaFreeHost(ctx->h_pose_stage);
    cudaFree(ctx->d_q); cudaFree(ctx->d_block_sums); cudaFree(ctx->d_weights);
    cudaFree(ctx->d_ancestors); cudaFree(ctx->d_wsum); cudaFreeHost(ctx->h_wsum);
    cudaFree(ctx->d_edt_scratch);
    cudaFree(ctx->d_points); cudaFreeHost(ctx->h_points);
    frontend_free(ctx);
    for (int i = 0; i < LAT_SLOTS; ++i)
        if (ctx->lat_event[i]) cudaEventDestroy(ctx->lat_event[i]);
    for (int i = 0; i < 4096; ++i)
        if (ctx->timing[i]) cudaEventDestroy(ctx->timing[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int b200slam_sync(b200slam_ctx *ctx)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    // the sticky device-side error word rides along: a bounded wait that gave up is reported here
    CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->h_match->error, &ctx->d_match->error, sizeof(unsigned int),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return device_error_check(ctx, ctx->h_match->error);
}

void *b200slam_stream(b200slam_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int b200slam_set_match_mode(b200slam_ctx *ctx, int mode)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (mode != B200SLAM_MATCH_LATENCY && mode != B200SLAM_MATCH_THROUGHPUT)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "unknown match mode %d", mode);
    ctx->match_mode = mode;
    return B200SLAM_OK;
}

uint64_t b200slam_launch_count(const b200slam_ctx *ctx) { return ctx ? ctx->launches : 0; }

int b200slam_device_info(const b200slam_ctx *ctx, char *name, int *sm_count, int *cc_major,
                         int *cc_minor, size_t *total_mem)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, ctx->device) != cudaSuccess) return B200SLAM_ERR_CUDA;
    if (name) { strncpy(name, p.name, 63); name[63] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return B200SLAM_OK;
}

int b200slam_host_alloc(b200slam_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return B200SLAM_ERR_ARG;
    CUDA_TRY(ctx, cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return B200SLAM_OK;
}

int b200slam_host_free(b200slam_ctx *ctx, void *p)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    CUDA_TRY(ctx, cudaFreeHost(p));
    return B200SLAM_OK;
}

/* ---- maps ---------------------------------------------------------------------- */

int b200slam_map_create(b200slam_ctx *ctx, int rows, int cols, b200slam_map **out)
{
    if (!ctx || !out) return B200SLAM_ERR_ARG;
    *out = nullptr;
    if (rows <= 0 || cols <= 0 || (long long)rows * (((long long)cols + 31) & ~31ll) >= (1ll << 30))
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "map %d x %d unsupported (need 0 < cells < 2^30)",
                                  rows, cols);
    b200slam_map *m = new (std::nothrow) b200slam_map();
    if (!m) return B200SLAM_ERR_NOMEM;
    m->rows = m->cap_rows = rows;
    m->cols = m->cap_cols = cols;
    m->occ_pitch = (cols + 31) & ~31;        // rows start on 128-byte lines
    m->field_pitch = (cols + 31) & ~31;
    cudaError_t e = cudaMalloc(&m->d_occ, sizeof(int32_t) * (size_t)m->occ_pitch * rows);
    if (e == cudaSuccess)
        e = cudaMalloc(&m->d_field_alloc,
                       sizeof(float) * ((size_t)m->field_pitch * rows + field_pad_floats(m->field_pitch)));
    if (e == cudaSuccess)
        e = cudaMemsetAsync(m->d_field_alloc, 0, sizeof(float) * field_pad_floats(m->field_pitch), ctx->stream);
    if (e == cudaSuccess && (long long)rows * cols >= (1ll << 20) && !getenv("B200SLAM_EDT_NO_BYTES"))
        e = cudaMalloc(&m->d_occ8, (size_t)m->occ_pitch * rows);      // byte shadow: big maps only (HBM-bound transform)
    if (e != cudaSuccess) {
        cudaFree(m->d_occ8);
        cudaFree(m->d_occ);
        cudaFree(m->d_field_alloc);
        delete m;
        return b200slam_set_error(ctx, B200SLAM_ERR_CUDA, "map alloc %d x %d -> %s", rows, cols,
                                  cudaGetErrorString(e));
    }
    m->d_field = m->d_field_alloc + field_pad_floats(m->field_pitch);
    *out = m;
    return B200SLAM_OK;
}

void b200slam_map_destroy(b200slam_ctx *ctx, b200slam_map *map)
{
    if (!map) return;
    if (ctx && ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (map->shared_nranks) comm_unshare_map(ctx, map);
    cudaFree(map->d_occ);
    cudaFree(map->d_occ8);
    cudaFree(map->d_field_alloc);
    cudaFree(map->d_raster_cells);
    delete map;
}

int b200slam_map_resize(b200slam_map *map, int rows, int cols)
{
    if (!map || rows <= 0 || cols <= 0 || rows > map->cap_rows || cols > map->cap_cols) return B200SLAM_ERR_ARG;
    map->rows = rows;
    map->cols = cols;
    return B200SLAM_OK;
}

int b200slam_map_set_geometry(b200slam_map *map, float pixel_size, float top_left_x, float top_left_y)
{
    if (!map || !(pixel_size > 0.0f)) return B200SLAM_ERR_ARG;
    map->pixel_size = pixel_size;
    map->top_left_x = top_left_x;
    map->top_left_y = top_left_y;
    map->has_geometry = true;
    return B200SLAM_OK;
}

int b200slam_map_upload_occupancy(b200slam_ctx *ctx, b200slam_map *map, const int32_t *occ, int stride)
{
    if (!ctx || !map || !occ || stride < map->cols) return B200SLAM_ERR_ARG;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(map->d_occ, sizeof(int32_t) * (size_t)map->occ_pitch, occ,
                                    sizeof(int32_t) * (size_t)stride, sizeof(int32_t) * (size_t)map->cols,
                                    map->rows, cudaMemcpyHostToDevice, ctx->stream));
    map->raster_cells_n = -1;                 // contents no longer "zero except the rasterised cells"
    map->occ8_valid = false;
    if (map->d_occ8 && !map->occ_exposed) {
        int rc = occ_pack_launch(ctx, map);
        if (rc) return rc;
        map->occ8_valid = true;
        map->occ8_rows = map->rows; map->occ8_cols = map->cols;
    }
    return B200SLAM_OK;
}

}  // extern "C" (internal helpers below have C++ linkage)

int ensure_points_capacity(b200slam_ctx *ctx, size_t npoints)
{
    if (npoints <= ctx->points_cap) return B200SLAM_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t cap = (npoints + 4095) & ~(size_t)4095;
    float *d = nullptr, *h = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d, sizeof(float) * 2 * cap));
    CUDA_TRY(ctx, cudaHostAlloc(&h, sizeof(float) * 2 * cap, cudaHostAllocDefault));
    if (ctx->d_points && ctx->local_n > 0) {       // keep a resident local map across the growth
        CUDA_TRY(ctx, cudaMemcpy(d, ctx->d_points, sizeof(float) * ctx->local_n, cudaMemcpyDeviceToDevice));
        CUDA_TRY(ctx, cudaMemcpy(d + cap, ctx->d_points + ctx->points_cap, sizeof(float) * ctx->local_n,
                                 cudaMemcpyDeviceToDevice));
    }
    cudaFree(ctx->d_points); cudaFreeHost(ctx->h_points);
    ctx->d_points = d; ctx->h_points = h;
    ctx->points_cap = cap;
    return B200SLAM_OK;
}

// main.c:272-289: bounding box of the points -- strict compares, seeded with point 0.  Minimum and maximum do not
// depend on the order the points are visited in (a NaN never passes a strict compare, a NaN seed stays; which
// zero survives among +0 / -0 changes nothing once the margin is subtracted), so the walk is 4 lanes wide:
// _mm_min_ps(v, acc) is exactly `v < acc ? v : acc`.
static void points_bbox(const float *x, const float *y, int n, float bbox[4])
{
#ifndef B200SLAM_HAVE_SSE
    bbox[0] = bbox[2] = x[0]; bbox[1] = bbox[3] = y[0];
    for (int a = 0; a < n; ++a) {
        if (x[a] < bbox[0]) bbox[0] = x[a];
        if (x[a] > bbox[2]) bbox[2] = x[a];
        if (y[a] < bbox[1]) bbox[1] = y[a];
        if (y[a] > bbox[3]) bbox[3] = y[a];
    }
#else
    __m128 lox = _mm_set1_ps(x[0]), hix = lox, loy = _mm_set1_ps(y[0]), hiy = loy;
    int a = 0;
    for (; a + 4 <= n; a += 4) {
        const __m128 vx = _mm_loadu_ps(x + a), vy = _mm_loadu_ps(y + a);
        lox = _mm_min_ps(vx, lox); hix = _mm_max_ps(vx, hix);
        loy = _mm_min_ps(vy, loy); hiy = _mm_max_ps(vy, hiy);
    }
    float l[4][4];
    _mm_storeu_ps(l[0], lox); _mm_storeu_ps(l[1], loy); _mm_storeu_ps(l[2], hix); _mm_storeu_ps(l[3], hiy);
    for (int k = 0; k < 4; ++k) bbox[k] = l[k][0];
    for (int k = 0; k < 2; ++k)
        for (int j = 1; j < 4; ++j) {
            if (l[k][j] < bbox[k]) bbox[k] = l[k][j];
            if (l[k + 2][j] > bbox[k + 2]) bbox[k + 2] = l[k + 2][j];
        }
    for (; a < n; ++a) {
        if (x[a] < bbox[0]) bbox[0] = x[a];
        if (x[a] > bbox[2]) bbox[2] = x[a];
        if (y[a] < bbox[1]) bbox[1] = y[a];
        if (y[a] > bbox[3]) bbox[3] = y[a];
    }
#endif
}

int rasterise_from_bbox(b200slam_ctx *ctx, b200slam_map *map, int npoints, const float bbox[4], float pixel_size,
                        int *rows_out, int *cols_out, float top_left_out[2])
{
    float minXY[2] = {bbox[0], bbox[1]}, maxXY[2] = {bbox[2], bbox[3]};
    // main.c:296-305: 3-pixel margin; size = (int)roundf(extent / pixel) + 1, every step rounded to float
    int S[2];
    for (int a = 0; a < 2; ++a) {
        volatile float margin = 3 * pixel_size;
        volatile float lo = minXY[a] - margin, hi = maxXY[a] + margin;
        volatile float ext = hi - lo;
        volatile float q = ext / pixel_size;
        minXY[a] = lo;
        S[a] = (int)roundf(q) + 1;
    }
    const int cols = S[0], rows = S[1];                                    // main.c:313-314
    if (rows_out) *rows_out = rows;
    if (cols_out) *cols_out = cols;
    if (top_left_out) { top_left_out[0] = minXY[0]; top_left_out[1] = minXY[1]; }
    if (rows > map->cap_rows || cols > map->cap_cols || rows <= 0 || cols <= 0)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "rasterised grid %d x %d exceeds the map capacity %d x %d",
                                  rows, cols, map->cap_rows, map->cap_cols);
    map->rows = rows;
    map->cols = cols;
    map->pixel_size = pixel_size;                                          // main.c:357-362
    map->top_left_x = minXY[0];
    map->top_left_y = minXY[1];
    map->has_geometry = true;
    return rasterise_launch(ctx, map, npoints, minXY[0], minXY[1], pixel_size);
}

extern "C" {

int b200slam_map_rasterise(b200slam_ctx *ctx, b200slam_map *map, const float *x, const float *y, int npoints,
                           float pixel_size, int *rows_out, int *cols_out, float top_left_out[2])
{
    if (!ctx || !map || !x || !y || npoints <= 0 || !(pixel_size > 0.0f)) return B200SLAM_ERR_ARG;
    float bbox[4];
    points_bbox(x, y, npoints, bbox);
    int rc = ensure_points_capacity(ctx, (size_t)npoints);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));          // pinned staging buffer free again
    memcpy(ctx->h_points, x, sizeof(float) * npoints);
    memcpy(ctx->h_points + ctx->points_cap, y, sizeof(float) * npoints);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_points, ctx->h_points, sizeof(float) * npoints, cudaMemcpyHostToDevice,
                                  ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_points + ctx->points_cap, ctx->h_points + ctx->points_cap,
                                  sizeof(float) * npoints, cudaMemcpyHostToDevice, ctx->stream));
    ctx->local_n = npoints;                                     // these points ARE the resident local map now
    for (int i = 0; i < 4; ++i) ctx->local_bbox[i] = bbox[i];
    return rasterise_from_bbox(ctx, map, npoints, bbox, pixel_size, rows_out, cols_out, top_left_out);
}

int b200slam_map_rasterise_async(b200slam_ctx *ctx, b200slam_map *map, const float *x, const float *y, int npoints,
                                 float pixel_size, int *rows_out, int *cols_out, float top_left_out[2])
{
    if (!ctx || !map || !x || !y || npoints <= 0 || !(pixel_size > 0.0f)) return B200SLAM_ERR_ARG;
    cudaPointerAttributes ax, ay;
    if (cudaPointerGetAttributes(&ax, x) != cudaSuccess || cudaPointerGetAttributes(&ay, y) != cudaSuccess ||
        ax.type != cudaMemoryTypeHost || ay.type != cudaMemoryTypeHost) {
        cudaGetLastError();
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG,
                                  "b200slam_map_rasterise_async: x and y must be page-locked host memory "
                                  "(b200slam_host_alloc); use b200slam_map_rasterise for pageable arrays");
    }
    float bbox[4];
    points_bbox(x, y, npoints, bbox);
    int rc = ensure_points_capacity(ctx, (size_t)npoints);
    if (rc) return rc;
    // queued straight from the caller's page-locked arrays: no staging copy, no wait
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_points, x, sizeof(float) * npoints, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_points + ctx->points_cap, y, sizeof(float) * npoints, cudaMemcpyHostToDevice,
                                  ctx->stream));
    ctx->local_n = npoints;
    for (int i = 0; i < 4; ++i) ctx->local_bbox[i] = bbox[i];
    return rasterise_from_bbox(ctx, map, npoints, bbox, pixel_size, rows_out, cols_out, top_left_out);
}

int b200slam_map_edt(b200slam_ctx *ctx, b200slam_map *map, float max_dist)
{
    if (!ctx || !map) return B200SLAM_ERR_ARG;
    if (map->d_occ8 && map->occ8_valid && map->rows <= map->occ8_rows && map->cols <= map->occ8_cols && max_dist > 1.0f &&
        max_dist <= 15.0f)
        return edt_launch_bytes(ctx, map->d_occ8, map->occ_pitch, map->d_field, map->field_pitch, map->rows, map->cols, max_dist);
    return edt_launch(ctx, map->d_occ, map->occ_pitch, map->d_field, map->field_pitch, map->rows,
                      map->cols, max_dist);
}

int b200slam_map_download_field(b200slam_ctx *ctx, b200slam_map *map, float *out, int stride)
{
    if (!ctx || !map || !out || stride < map->cols) return B200SLAM_ERR_ARG;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(out, sizeof(float) * (size_t)stride, map->d_field,
                                    sizeof(float) * (size_t)map->field_pitch,
                                    sizeof(float) * (size_t)map->cols, map->rows,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200SLAM_OK;
}

int b200slam_map_download_occupancy(b200slam_ctx *ctx, b200slam_map *map, int32_t *out, int stride)
{
    if (!ctx || !map || !out || stride < map->cols) return B200SLAM_ERR_ARG;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(out, sizeof(int32_t) * (size_t)stride, map->d_occ,
                                    sizeof(int32_t) * (size_t)map->occ_pitch, sizeof(int32_t) * (size_t)map->cols,
                                    map->rows, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200SLAM_OK;
}

int b200slam_map_upload_field(b200slam_ctx *ctx, b200slam_map *map, const float *field, int stride)
{
    if (!ctx || !map || !field || stride < map->cols) return B200SLAM_ERR_ARG;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(map->d_field, sizeof(float) * (size_t)map->field_pitch, field,
                                    sizeof(float) * (size_t)stride, sizeof(float) * (size_t)map->cols,
                                    map->rows, cudaMemcpyHostToDevice, ctx->stream));
    return B200SLAM_OK;
}

int b200slam_map_device_ptrs(b200slam_map *map, int32_t **occ, int *occ_pitch, float **field,
                             int *field_pitch)
{
    if (!map) return B200SLAM_ERR_ARG;
    if (occ) { *occ = map->d_occ; map->occ_exposed = true; map->raster_cells_n = -1; map->occ8_valid = false; }
    if (occ_pitch) *occ_pitch = map->occ_pitch;
    if (field) *field = map->d_field;
    if (field_pitch) *field_pitch = map->field_pitch;
    return B200SLAM_OK;
}

int b200slam_edt(b200slam_ctx *ctx, const int32_t *occ, int occ_stride, float *out, int out_stride,
                 int rows, int cols, float max_dist)
{
    if (!ctx || !occ || !out) return B200SLAM_ERR_ARG;
    if (rows <= 0 || cols <= 0) return B200SLAM_OK;       // empty grid: nothing to write
    if (occ_stride < cols || out_stride < cols)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "stride smaller than cols");
    b200slam_map *m = ctx->edt_map;
    if (!m || m->rows != rows || m->cols != cols) {
        if (m) b200slam_map_destroy(ctx, m);
        ctx->edt_map = nullptr;
        int rc = b200slam_map_create(ctx, rows, cols, &m);
        if (rc) return rc;
        ctx->edt_map = m;
    }
    int rc = b200slam_map_upload_occupancy(ctx, m, occ, occ_stride);
    if (rc) return rc;
    rc = b200slam_map_edt(ctx, m, max_dist);
    if (rc) return rc;
    return b200slam_map_download_field(ctx, m, out, out_stride);
}

/* ---- scan + lattice ------------------------------------------------------------ */

}  // extern "C"

int ensure_scan_capacity(b200slam_ctx *ctx, int nbeams)
{
    if (nbeams <= ctx->scan_cap && ctx->d_scan_x) return B200SLAM_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_scan_x); cudaFree(ctx->d_hit_values); cudaFreeHost(ctx->h_scan); cudaFreeHost(ctx->h_hit_values);
    ctx->d_scan_x = ctx->d_scan_y = ctx->d_hit_values = ctx->h_scan = ctx->h_hit_values = nullptr;
    ctx->scan_cap = 0;
    // at least the reference's bestHits[2500] (main.c:376), so that buffer is never re-allocated (and its
    // stale tail lost) for scans the reference itself could hold
    const int cap = ((nbeams > 2500 ? nbeams : 2500) + 255) & ~255;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_scan_x, sizeof(float) * 2 * cap));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_hit_values, sizeof(float) * 2 * cap));
    // FastMatchParameters.bestHits is a zero-initialised global that every match overwrites from
    // the front (main.c:374-378, 515); entries past the last candidate's count keep older values
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_hit_values, 0, sizeof(float) * 2 * cap, ctx->stream));
    CUDA_TRY(ctx, cudaHostAlloc(&ctx->h_scan, sizeof(float) * 2 * cap, cudaHostAllocDefault));
    CUDA_TRY(ctx, cudaHostAlloc(&ctx->h_hit_values, sizeof(float) * cap, cudaHostAllocDefault));
    if (!ctx->scan_event) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->scan_event, cudaEventDisableTiming));
    ctx->d_scan_y = ctx->d_scan_x + cap;
    ctx->scan_cap = cap;
    return B200SLAM_OK;
}

extern "C" {

int b200slam_scan_upload(b200slam_ctx *ctx, const float *x, const float *y, int nbeams)
{
    if (!ctx || nbeams < 0 || (nbeams > 0 && (!x || !y))) return B200SLAM_ERR_ARG;
    int rc = ensure_scan_capacity(ctx, nbeams);
    if (rc) return rc;
    ctx->scan_t_valid = false;
    ctx->scan_n_dev = false;              // the host's count is exact again
    if (nbeams > 0) {
        // one pinned, truly asynchronous copy of x | y (callers hand in pageable arrays)
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->scan_event));
        memcpy(ctx->h_scan, x, sizeof(float) * nbeams);
        memcpy(ctx->h_scan + ctx->scan_cap, y, sizeof(float) * nbeams);
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_scan_x, ctx->h_scan, sizeof(float) * ((size_t)ctx->scan_cap + nbeams),
                                      cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->scan_event, ctx->stream));
    }
    ctx->nbeams = nbeams;
    return B200SLAM_OK;
}

float b200slam_lattice_value(float p, float s, int k, int n)
{
    volatile float off = (float)(k - n / 2) * s;     // product rounded on its own
    return p + off;
}

}  // extern "C"

namespace {

bool is_capturing(b200slam_ctx *ctx)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(ctx->stream, &st);
    return st != cudaStreamCaptureStatusNone;
}

// Computes the lattice axis tables [ct | st | sxt | syt] with the host libm and the
// reference's float operations (main.c:424-437).  Small lattices (<= 960 floats) hand them
// to the kernel as launch parameters: nothing is copied on the stream.  Larger ones go
// through a ring of LAT_SLOTS pinned slots, each guarded by an event recorded after the
// kernel that reads its device copy, so back-to-back matches never wait on the host.
int stage_lattice(b200slam_ctx *ctx, const b200slam_map *map, const float pose0[3], const float step[3],
                  const int n[3], int64_t row_begin, int64_t row_end, LatticeLaunch *L)
{
    const int nth_all = n[0], ntx = n[1], nty = n[2];
    // only the angles this launch's (theta, tx) rows touch (a shard of a big lattice stays small)
    const int th_first = row_end > row_begin ? (int)(row_begin / ntx) : 0;
    const int th_last = row_end > row_begin ? (int)((row_end - 1) / ntx) : 0;
    const int nth = th_last - th_first + 1;
    const size_t need = (size_t)2 * nth + ntx + nty;
    const bool capturing = is_capturing(ctx);
    const bool by_param = need <= LATTICE_PARAM_FLOATS;
    static_assert(sizeof(ctx->h_param_tab) / sizeof(float) >= LATTICE_PARAM_FLOATS, "parameter-table scratch");
    if (!by_param && need > ctx->lat_cap) {
        if (capturing)
            return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "lattice larger than warmed-up scratch during graph capture");
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFreeHost(ctx->h_lat); cudaFree(ctx->d_lat);
        ctx->h_lat = ctx->d_lat = nullptr;
        ctx->lat_cap = 0;
        const size_t cap = (need + 1023) & ~(size_t)1023;
        CUDA_TRY(ctx, cudaHostAlloc(&ctx->h_lat, sizeof(float) * cap * LAT_SLOTS, cudaHostAllocDefault));
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_lat, sizeof(float) * cap * LAT_SLOTS));
        for (int i = 0; i < LAT_SLOTS; ++i)
            if (!ctx->lat_event[i])
                CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->lat_event[i], cudaEventDisableTiming));
        ctx->lat_cap = cap;
    }
    int slot = 0;
    if (!by_param) {
        slot = ctx->lat_next;
        ctx->lat_next = (slot + 1) % LAT_SLOTS;
        ctx->lat_cur = slot;
        if (!capturing) CUDA_TRY(ctx, cudaEventSynchronize(ctx->lat_event[slot]));
    }
    const float ipixel = 1 / map->pixel_size;                             // main.c:383
    // By-parameter tables are copied into the launch at cudaLaunchKernelEx time, so plain host scratch will
    // do; the pinned ring slots are only ever written here after waiting for the event behind the DMA that
    // reads them (a small lattice queued behind a large one must not touch slot 0).
    float *ct = by_param ? ctx->h_param_tab : ctx->h_lat + (size_t)slot * ctx->lat_cap, *st = ct + nth, *sxt = st + nth,
          *syt = sxt + ntx;
    for (int i = 0; i < nth; ++i) {
        const float th = b200slam_lattice_value(pose0[2], step[2], th_first + i, nth_all);   // main.c:424
        ct[i] = cosf(th);                                                 // main.c:434
        st[i] = sinf(th);                                                 // main.c:435
    }
    for (int i = 0; i < ntx; ++i) {
        const float tx = b200slam_lattice_value(pose0[0], step[0], i, ntx);   // main.c:425
        volatile float d = tx - map->top_left_x;
        sxt[i] = d * ipixel;                                              // main.c:436
    }
    for (int i = 0; i < nty; ++i) {
        const float ty = b200slam_lattice_value(pose0[1], step[1], i, nty);   // main.c:426
        volatile float d = ty - map->top_left_y;
        syt[i] = d * ipixel;                                              // main.c:437
    }
    L->map = map;
    L->nth = nth_all; L->ntx = ntx; L->nty = nty;
    L->th_first = th_first; L->nth_tab = nth;
    L->h_tables = ct;
    L->d_tables = nullptr;
    if (!by_param) {
        float *dst = ctx->d_lat + (size_t)slot * ctx->lat_cap;
        CUDA_TRY(ctx, cudaMemcpyAsync(dst, ct, sizeof(float) * need, cudaMemcpyHostToDevice, ctx->stream));
        L->d_tables = dst;
    }
    return B200SLAM_OK;
}

int check_lattice_args(b200slam_ctx *ctx, const b200slam_map *map, const float *pose0, const float *step,
                       const int *n)
{
    if (!ctx || !map || !pose0 || !step || !n) return B200SLAM_ERR_ARG;
    if (!map->has_geometry) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "map geometry not set");
    if (ctx->nbeams < 0 || !ctx->d_scan_x) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no scan uploaded");
    if (n[0] <= 0 || n[1] <= 0 || n[2] <= 0 || (long long)n[0] * n[1] * n[2] > 0xffffffffll)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "lattice %d x %d x %d unsupported", n[0], n[1], n[2]);
    return B200SLAM_OK;
}

// allreduce: 0 = this rank only; 1 = exchange the per-rank results and merge; 2 = post this
// rank's result to its peers but leave the merge to a later b200slam_exchange_collect_async.
int queue_lattice(b200slam_ctx *ctx, b200slam_map *map, const float pose0[3], const float step[3],
                  const int n[3], int64_t row_begin, int64_t row_end, bool want_scores, int allreduce)
{
    int rc = check_lattice_args(ctx, map, pose0, step, n);
    if (rc) return rc;
    const int64_t nrows = (int64_t)n[0] * n[1];
    if (row_begin < 0 || row_end > nrows || row_begin > row_end)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "row range [%lld,%lld) outside [0,%lld)",
                                  (long long)row_begin, (long long)row_end, (long long)nrows);
    LatticeLaunch L;
    rc = stage_lattice(ctx, map, pose0, step, n, row_begin, row_end, &L);
    if (rc) return rc;
    L.row_begin = row_begin;
    L.row_end = row_end;
    L.d_scores = nullptr;
    const bool multi = allreduce != 0 && ctx->nccl_comm && ctx->nranks > 1;
    L.exchange = multi && ctx->p2p_ready;      // posted from the kernel's tail over NVLink peer memory
    L.collect_prev = L.exchange && allreduce == 2;
    L.post_deferred = L.exchange && allreduce == 3;
    if (L.post_deferred) {                                    // recorded only: the collect kernel posts the burst
        if (ctx->posted_uncollected >= XCHG_MAX_POSTED)
            return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "more than %d posted matches without "
                                      "b200slam_exchange_collect_async", XCHG_MAX_POSTED);
        ctx->posted_uncollected++;
    }
    if (want_scores) {
        const size_t need = (size_t)nrows * n[2];
        if (need > ctx->scores_cap) {
            if (is_capturing(ctx))
                return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "score table allocation during graph capture");
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->d_scores);
            ctx->d_scores = nullptr;
            ctx->scores_cap = 0;
            CUDA_TRY(ctx, cudaMalloc(&ctx->d_scores, sizeof(float) * need));
            ctx->scores_cap = need;
        }
        L.d_scores = ctx->d_scores;
    }
    rc = lattice_launch(ctx, L);
    if (rc) return rc;
    bool gathered = false;
    if (multi && !L.exchange) {
        // {key, best_hits | last_hits} of every rank; b200slam_match_fetch merges them
        rc = comm_allgather_u64(ctx, &ctx->d_match->key, ctx->d_keys, 2);
        if (rc) return rc;
        gathered = true;
    }
    if (L.exchange && allreduce == 1) {
        rc = exchange_collect_launch(ctx);
        if (rc) return rc;
    }
    if (L.d_tables && !is_capturing(ctx)) CUDA_TRY(ctx, cudaEventRecord(ctx->lat_event[ctx->lat_cur], ctx->stream));
    ctx->last.valid = true;
    ctx->last.is_poses = false;
    ctx->last.gathered = gathered;
    ctx->last.exchanged = L.exchange;
    for (int i = 0; i < 3; ++i) {
        ctx->last.n[i] = n[i];
        ctx->last.pose0[i] = pose0[i];
        ctx->last.step[i] = step[i];
    }
    return B200SLAM_OK;
}

}  // namespace

extern "C" {

int b200slam_match_fetch(b200slam_ctx *ctx, b200slam_match *result)
{
    if (!ctx || !result) return B200SLAM_ERR_ARG;
    if (!ctx->last.valid) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no match queued");
    MatchDev m;
    m.error = 0;
    if (ctx->last.gathered) {
        // merge the all-gathered per-rank results: lowest key wins; the last candidate of the
        // whole lattice belongs to the last rank that scored anything
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_keys, ctx->d_keys, 16 * (size_t)ctx->nranks, cudaMemcpyDeviceToHost,
                                      ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        m.key = ~0ull; m.best_hits = 0; m.last_hits = 0;
        for (int r = 0; r < ctx->nranks; ++r) {
            const unsigned long long k = ctx->h_keys[2 * r], h = ctx->h_keys[2 * r + 1];
            if (k == ~0ull) continue;
            if (k < m.key) { m.key = k; m.best_hits = (int)(h & 0xffffffffull); }
            m.last_hits = (int)(h >> 32);
        }
    } else {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_match, ctx->d_match, sizeof(MatchDev), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        m = *ctx->h_match;
        if (ctx->last.exchanged && !ctx->last.is_poses) {     // the merged, global result
            m.key = m.gkey; m.best_hits = m.gbest_hits; m.last_hits = m.glast_hits;
        }
    }
    memset(result, 0, sizeof(*result));
    if (m.error) return device_error_check(ctx, m.error);
    if (m.key == ~0ull) {               // empty shard
        result->best_index = -1;
        result->best_score = INFINITY;
        return B200SLAM_OK;
    }
    const uint32_t bits = (uint32_t)(m.key >> 32);
    memcpy(&result->best_score, &bits, 4);
    result->best_index = (int64_t)(m.key & 0xffffffffull);
    result->best_hits = m.best_hits;
    result->last_hits = m.last_hits;
    if (!ctx->last.is_poses) {
        const int ntx = ctx->last.n[1], nty = ctx->last.n[2];
        const int64_t lin = result->best_index;
        const int ity = (int)(lin % nty);
        const int itx = (int)((lin / nty) % ntx);
        const int ith = (int)(lin / nty / ntx);
        result->best_pose[0] = b200slam_lattice_value(ctx->last.pose0[0], ctx->last.step[0], itx, ntx);
        result->best_pose[1] = b200slam_lattice_value(ctx->last.pose0[1], ctx->last.step[1], ity, nty);
        result->best_pose[2] = b200slam_lattice_value(ctx->last.pose0[2], ctx->last.step[2], ith, ctx->last.n[0]);
    } else if (ctx->last_poses_host) {
        const int64_t local = result->best_index - ctx->last_index_base;
        if (local >= 0 && local < ctx->last_P)
            for (int i = 0; i < 3; ++i) result->best_pose[i] = ctx->last_poses_host[3 * local + i];
    }
    return B200SLAM_OK;
}

int b200slam_match_fetch_hits(b200slam_ctx *ctx, float *hits, int count)
{
    if (!ctx || count < 0 || (count > 0 && !hits)) return B200SLAM_ERR_ARG;
    if (!ctx->d_hit_values) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no scan uploaded");
    if (count > ctx->scan_cap) count = ctx->scan_cap;
    if (count > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(hits, ctx->d_hit_values + ctx->scan_cap, sizeof(float) * (size_t)count,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200SLAM_OK;
}

int b200slam_score_lattice_async(b200slam_ctx *ctx, b200slam_map *map, const float pose0[3],
                                 const float step[3], const int n[3], int64_t row_begin, int64_t row_end,
                                 int allreduce)
{
    return queue_lattice(ctx, map, pose0, step, n, row_begin, row_end, false, allreduce);
}

int b200slam_exchange_collect_async(b200slam_ctx *ctx)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (!(ctx->nccl_comm && ctx->nranks > 1 && ctx->p2p_ready)) return B200SLAM_OK;   // nothing was deferred
    ctx->posted_uncollected = 0;
    return exchange_collect_launch(ctx);
}

int b200slam_score_lattice_rows(b200slam_ctx *ctx, b200slam_map *map, const float pose0[3],
                                const float step[3], const int n[3], int64_t row_begin, int64_t row_end,
                                int allreduce, b200slam_match *result)
{
    int rc = queue_lattice(ctx, map, pose0, step, n, row_begin, row_end, false, allreduce != 0 ? 1 : 0);
    if (rc) return rc;
    return b200slam_match_fetch(ctx, result);
}

int b200slam_score_lattice(b200slam_ctx *ctx, b200slam_map *map, const float pose0[3],
                           const float step[3], const int n[3], float *scores, float *last_hit_values,
                           b200slam_match *result)
{
    if (!n) return B200SLAM_ERR_ARG;
    const int64_t nrows = (int64_t)n[0] * n[1];
    int rc = queue_lattice(ctx, map, pose0, step, n, 0, nrows, scores != nullptr, 0);
    if (rc) return rc;
    if (scores)
        CUDA_TRY(ctx, cudaMemcpyAsync(scores, ctx->d_scores, sizeof(float) * (size_t)nrows * n[2],
                                      cudaMemcpyDeviceToHost, ctx->stream));
    // the last candidate's hit values ride along with the result: one synchronisation per match
    if (last_hit_values && ctx->nbeams > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_hit_values, ctx->d_hit_values + ctx->scan_cap,
                                      sizeof(float) * (size_t)ctx->nbeams, cudaMemcpyDeviceToHost, ctx->stream));
    b200slam_match tmp;
    rc = b200slam_match_fetch(ctx, result ? result : &tmp);       // copies the result block and synchronises
    if (rc) return rc;
    const b200slam_match *m = result ? result : &tmp;
    // FastMatch-sized lattices rewrite more than the last candidate's entries (main.c:515, see MatchDev)
    int written = m->last_hits;
    if (!ctx->last.gathered && ctx->h_match->written_hits > written && ctx->h_match->written_hits <= ctx->nbeams)
        written = ctx->h_match->written_hits;
    if (last_hit_values && written > 0)
        memcpy(last_hit_values, ctx->h_hit_values, sizeof(float) * (size_t)written);
    return B200SLAM_OK;
}

int b200slam_fastmatch(b200slam_ctx *ctx, b200slam_map *map, const float pose[3],
                       const float search_resolution[3], float pose_out[3], float *best_hits,
                       int *best_hits_size)
{
    if (!pose || !search_resolution || !pose_out) return B200SLAM_ERR_ARG;
    // main.c:386-387: t = searchResolution[0] for both translations, r = searchResolution[2].
    // The five sweeps of the while loop (main.c:440-591) re-score the same 27 candidates
    // (the lattice is built once, :422-438, and the refinement at :577-580 is commented
    // out), so one sweep gives the identical FastMatchParameters.
    const float step[3] = {search_resolution[0], search_resolution[0], search_resolution[2]};
    const int n[3] = {3, 3, 3};
    b200slam_match m;
    int rc = b200slam_score_lattice(ctx, map, pose, step, n, nullptr, best_hits, &m);
    if (rc) return rc;
    pose_out[0] = m.best_pose[0];                                         // main.c:592-594
    pose_out[1] = m.best_pose[1];
    pose_out[2] = m.best_pose[2];
    if (best_hits_size) *best_hits_size = m.best_hits;                    // main.c:557
    return B200SLAM_OK;
}

/* ---- FastMatch + FastMatch2 without a host round trip in between ------------------------ */

// FastMatch2 starts from FastMatch's result (main.c:918), which is one of 3 x 3 x 3 lattice points: the axis
// tables of every possible centre, computed exactly as stage_lattice would from the fetched pose (main.c:424-437:
// the centre is itself a lattice value of the first match).  t: ct[3][3] | st[3][3] | sxt[3][3] | syt[3][3].
static void seeded_tables(const b200slam_map *map_b, const float pose[3], const float step_a[3], const float step_b[3],
                          float t[36])
{
    const float ipixel = 1 / map_b->pixel_size;                          // main.c:383
    for (int s1 = 0; s1 < 3; ++s1) {
        const float th1 = b200slam_lattice_value(pose[2], step_a[2], s1, 3);
        const float tx1 = b200slam_lattice_value(pose[0], step_a[0], s1, 3);
        const float ty1 = b200slam_lattice_value(pose[1], step_a[1], s1, 3);
        for (int k = 0; k < 3; ++k) {
            const float th = b200slam_lattice_value(th1, step_b[2], k, 3);        // main.c:424
            t[3 * s1 + k] = cosf(th);                                             // main.c:434
            t[9 + 3 * s1 + k] = sinf(th);                                         // main.c:435
            const float tx = b200slam_lattice_value(tx1, step_b[0], k, 3);        // main.c:425
            volatile float dx = tx - map_b->top_left_x;
            t[18 + 3 * s1 + k] = dx * ipixel;                                     // main.c:436
            const float ty = b200slam_lattice_value(ty1, step_b[1], k, 3);        // main.c:426
            volatile float dy = ty - map_b->top_left_y;
            t[27 + 3 * s1 + k] = dy * ipixel;                                     // main.c:437
        }
    }
}

static int pair_bookkeeping(b200slam_ctx *ctx, const float pose[3], const float step_a[3], const float step_b[3])
{
    ctx->last.valid = true; ctx->last.is_poses = false; ctx->last.gathered = false; ctx->last.exchanged = false;
    for (int i = 0; i < 3; ++i) {
        ctx->last.n[i] = 3;
        ctx->last.step[i] = step_b[i];
        ctx->pair.guess[i] = pose[i]; ctx->pair.step_a[i] = step_a[i]; ctx->pair.step_b[i] = step_b[i];
    }
    ctx->pair.valid = true;
    return B200SLAM_OK;
}

static int ensure_host_result(b200slam_ctx *ctx)
{
    if (!ctx->h_result) {
        CUDA_TRY(ctx, cudaHostAlloc(&ctx->h_result, sizeof(MatchHost), cudaHostAllocMapped));
        memset(ctx->h_result, 0, sizeof(MatchHost));
    }
    return B200SLAM_OK;
}

int b200slam_fastmatch_pair_async(b200slam_ctx *ctx, b200slam_map *map_a, b200slam_map *map_b, const float pose[3],
                                  const float res_a[3], const float res_b[3])
{
    if (!ctx || !map_a || !map_b || !pose || !res_a || !res_b) return B200SLAM_ERR_ARG;
    const float step_a[3] = {res_a[0], res_a[0], res_a[2]};               // main.c:386-387
    const float step_b[3] = {res_b[0], res_b[0], res_b[2]};
    const int n[3] = {3, 3, 3};
    int rc = queue_lattice(ctx, map_a, pose, step_a, n, 0, 9, false, 0);  // FastMatch (main.c:902 / :909)
    if (rc) return rc;
    rc = check_lattice_args(ctx, map_b, pose, step_b, n);
    if (rc) return rc;
    float *t = ctx->h_param_tab;
    seeded_tables(map_b, pose, step_a, step_b, t);
    LatticeLaunch L;
    L.map = map_b;
    L.nth = L.ntx = L.nty = 3;
    L.th_first = 0; L.nth_tab = 3;
    L.h_tables = t; L.d_tables = nullptr;
    L.row_begin = 0; L.row_end = 9;
    L.d_scores = nullptr;
    L.exchange = L.collect_prev = L.post_deferred = false;
    L.seeded = true;
    rc = ensure_host_result(ctx);
    if (rc) return rc;
    L.host_result = true;
    ctx->result_seq++;
    rc = lattice_launch(ctx, L);
    if (rc) return rc;
    return pair_bookkeeping(ctx, pose, step_a, step_b);
}

// readAScan + FastMatch + FastMatch2 as ONE kernel.  d_src: the scan's ranges on the device.
static int scan_step_queue(b200slam_ctx *ctx, const float *d_src, int max_range, b200slam_map *map_a, b200slam_map *map_b,
                           const float pose[3], const float res_a[3], const float res_b[3])
{
    const float step_a[3] = {res_a[0], res_a[0], res_a[2]};               // main.c:386-387
    const float step_b[3] = {res_b[0], res_b[0], res_b[2]};
    const int n[3] = {3, 3, 3};
    if (!map_a->has_geometry || !map_b->has_geometry) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "map geometry not set");
    LatticeLaunch L;
    int rc = stage_lattice(ctx, map_a, pose, step_a, n, 0, 9, &L);        // 12 floats in ctx->h_param_tab
    if (rc) return rc;
    float t12[12], t36[36];
    memcpy(t12, L.h_tables, sizeof t12);
    seeded_tables(map_b, pose, step_a, step_b, t36);
    rc = ensure_host_result(ctx);
    if (rc) return rc;
    ctx->result_seq++;
    rc = scan_step_launch(ctx, map_a, map_b, t12, t36, d_src, max_range);
    if (rc) return rc;
    ctx->nbeams = ctx->lidar_n;            // upper bound until the fetch
    ctx->scan_n_dev = true;
    ctx->scan_t_valid = false;
    return pair_bookkeeping(ctx, pose, step_a, step_b);
}

int b200slam_scan_step_async(b200slam_ctx *ctx, const float *ranges, int max_range, b200slam_map *map_a, b200slam_map *map_b,
                             const float pose[3], const float res_a[3], const float res_b[3])
{
    if (!ctx || !ranges || !map_a || !map_b || !pose || !res_a || !res_b) return B200SLAM_ERR_ARG;
    if (ctx->lidar_n <= 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_lidar_set first");
    int rc = ensure_scan_capacity(ctx, ctx->lidar_n);
    if (!rc) rc = ensure_front(ctx);
    if (rc) return rc;
    const int nl = ctx->lidar_n;
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->scan_event));                  // pinned staging free again
    memcpy(ctx->h_ranges, ranges, sizeof(float) * nl);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_ranges, ctx->h_ranges, sizeof(float) * nl, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->scan_event, ctx->stream));
    return scan_step_queue(ctx, ctx->d_ranges, max_range, map_a, map_b, pose, res_a, res_b);
}

int b200slam_scan_step_resident_async(b200slam_ctx *ctx, int64_t first_value, int max_range, b200slam_map *map_a,
                                      b200slam_map *map_b, const float pose[3], const float res_a[3], const float res_b[3])
{
    if (!ctx || first_value < 0 || !map_a || !map_b || !pose || !res_a || !res_b) return B200SLAM_ERR_ARG;
    if (ctx->lidar_n <= 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_lidar_set first");
    if (!ctx->d_csv_values || first_value + ctx->lidar_n > ctx->csv_count)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "values [%lld, %lld) are not in the ingested CSV (%lld values)",
                                  (long long)first_value, (long long)(first_value + ctx->lidar_n), (long long)ctx->csv_count);
    int rc = ensure_scan_capacity(ctx, ctx->lidar_n);
    if (!rc) rc = ensure_front(ctx);
    if (rc) return rc;
    return scan_step_queue(ctx, ctx->d_csv_values + first_value, max_range, map_a, map_b, pose, res_a, res_b);
}

int b200slam_fastmatch_pair_fetch(b200slam_ctx *ctx, float pose_a[3], float pose_b[3], int *scan_size, int *best_hits_size)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (!ctx->pair.valid) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no b200slam_fastmatch_pair_async queued");
    // The second kernel's tail writes the result block into mapped host memory, seq last: watch it arrive (no
    // copy queued, no driver call).  Bounded: after two seconds fall back to a stream synchronisation, which
    // also surfaces a launch failure.
    volatile MatchHost *h = ctx->h_result;
    {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        unsigned spins = 0;
        while (h->seq != ctx->result_seq) {
            if ((++spins & 0xfffu) == 0) {
                clock_gettime(CLOCK_MONOTONIC, &t1);
                if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > 2.0) {
                    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
                    if (h->seq != ctx->result_seq)
                        return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "pair match finished without publishing its result");
                }
            }
        }
        __sync_synchronize();
    }
    MatchHost m;
    m.key = h->key; m.seed_key = h->seed_key; m.best_hits = h->best_hits; m.error = h->error;
    if (getenv("B200SLAM_FM_TRACE")) {              // diagnostics: SM cycles between the fused kernel's phase boundaries
        static int printed = 0;
        if (printed++ % 500 == 100) {
            fprintf(stderr, "[b200slam] fastmatch kernel phases (cycles):");
            for (int i = 1; i < 16 && h->trace[i]; ++i) fprintf(stderr, " %lld", h->trace[i] - h->trace[i - 1]);
            fprintf(stderr, "\n");
        }
    }
    if (m.error) return device_error_check(ctx, m.error);
    if (ctx->scan_n_dev) ctx->nbeams = h->scan_n;
    if (ctx->mp_n_dev) ctx->mp_size = h->mp_n;
    if (m.key == ~0ull || m.seed_key == ~0ull) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "pair match scored nothing");
    const int l1 = (int)(m.seed_key & 0xffffffffull), l2 = (int)(m.key & 0xffffffffull);
    float pa[3];
    pa[0] = b200slam_lattice_value(ctx->pair.guess[0], ctx->pair.step_a[0], (l1 / 3) % 3, 3);   // main.c:592-594
    pa[1] = b200slam_lattice_value(ctx->pair.guess[1], ctx->pair.step_a[1], l1 % 3, 3);
    pa[2] = b200slam_lattice_value(ctx->pair.guess[2], ctx->pair.step_a[2], l1 / 9, 3);
    if (pose_a) for (int i = 0; i < 3; ++i) pose_a[i] = pa[i];
    if (pose_b) {
        pose_b[0] = b200slam_lattice_value(pa[0], ctx->pair.step_b[0], (l2 / 3) % 3, 3);
        pose_b[1] = b200slam_lattice_value(pa[1], ctx->pair.step_b[1], l2 % 3, 3);
        pose_b[2] = b200slam_lattice_value(pa[2], ctx->pair.step_b[2], l2 / 9, 3);
    }
    for (int i = 0; i < 3; ++i) ctx->last.pose0[i] = pa[i];     // b200slam_match_fetch then describes the second match
    if (scan_size) *scan_size = ctx->nbeams;
    if (best_hits_size) *best_hits_size = m.best_hits;           // main.c:557
    return B200SLAM_OK;
}

/* ---- the per-scan loop on the device: scans queued ahead of their results (main.c:859-970) ------------- */

}  // extern "C"

namespace {
__global__ void chain_set_kernel(ChainDev *dst, const ChainDev v) { *dst = v; }
}  // namespace

extern "C" {

int b200slam_scan_chain_begin(b200slam_ctx *ctx, int scan_index, const float pose[3], const float *prev_pose,
                              const float map_pose[3], float mini_update_dt, float mini_update_dr)
{
    if (!ctx || !pose || !map_pose || scan_index < 0) return B200SLAM_ERR_ARG;
    int rc = chain_trig_selftest(ctx);                    // once per context: the device's cosf / sinf == this host's libm
    if (rc) return rc;
    if (!ctx->d_chain) CUDA_TRY(ctx, cudaMalloc(&ctx->d_chain, sizeof(ChainDev)));
    if (!ctx->h_chain_ring) {
        CUDA_TRY(ctx, cudaHostAlloc(&ctx->h_chain_ring, sizeof(ChainSlot) * CHAIN_RING, cudaHostAllocMapped));
        memset(ctx->h_chain_ring, 0, sizeof(ChainSlot) * CHAIN_RING);
    }
    ChainDev v;
    for (int i = 0; i < 3; ++i) {
        v.pose[i] = pose[i];
        v.prev[i] = prev_pose ? prev_pose[i] : pose[i];
        v.map_pose[i] = map_pose[i];
    }
    v.have_prev = prev_pose ? 1 : 0;
    v.next_scan = scan_index;
    v.stop = 0;
    ctx->chain_mini_dt = mini_update_dt;
    ctx->chain_mini_dr = mini_update_dr;
    chain_set_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_chain, v);
    LAUNCH_CHECK(ctx);
    return B200SLAM_OK;
}

int b200slam_scan_chain_step_async(b200slam_ctx *ctx, int scan_index, int nscans, int64_t first_value, int max_range,
                                   b200slam_map *map_a, b200slam_map *map_b, const float res_a[3], const float res_b[3])
{
    if (!ctx || scan_index < 0 || nscans < 1 || nscans > CHAIN_RING / 2 || first_value < 0 || !map_a || !map_b || !res_a || !res_b)
        return B200SLAM_ERR_ARG;
    if (!ctx->d_chain) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_scan_chain_begin first");
    if (ctx->lidar_n <= 0) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_lidar_set first");
    if (!ctx->d_csv_values || first_value + (int64_t)nscans * ctx->lidar_n > ctx->csv_count)
        return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "values [%lld, %lld) are not in the ingested CSV (%lld values)",
                                  (long long)first_value, (long long)(first_value + (int64_t)nscans * ctx->lidar_n), (long long)ctx->csv_count);
    if (!map_a->has_geometry || !map_b->has_geometry) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "map geometry not set");
    int rc = ensure_scan_capacity(ctx, ctx->lidar_n);
    if (!rc) rc = ensure_front(ctx);
    if (rc) return rc;
    ChainLaunch C;
    C.scan_index = scan_index;
    C.count = nscans;
    C.step_a[0] = C.step_a[1] = res_a[0]; C.step_a[2] = res_a[2];         // main.c:386-387
    C.step_b[0] = C.step_b[1] = res_b[0]; C.step_b[2] = res_b[2];
    rc = scan_chain_launch(ctx, map_a, map_b, C, ctx->d_csv_values + first_value, max_range);
    if (rc) return rc;
    ctx->nbeams = ctx->lidar_n;            // upper bound until a fetch
    ctx->scan_n_dev = true;
    ctx->scan_t_valid = false;
    ctx->last.valid = true; ctx->last.is_poses = false; ctx->last.gathered = false; ctx->last.exchanged = false;
    for (int i = 0; i < 3; ++i) { ctx->last.n[i] = 3; ctx->last.step[i] = C.step_b[i]; }
    ctx->pair.valid = false;
    return B200SLAM_OK;
}

int b200slam_scan_chain_fetch(b200slam_ctx *ctx, int scan_index, float pose_a[3], float pose_b[3], int *scan_size,
                              int *best_hits_size, int *stopped)
{
    if (!ctx || scan_index < 0) return B200SLAM_ERR_ARG;
    if (!ctx->h_chain_ring) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no chained scan queued");
    volatile ChainSlot *h = ctx->h_chain_ring + (scan_index % CHAIN_RING);
    const unsigned long long want = (unsigned long long)scan_index + 1ull;
    {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        unsigned spins = 0;
        while (h->seq != want) {
            if ((++spins & 0xfffu) == 0) {
                clock_gettime(CLOCK_MONOTONIC, &t1);
                if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > 2.0) {
                    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
                    if (h->seq != want)
                        return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "chained scan %d did not run (the chain had stopped, or "
                                                  "was begun at another scan)", scan_index);
                }
            }
        }
        __sync_synchronize();
    }
    if (h->error) return device_error_check(ctx, h->error);
    if (pose_a) for (int i = 0; i < 3; ++i) pose_a[i] = h->pose_a[i];
    if (pose_b) for (int i = 0; i < 3; ++i) pose_b[i] = h->pose_b[i];
    for (int i = 0; i < 3; ++i) ctx->last.pose0[i] = h->pose_a[i];
    ctx->nbeams = h->scan_n;
    if (ctx->mp_n_dev) ctx->mp_size = h->mp_n;
    if (scan_size) *scan_size = h->scan_n;
    if (best_hits_size) *best_hits_size = h->best_hits;
    if (stopped) *stopped = h->stopped;
    return B200SLAM_OK;
}

/* ---- pose lists / particles ---------------------------------------------------- */

// Sizes the particle buffers for P poses and stages poses (+ cos/sin from the host libm when not
// given, main.c:434-435) into the device SoA x | y | ct | st | theta.
static int stage_poses(b200slam_ctx *ctx, const float *poses, const float *ct, const float *st, int64_t P)
{
    if ((size_t)P > ctx->pose_cap) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        particles_unshare_blocks(ctx);               // peers' mappings of the old block die with it
        cudaFree(ctx->d_pose_block); cudaFree(ctx->d_hits); cudaFreeHost(ctx->h_pose_stage);
        cudaFree(ctx->d_q); cudaFree(ctx->d_block_sums); cudaFree(ctx->d_weights);
        ctx->d_pose_block = nullptr; ctx->d_pose_soa = nullptr; ctx->d_pose_alt = nullptr; ctx->d_anc_resident = nullptr;
        ctx->d_hits = nullptr; ctx->h_pose_stage = nullptr;
        ctx->d_q = nullptr; ctx->d_block_sums = nullptr; ctx->d_weights = nullptr;
        ctx->pose_cap = 0;
        const size_t cap = ((size_t)P + 4095) & ~(size_t)4095;
        // [2][5 * cap] floats (the two pose buffers) | [cap] int32 ancestors of the resident set
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_pose_block, sizeof(float) * 11 * cap));
        ctx->d_pose_soa = ctx->d_pose_block;
        ctx->d_pose_alt = ctx->d_pose_block + 5 * cap;
        ctx->d_anc_resident = reinterpret_cast<int32_t *>(ctx->d_pose_block + 10 * cap);
        ctx->pose_parity = 0;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_hits, sizeof(int32_t) * cap));
        CUDA_TRY(ctx, cudaHostAlloc(&ctx->h_pose_stage, sizeof(float) * 5 * cap, cudaHostAllocDefault));
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_q, sizeof(unsigned long long) * cap));
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_block_sums, sizeof(unsigned long long) * (cap / 1024 + 2)));
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_weights, sizeof(float) * cap));
        ctx->pose_cap = cap;
    }
    if ((size_t)P > ctx->scores_cap) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_scores);
        ctx->d_scores = nullptr;
        ctx->scores_cap = 0;
        const size_t cap = ((size_t)P + 4095) & ~(size_t)4095;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_scores, sizeof(float) * cap));
        ctx->scores_cap = cap;
    }
    if ((size_t)P > ctx->anc_cap) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_ancestors);
        ctx->d_ancestors = nullptr;
        ctx->anc_cap = 0;
        const size_t cap = ((size_t)P + 4095) & ~(size_t)4095;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_ancestors, sizeof(int32_t) * cap));
        ctx->anc_cap = cap;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // staging buffer free again
    const size_t cap = ctx->pose_cap;
    float *hx = ctx->h_pose_stage, *hy = hx + cap, *hct = hy + cap, *hst = hct + cap, *hth = hst + cap;
    for (int64_t p = 0; p < P; ++p) {
        hx[p] = poses[3 * p + 0];
        hy[p] = poses[3 * p + 1];
        hth[p] = poses[3 * p + 2];
        hct[p] = ct ? ct[p] : cosf(poses[3 * p + 2]);                     // main.c:434
        hst[p] = st ? st[p] : sinf(poses[3 * p + 2]);                     // main.c:435
    }
    if (P > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pose_soa, ctx->h_pose_stage, sizeof(float) * 5 * cap,
                                      cudaMemcpyHostToDevice, ctx->stream));
    return B200SLAM_OK;
}

int b200slam_score_poses(b200slam_ctx *ctx, b200slam_map *map, const float *poses, const float *ct,
                         const float *st, int64_t P, int64_t index_base, float *scores, int32_t *hits,
                         b200slam_match *result)
{
    if (!ctx || !map || P < 0 || (P > 0 && !poses)) return B200SLAM_ERR_ARG;
    if (!map->has_geometry) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "map geometry not set");
    if (!ctx->d_scan_x) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no scan uploaded");
    if (P + index_base > 0xffffffffll) return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "too many poses");
    int rc = stage_poses(ctx, poses, ct, st, P);
    if (rc) return rc;
    rc = poses_launch(ctx, map, P, index_base, ctx->d_scores, ctx->d_hits);
    if (rc) return rc;
    ctx->last.valid = true;
    ctx->last.is_poses = true;
    ctx->last.gathered = false;
    ctx->last.exchanged = false;
    ctx->last_P = P;
    ctx->last_index_base = index_base;
    ctx->last_poses_host = poses;
    if (scores && P > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(scores, ctx->d_scores, sizeof(float) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    if (hits && P > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(hits, ctx->d_hits, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    if (result) {
        rc = b200slam_match_fetch(ctx, result);       // syncs; the kernel's last CTA published the hits
        if (rc) return rc;
    } else {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->last_poses_host = nullptr;      // caller's buffer is only borrowed during this call
    return B200SLAM_OK;
}

/* ---- device-resident particle set ------------------------------------------------------- */

int b200slam_particles_upload(b200slam_ctx *ctx, const float *poses, const float *ct, const float *st, int64_t P)
{
    if (!ctx || P <= 0 || !poses || P > 0x7fffffffll) return B200SLAM_ERR_ARG;
    int rc = stage_poses(ctx, poses, ct, st, P);
    if (rc) return rc;
    ctx->last_P = P;
    ctx->last_index_base = 0;
    ctx->last_poses_host = nullptr;
    ctx->last.valid = false;
    ctx->pf_sharded = false;
    return B200SLAM_OK;
}

int b200slam_particles_shard(b200slam_ctx *ctx, const float *poses, const float *ct, const float *st, int64_t P,
                             int64_t index_base, int64_t n_global)
{
    if (!ctx || P <= 0 || !poses || index_base < 0 || n_global < index_base + P || n_global > 0x7fffffffll)
        return B200SLAM_ERR_ARG;
    if (!ctx->nccl_comm || ctx->nranks < 2 || !ctx->p2p_ready)
        return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_particles_shard needs b200slam_comm_init with NVLink peer "
                                                            "memory (one GPU: b200slam_particles_upload)");
    int rc = stage_poses(ctx, poses, ct, st, P);
    if (rc) return rc;
    rc = particles_share_blocks(ctx);              // collective
    if (rc) return rc;
    ctx->last_P = P;
    ctx->last_index_base = index_base;
    ctx->last_poses_host = nullptr;
    ctx->last.valid = false;
    ctx->pf_sharded = true;
    ctx->pf_nglobal = n_global;
    // nobody pushes offspring into this rank's buffers before every rank has staged its slice
    return comm_peer_barrier(ctx);
}

int b200slam_particles_score_async(b200slam_ctx *ctx, b200slam_map *map)
{
    if (!ctx || !map) return B200SLAM_ERR_ARG;
    if (!map->has_geometry) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "map geometry not set");
    if (!ctx->d_scan_x) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no scan uploaded");
    if (ctx->last_P <= 0 || !ctx->d_pose_soa) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no particles uploaded");
    const int64_t base = ctx->pf_sharded ? ctx->last_index_base : 0;
    int rc = poses_launch(ctx, map, ctx->last_P, base, ctx->d_scores, ctx->d_hits, ctx->pf_sharded);
    if (rc) return rc;
    ctx->last.valid = true;
    ctx->last.is_poses = true;
    ctx->last.gathered = false;
    ctx->last.exchanged = false;
    ctx->last_index_base = base;
    return B200SLAM_OK;
}

int b200slam_particles_resample_async(b200slam_ctx *ctx, float beta, uint32_t u0_q32)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (!ctx->last.valid || !ctx->last.is_poses)
        return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_particles_score_async must run first");
    // the scores must be those of the resident set, not of a b200slam_score_poses call in between
    return particles_resample_resident(ctx, ctx->last_P, beta, u0_q32);
}

int b200slam_particles_download(b200slam_ctx *ctx, float *poses, float *scores, float *weights, int32_t *ancestors)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    const int64_t P = ctx->last_P;
    if (P <= 0 || !ctx->d_pose_soa) return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "no particles uploaded");
    const size_t cap = ctx->pose_cap;
    if (poses) {
        float *h = ctx->h_pose_stage;
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_pose_soa, sizeof(float) * 2 * cap, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(h + 4 * cap, ctx->d_pose_soa + 4 * cap, sizeof(float) * (size_t)P,
                                      cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        for (int64_t p = 0; p < P; ++p) {
            poses[3 * p + 0] = h[p];
            poses[3 * p + 1] = h[cap + p];
            poses[3 * p + 2] = h[4 * cap + p];
        }
    }
    if (scores) CUDA_TRY(ctx, cudaMemcpyAsync(scores, ctx->d_scores, sizeof(float) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    if (weights) CUDA_TRY(ctx, cudaMemcpyAsync(weights, ctx->d_weights, sizeof(float) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    if (ancestors) CUDA_TRY(ctx, cudaMemcpyAsync(ancestors, ctx->d_anc_resident, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->h_match->error, &ctx->d_match->error, sizeof(unsigned int), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return device_error_check(ctx, ctx->h_match->error);
}

int b200slam_weights_resample(b200slam_ctx *ctx, float beta, uint32_t u0_q32, float *weights,
                              uint64_t *wsum, int32_t *ancestors, int64_t *k_begin, int64_t *k_count)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (!ctx->last.valid || !ctx->last.is_poses)
        return b200slam_set_error(ctx, B200SLAM_ERR_STATE, "b200slam_score_poses must run first");
    return particles_weights_resample(ctx, ctx->last_P, beta, u0_q32, weights, wsum, ancestors, k_begin,
                                      k_count);
}

int b200slam_pyramid_match(b200slam_ctx *ctx, b200slam_map *const *maps, int levels,
                           const float pose0[3], const float *steps, const int *n,
                           b200slam_match *results)
{
    if (!ctx || !maps || levels <= 0 || !pose0 || !steps || !n || !results) return B200SLAM_ERR_ARG;
    // main.c:901-918: each level is seeded with the previous level's winner.
    float seed[3] = {pose0[0], pose0[1], pose0[2]};
    const bool multi = ctx->nccl_comm && ctx->nranks > 1;
    for (int l = 0; l < levels; ++l) {
        const int *nl = n + 3 * l;
        const int64_t nrows = (int64_t)nl[0] * nl[1];
        int64_t rb = 0, re = nrows;
        if (multi) b200slam_shard_range(nrows, ctx->nranks, ctx->rank, &rb, &re);
        int rc = b200slam_score_lattice_rows(ctx, maps[l], seed, steps + 3 * l, nl, rb, re, multi ? 1 : 0,
                                             &results[l]);
        if (rc) return rc;
        for (int i = 0; i < 3; ++i) seed[i] = results[l].best_pose[i];
    }
    return B200SLAM_OK;
}

/* ---- CUDA graphs: capture a launch-bound sequence once, replay it ------------------ */

struct b200slam_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    uint64_t kernels = 0;
};

int b200slam_graph_begin(b200slam_ctx *ctx)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    ctx->graph_launch_mark = ctx->launches;
    CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    return B200SLAM_OK;
}

int b200slam_graph_end(b200slam_ctx *ctx, b200slam_graph **out)
{
    if (!ctx || !out) return B200SLAM_ERR_ARG;
    *out = nullptr;
    cudaGraph_t g = nullptr;
    CUDA_TRY(ctx, cudaStreamEndCapture(ctx->stream, &g));
    b200slam_graph *G = new (std::nothrow) b200slam_graph();
    if (!G) { cudaGraphDestroy(g); return B200SLAM_ERR_NOMEM; }
    G->graph = g;
    G->kernels = ctx->launches - ctx->graph_launch_mark;
    ctx->launches = ctx->graph_launch_mark;      // captured launches have not run yet
    cudaError_t e = cudaGraphInstantiate(&G->exec, g, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(g);
        delete G;
        return b200slam_set_error(ctx, B200SLAM_ERR_CUDA, "cudaGraphInstantiate -> %s", cudaGetErrorString(e));
    }
    *out = G;
    return B200SLAM_OK;
}

int b200slam_graph_launch(b200slam_ctx *ctx, b200slam_graph *g)
{
    if (!ctx || !g) return B200SLAM_ERR_ARG;
    CUDA_TRY(ctx, cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->kernels;
    return B200SLAM_OK;
}

void b200slam_graph_destroy(b200slam_ctx *ctx, b200slam_graph *g)
{
    if (!g) return;
    if (ctx && ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
}

/* ---- device timing ------------------------------------------------------------------ */

int b200slam_event_record(b200slam_ctx *ctx, int slot)
{
    if (!ctx || slot < 0 || slot >= 4096) return B200SLAM_ERR_ARG;
    if (!ctx->timing[slot]) CUDA_TRY(ctx, cudaEventCreate(&ctx->timing[slot]));
    CUDA_TRY(ctx, cudaEventRecord(ctx->timing[slot], ctx->stream));
    return B200SLAM_OK;
}

int b200slam_event_wait(b200slam_ctx *ctx, b200slam_ctx *other, int slot)
{
    if (!ctx || !other || slot < 0 || slot >= 4096 || !other->timing[slot]) return B200SLAM_ERR_ARG;
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, other->timing[slot], 0));
    return B200SLAM_OK;
}

int b200slam_event_elapsed_ms(b200slam_ctx *ctx, int slot_start, int slot_stop, float *ms)
{
    if (!ctx || !ms || slot_start < 0 || slot_start >= 4096 || slot_stop < 0 || slot_stop >= 4096 ||
        !ctx->timing[slot_start] || !ctx->timing[slot_stop])
        return B200SLAM_ERR_ARG;
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->timing[slot_stop]));
    CUDA_TRY(ctx, cudaEventElapsedTime(ms, ctx->timing[slot_start], ctx->timing[slot_stop]));
    return B200SLAM_OK;
}

/* ---- sharding helpers (pure host) ----------------------------------------------- */

void b200slam_shard_range(int64_t total, int nranks, int rank, int64_t *begin, int64_t *end)
{
    if (nranks <= 0) nranks = 1;
    const int64_t base = total / nranks, rem = total % nranks;
    const int64_t b = rank * base + (rank < rem ? rank : rem);
    if (begin) *begin = b;
    if (end) *end = b + base + (rank < rem ? 1 : 0);
}

uint64_t b200slam_pack_key(float score, uint32_t index) { return pack_key(score, index); }

void b200slam_unpack_key(uint64_t key, float *score, uint32_t *index)
{
    const uint32_t bits = (uint32_t)(key >> 32);
    if (score) memcpy(score, &bits, 4);
    if (index) *index = (uint32_t)(key & 0xffffffffull);
}

uint64_t b200slam_merge_keys(const uint64_t *keys, int n)
{
    uint64_t best = ~0ull;
    for (int i = 0; i < n; ++i) best = keys[i] < best ? keys[i] : best;
    return best;
}

}  // extern "C"
