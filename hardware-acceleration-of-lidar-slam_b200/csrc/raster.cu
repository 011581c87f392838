// raster.cu -- map points -> occupancy grid on the device: the scatter loop of
// OccupationalGrid (Subsystem_1/main.c:332-353), one level.  The bounding box, margin and
// grid size (main.c:272-305) are computed by the caller on the host with the reference's
// float operations (api.cu: b200slam_map_rasterise); the per-point arithmetic here is
//     hits = (int)roundf((p - min) / PIXELSIZE) + 1          (IEEE division, roundf)
//     idx  = (hits_y - 1) * Sgrid_x + hits_x - 1 ;  row = idx / Sgrid_x ; col = idx % Sgrid_x
// exactly as the reference, including its index wrap when hits_x runs past the row.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
raster_clear_kernel(int4 *__restrict__ occ, long n4)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) occ[i] = make_int4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(256)
raster_erase_kernel(const int32_t *__restrict__ cells, int n, int32_t *__restrict__ occ, uint8_t *__restrict__ occ8)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int off = cells[a];
    if (off >= 0) {
        occ[off] = 0;
        if (occ8) occ8[off] = 0;
    }
}

__global__ void __launch_bounds__(256)
raster_scatter_kernel(const float *__restrict__ px, const float *__restrict__ py, int n, float min_x,
                      float min_y, float pixel, int sgrid_x, int rows, int32_t *__restrict__ occ, int pitch,
                      int32_t *__restrict__ cells, uint8_t *__restrict__ occ8)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const float dx = __fsub_rn(px[a], min_x);                               // main.c:333
    const float dy = __fsub_rn(py[a], min_y);                               // main.c:334
    const int hx = (int)roundf(__fdiv_rn(dx, pixel)) + 1;                   // main.c:339
    const int hy = (int)roundf(__fdiv_rn(dy, pixel)) + 1;                   // main.c:340
    const int idx = (((hy - 1) * sgrid_x) + hx) - 1;                        // main.c:345
    const int row = idx / sgrid_x, col = idx % sgrid_x;                     // main.c:348-349
    const bool in = row >= 0 && row < rows && col >= 0;
    const int off = in ? row * pitch + col : -1;                            // cells < 2^30 (b200slam_map_create)
    if (in) {
        occ[off] = 1;                                                       // main.c:355
        if (occ8) occ8[off] = 1;
    }
    if (cells) cells[a] = off;
}

}  // namespace

int rasterise_launch(b200slam_ctx *ctx, b200slam_map *map, int npoints, float min_x, float min_y, float pixel_size)
{
    // main.c:319 clears the whole array.  Here the occupancy is kept zero outside the cells the previous
    // rasterisation set, which are remembered, so clearing is erasing those (O(points), not O(cells)).
    int32_t *cells = nullptr;
    uint8_t *occ8 = map->occ_exposed ? nullptr : map->d_occ8;          // the byte shadow follows the list mode only
    if (!map->occ_exposed) {
        if ((size_t)npoints > map->raster_cells_cap) {
            const size_t cap = ((size_t)npoints + 4095) & ~(size_t)4095;
            int32_t *d = nullptr;
            CUDA_TRY(ctx, cudaMalloc(&d, sizeof(int32_t) * cap));
            if (map->raster_cells_n > 0) {                 // the old list still has to erase its cells
                CUDA_TRY(ctx, cudaMemcpyAsync(d, map->d_raster_cells, sizeof(int32_t) * map->raster_cells_n,
                                              cudaMemcpyDeviceToDevice, ctx->stream));
                CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            }
            cudaFree(map->d_raster_cells);
            map->d_raster_cells = d;
            map->raster_cells_cap = cap;
        }
        cells = map->d_raster_cells;
    }
    if (cells && map->raster_cells_n >= 0) {
        if (map->raster_cells_n > 0) {
            raster_erase_kernel<<<(map->raster_cells_n + 255) / 256, 256, 0, ctx->stream>>>(cells, map->raster_cells_n,
                                                                                          map->d_occ, occ8);
            LAUNCH_CHECK(ctx);
        }
    } else {
        const long n4 = (long)(cells ? map->cap_rows : map->rows) * map->occ_pitch / 4;
        raster_clear_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<int4 *>(map->d_occ), n4);
        LAUNCH_CHECK(ctx);
        if (occ8) CUDA_TRY(ctx, cudaMemsetAsync(occ8, 0, (size_t)map->cap_rows * map->occ_pitch, ctx->stream));
    }
    raster_scatter_kernel<<<(npoints + 255) / 256, 256, 0, ctx->stream>>>(
        ctx->d_points, ctx->d_points + ctx->points_cap, npoints, min_x, min_y, pixel_size, map->cols, map->rows,
        map->d_occ, map->occ_pitch, cells, occ8);
    LAUNCH_CHECK(ctx);
    map->raster_cells_n = cells ? npoints : -1;
    map->occ8_valid = occ8 != nullptr;
    map->occ8_rows = map->cap_rows; map->occ8_cols = map->cap_cols;
    return B200SLAM_OK;
}
