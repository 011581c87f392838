/*
 * slam_oracle.h -- CPU restatement of the reference's EDT + scan-matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker or the reported CPU baseline.  The
 * product (libb200slam.so) never links, loads or calls anything in this directory
 * and has no CPU fallback.
 *
 * Reference: circuitpotato/Hardware-Acceleration-of-LIDAR-SLAM.  Every function
 * cites the reference file:line it restates (paths relative to the reference root).
 *
 * Parity status
 *   - orc_edt*, orc_score_lattice (incl. the 3x3x3 FastMatch schedule) are PINNED:
 *     tests/test_oracle_vs_reference.py checks them bit-for-bit against the
 *     reference's own functions compiled unmodified into oracle/_ref/ (see
 *     oracle/Makefile) and against fixtures under tests/golden/ generated from
 *     those same objects (tests/golden/make_golden.py).
 *   - orc_occupational_grid is PINNED the same way (reference OccupationalGrid run on its own
 *     globals, both pixel sizes).
 *   - orc_read_scan, orc_transform, orc_extract_local_map are PINNED against the reference's
 *     readAScan / Transform / ExtractLocalMap run on its own globals (tests/test_frontend.py);
 *     orc_grow_map restates code that is inline in main() and is pinned end to end by the replay.
 *   - orc_score_poses restates the same per-beam arithmetic for an arbitrary pose
 *     list; it is pinned through the lattice (a lattice expanded to a pose list
 *     must give identical scores).
 *   - orc_weights_resample and orc_pyramid_match are "PARITY UNPINNED": the
 *     reference has no particle filter and no pyramid; these
 *     functions ARE the definition the CUDA path is held to.
 *
 * Build: gcc -O2 -ffp-contract=off (never -ffast-math / -march=native): float
 * products and sums must round separately, exactly like the reference build.
 */
#ifndef SLAM_ORACLE_H
#define SLAM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- occupancy-grid rasterisation ---------- */

/* One level of OccupationalGrid, Subsystem_1/main.c:271-354: bounding box of the points,
 * 3-pixel margin, grid size, and a 1 in every cell a point rounds into.  `grid` is
 * [cap_rows][stride]; the whole capacity is zeroed first (main.c:319).  Returns 0, or -1 when
 * the grid does not fit the capacity (the reference would overflow its fixed arrays). */
int orc_occupational_grid(const float *x, const float *y, int n, float pixel_size, int32_t *grid,
                          int stride, int cap_rows, int cap_cols, int *rows, int *cols,
                          float *min_x, float *min_y);

/* ---- scan front end and map points ------ */

/* SetLidarParameters, Subsystem_1/main.c:45-58: angles by repeated float addition. */
void orc_lidar_angles(float angle_min, float angle_increment, int n, float *angles);
/* readAScan, main.c:71-95: drops r < range_min | r > max_range (int), x = r * cosf(a),
 * y = r * sinf(a), compacted in beam order.  Returns scan.size. */
/* readDatasetLineByLine (Subsystem_1/main.c:22-30) applied to a whole file: fscanf(fp, "%f,", &value) until it
 * stops converting; returns the number of values (at most max_values). */
long orc_read_csv(const char *path, float *out, long max_values);

int orc_read_scan(const float *ranges, const float *angles, int n, float range_min, int max_range,
                  float *x, float *y);
/* Transform, main.c:97-118. */
void orc_transform(const float *x, const float *y, int n, const float pose[3], float *tx, float *ty);
/* ExtractLocalMap, main.c:155-198.  Returns local_map.size. */
int orc_extract_local_map(const float *tx, const float *ty, int n, const float *map_x, const float *map_y,
                          int map_size, float border, float *local_x, float *local_y);
/* Map growth inside main(), main.c:942-948: for j < best_hits_size: best_hits[j] > 1.5 appends
 * (tx[j], ty[j]) at map_size + k.  Returns newPointSize.  PINNED only end to end (the code is inline
 * in main(): tests/test_replay.py compares a whole replay's pose trace and map dump). */
int orc_grow_map(const float *best_hits, int best_hits_size, const float *tx, const float *ty,
                 float *map_x, float *map_y, int map_size);

/* ---- distance transform ------------------------------------------------- */

/* Window radius implied by max_dist: number of integers d >= 1 with
 * (float)(d*d) < max_dist*max_dist  (the compare at Subsystem_1/main.c:235). */
int orc_edt_radius(float max_dist);

/* Clamped exact Euclidean distance transform, separable integer form.
 * Semantics of Subsystem_1/main.c:223-245 (== main_accelerated.c:215-248):
 *   out[r][c] = d2min < max_dist^2 ? sqrtf(d2min) : max_dist
 * where d2min is the integer squared distance to the nearest non-zero cell of the
 * rows x cols sub-rectangle.  Strides are in elements.  Only [0,rows)x[0,cols) of
 * `out` is written (main.c:225-226 loop bounds). */
void orc_edt(const int32_t *occ, int occ_stride, float *out, int out_stride,
             int rows, int cols, float max_dist);

/* Literal restatement of the per-cell search loop nest of main.c:223-245
 * (O(cells^2)); used only to validate orc_edt on small grids. */
void orc_edt_percell(const int32_t *occ, int occ_stride, float *out, int out_stride,
                     int rows, int cols, float max_dist);

/* Literal restatement of the per-obstacle scatter form of
 * main_accelerated.c:215-248 (double `dist`, sqrt((float)dist)). */
void orc_edt_scatter(const int32_t *occ, int occ_stride, float *out, int out_stride,
                     int rows, int cols, float max_dist);

/* ---- scan matching ------------------------------------------------------ */

typedef struct {
    const float *field;   /* distance field, row-major (metric_grid, main.c:203) */
    int rows, cols;       /* grid_size[0], grid_size[1] (main.c:202)            */
    int stride;           /* elements per row (200 / 400 in the reference)      */
    float pixel_size;     /* main.c:204 */
    float top_left_x;     /* top_left_corner[0] = minX (main.c:359) */
    float top_left_y;     /* top_left_corner[1] = minY (main.c:360) */
} orc_map;

typedef struct {
    int64_t best_index;      /* linear index (itheta*ntx + itx)*nty + ity; lowest on ties */
    float   best_score;
    float   best_pose[3];    /* x, y, theta of the winner (main.c:554-556) */
    int     best_hits;       /* in-bounds beam count of the winner (main.c:557) */
    int     last_hits;       /* in-bounds beam count of the LAST candidate scored */
} orc_match;

/* Axis value k of an n-point lattice axis centred on p with step s:
 * p + (float)(k - n/2) * s, product and sum rounded separately.  For n == 3 this
 * is bit-identical to {p - s, p, p + s} of main.c:424-426. */
float orc_lattice_value(float p, float s, int k, int n);

/* Correlative scan match over the lattice {theta} x {tx} x {ty}.
 * Restates main.c:381-596 for one sweep of an n[0] x n[1] x n[2] (theta, tx, ty)
 * lattice: loop order theta -> tx -> ty, per-beam arithmetic of main.c:417-421,
 * 433-438, 459-465, 482-485, 500-503, 508-521, strict `<` update of :549-569.
 * step = {tx step, ty step, theta step} (the reference passes t, t, r).
 * scores     : optional [n0*n1*n2] output of every candidate's score
 * last_hit_values : optional [nbeams] output; receives the in-bounds field values
 *              of the LAST candidate in beam order (main.c:515 overwrites
 *              bestHits for every candidate -- SURVEY.md section 7 hard part 4). */
void orc_score_lattice(const orc_map *map, const float *scan_x, const float *scan_y,
                       int nbeams, const float pose0[3], const float step[3],
                       const int n[3], float *scores, float *last_hit_values,
                       orc_match *result);

/* Same per-beam arithmetic for an arbitrary list of poses / particles.
 * poses: [P][3] = x, y, theta.  ct/st: optional [P] cos/sin of theta (host libm
 * cosf/sinf when NULL, as main.c:434-435).  scores: [P].  hits: optional [P]. */
void orc_score_poses(const orc_map *map, const float *scan_x, const float *scan_y,
                     int nbeams, const float *poses, const float *ct, const float *st,
                     int64_t P, float *scores, int32_t *hits, orc_match *result);

/* The reference's FastMatch / FastMatch2 call (main.c:381-596 / 598-809): 3x3x3
 * lattice, five identical sweeps, returns pose, bestHits_size
 * (winner's) and bestHits[] (last candidate's). */
void orc_fastmatch(const orc_map *map, const float *scan_x, const float *scan_y,
                   int nbeams, const float pose[3], const float search_resolution[3],
                   float pose_out[3], float *best_hits, int *best_hits_size);

/* ---- particle weighting + systematic resampling (PARITY UNPINNED) -------- */

/* Deterministic exp(x) for x <= 0 (FMA-free double arithmetic, rounded to float);
 * the CUDA path carries the same constants so host and device agree bitwise. */
float orc_exp_det(float x);

/* w_i = exp_det(-beta * (score_i - score_min)); q_i = (uint64)(w_i * 2^32);
 * W = sum q_i; weights[i] = (float)((double)q_i / (double)W);
 * thresholds T_k = U + floor(k*W/N), U = ((W div N) * u0_q32) >> 32;
 * ancestors[k] = first i with inclusive prefix sum C_i > T_k.
 * Returns W through wsum.  weights/ancestors may be NULL. */
void orc_weights_resample(const float *scores, int64_t N, float beta, uint32_t u0_q32,
                          float *weights, uint64_t *q, uint64_t *wsum,
                          int32_t *ancestors);

/* ---- multi-resolution match (PARITY UNPINNED beyond the 2-level a9 precedent) */

/* Coarse-to-fine search over `levels` independent maps (coarsest first), each
 * level seeded with the previous level's best pose and using steps[level][3],
 * n[level][3]; the reference's own 2-level schedule is main.c:901-918. */
void orc_pyramid_match(const orc_map *maps, int levels, const float *scan_x,
                       const float *scan_y, int nbeams, const float pose0[3],
                       const float *steps /*[levels][3]*/, const int *n /*[levels][3]*/,
                       orc_match *results /*[levels]*/);

#ifdef __cplusplus
}
#endif
#endif
