/*
 * slam_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE
 * ONLY (see slam_oracle.h): the product never links or loads this file.
 *
 * Written in the reference's own style (plain single-threaded C, float arithmetic
 * with separately rounded products and sums).  Build with
 *     gcc -O2 -ffp-contract=off -fPIC -shared slam_oracle.c -lm
 * Citations are to the reference tree (Subsystem_1/main.c unless noted).
 */
#include "slam_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* EDT                                                                        */
/* ------------------------------------------------------------------------- */

int orc_edt_radius(float max_dist)
{
    /* main.c:235  `(float)dist_square < min_dist * min_dist` with min_dist ==
     * MAX_DIST: a cell at integer distance d can only matter if d*d passes it. */
    float thr = max_dist * max_dist;
    int r = 0;
    while ((float)((r + 1) * (r + 1)) < thr) r++;
    return r;
}

void orc_edt(const int32_t *occ, int occ_stride, float *out, int out_stride,
             int rows, int cols, float max_dist)
{
    /* Closed form of main.c:223-245: the loop keeps min_dist = sqrtf(d2) of the
     * smallest integer d2 seen so far that passes `(float)d2 < min_dist*min_dist`
     * (SURVEY.md appendix 9.1 shows the float compare never mis-orders integers
     * below 100), so the result depends only on the integer d2min.  d2min is found
     * separably: nearest occupied cell in the same column within R rows, then a
     * min over |dx| <= R of dx^2 + g^2. */
    const int R = orc_edt_radius(max_dist);
    const float thr = max_dist * max_dist;
    const int BIG = 1 << 20;
    int *g = (int *)malloc(sizeof(int) * (size_t)rows * (size_t)cols);
    if (!g) abort();

    for (int c = 0; c < cols; ++c) {
        for (int r = 0; r < rows; ++r) {
            int best = BIG;
            int lo = r - R < 0 ? 0 : r - R;
            int hi = r + R > rows - 1 ? rows - 1 : r + R;
            for (int j = lo; j <= hi; ++j) {
                if (occ[(size_t)j * occ_stride + c]) {      /* main.c:233 */
                    int d = j > r ? j - r : r - j;
                    if (d < best) best = d;
                }
            }
            g[(size_t)r * cols + c] = best;
        }
    }
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < cols; ++c) {
            int best = BIG;
            int lo = c - R < 0 ? 0 : c - R;
            int hi = c + R > cols - 1 ? cols - 1 : c + R;
            for (int i = lo; i <= hi; ++i) {
                int gv = g[(size_t)r * cols + i];
                if (gv < BIG) {
                    int dx = i - c;
                    int d2 = dx * dx + gv * gv;              /* main.c:216-220 */
                    if (d2 < best) best = d2;
                }
            }
            /* main.c:230,235-236,241 */
            out[(size_t)r * out_stride + c] =
                (best < BIG && (float)best < thr) ? sqrtf((float)best) : max_dist;
        }
    }
    free(g);
}

void orc_edt_percell(const int32_t *occ, int occ_stride, float *out, int out_stride,
                     int rows, int cols, float max_dist)
{
    /* main.c:223-245, loop for loop (width == cols, height == rows). */
    for (int y = 0; y < rows; ++y) {
        for (int x = 0; x < cols; ++x) {
            if (occ[(size_t)y * occ_stride + x]) {           /* :227 */
                out[(size_t)y * out_stride + x] = 0;
            } else {
                float min_dist = max_dist;                   /* :230 */
                for (int j = 0; j < rows; ++j) {
                    for (int i = 0; i < cols; ++i) {
                        if (occ[(size_t)j * occ_stride + i]) {
                            int xt = x - i, yt = y - j;      /* :216-220 */
                            int dist_square = xt * xt + yt * yt;
                            if ((float)dist_square < min_dist * min_dist)   /* :235 */
                                min_dist = sqrtf((float)dist_square);
                        }
                    }
                }
                out[(size_t)y * out_stride + x] = min_dist;  /* :241 */
            }
        }
    }
}

void orc_edt_scatter(const int32_t *occ, int occ_stride, float *out, int out_stride,
                     int rows, int cols, float max_dist)
{
    /* Subsystem_1/main_accelerated.c:215-248: first index bounded by the 4th
     * argument (rows), second by the 3rd (cols); `== 1` test; double dist. */
    float *distance = (float *)malloc(sizeof(float) * (size_t)rows * (size_t)cols);
    if (!distance) abort();
    for (size_t k = 0; k < (size_t)rows * (size_t)cols; ++k) distance[k] = max_dist;   /* :219-224 */
    for (int x = 0; x < rows; ++x) {
        for (int y = 0; y < cols; ++y) {
            if (occ[(size_t)x * occ_stride + y] == 1) {       /* :229 */
                for (int i = 0; i < rows; ++i) {
                    for (int j = 0; j < cols; ++j) {
                        double dist = (x - i) * (x - i) + (y - j) * (y - j);     /* :232 */
                        float dd = distance[(size_t)i * cols + j];
                        if (dist < dd * dd)                                       /* :233 */
                            distance[(size_t)i * cols + j] = sqrt((float)dist);   /* :234 */
                    }
                }
            }
        }
    }
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j)
            out[(size_t)i * out_stride + j] = distance[(size_t)i * cols + j];    /* :242-246 */
    free(distance);
}

/* ------------------------------------------------------------------------- */
/* Scan matching                                                              */
/* ------------------------------------------------------------------------- */

float orc_lattice_value(float p, float s, int k, int n)
{
    float off = (float)(k - n / 2) * s;
    return p + off;
}

void orc_score_lattice(const orc_map *map, const float *scan_x, const float *scan_y,
                       int nbeams, const float pose0[3], const float step[3],
                       const int n[3], float *scores, float *last_hit_values,
                       orc_match *result)
{
    const int nth = n[0], ntx = n[1], nty = n[2];
    const float ipixel = 1 / map->pixel_size;                /* :383 */
    const float minX = map->top_left_x;                      /* :384 */
    const float minY = map->top_left_y;                      /* :385 */
    const int nRows = map->rows;                             /* :388 */
    const int nCols = map->cols;                             /* :389 */

    float *pixelScan_x = (float *)malloc(sizeof(float) * (size_t)(nbeams + 1));
    float *pixelScan_y = (float *)malloc(sizeof(float) * (size_t)(nbeams + 1));
    float *S_x = (float *)malloc(sizeof(float) * (size_t)(nbeams + 1));
    float *S_y = (float *)malloc(sizeof(float) * (size_t)(nbeams + 1));
    int *Sx = (int *)malloc(sizeof(int) * (size_t)(nbeams + 1));
    int *Sy = (int *)malloc(sizeof(int) * (size_t)(nbeams + 1));
    float *theta = (float *)malloc(sizeof(float) * (size_t)nth);
    float *ct = (float *)malloc(sizeof(float) * (size_t)nth);
    float *st = (float *)malloc(sizeof(float) * (size_t)nth);
    float *tx = (float *)malloc(sizeof(float) * (size_t)ntx);
    float *ty = (float *)malloc(sizeof(float) * (size_t)nty);
    float *Sx_temp = (float *)malloc(sizeof(float) * (size_t)ntx);
    float *Sy_temp = (float *)malloc(sizeof(float) * (size_t)nty);

    for (int a = 0; a < nbeams; a++) {                       /* :417-421 */
        pixelScan_x[a] = scan_x[a] * ipixel;
        pixelScan_y[a] = scan_y[a] * ipixel;
    }
    for (int i = 0; i < nth; i++) {                          /* :424, :433-435 */
        theta[i] = orc_lattice_value(pose0[2], step[2], i, nth);
        ct[i] = cosf(theta[i]);
        st[i] = sinf(theta[i]);
    }
    for (int i = 0; i < ntx; i++) {                          /* :425, :436 */
        tx[i] = orc_lattice_value(pose0[0], step[0], i, ntx);
        Sx_temp[i] = (tx[i] - minX) * ipixel;
    }
    for (int i = 0; i < nty; i++) {                          /* :426, :437 */
        ty[i] = orc_lattice_value(pose0[1], step[1], i, nty);
        Sy_temp[i] = (ty[i] - minY) * ipixel;
    }

    float bestScore = INFINITY;                              /* :403 */
    result->best_index = -1;
    result->best_score = INFINITY;
    result->best_pose[0] = pose0[0];                         /* :422 */
    result->best_pose[1] = pose0[1];
    result->best_pose[2] = pose0[2];
    result->best_hits = 0;
    result->last_hits = 0;

    for (int theta_index = 0; theta_index < nth; theta_index++) {            /* :443 */
        for (int q = 0; q < nbeams; q++) {                                   /* :459-465 */
            S_x[q] = (pixelScan_x[q] * ct[theta_index]) + (pixelScan_y[q] * (st[theta_index]));
            S_y[q] = (pixelScan_x[q] * (-st[theta_index])) + (pixelScan_y[q] * ct[theta_index]);
        }
        for (int tx_index = 0; tx_index < ntx; tx_index++) {                 /* :468 */
            for (int i = 0; i < nbeams; i++)                                 /* :482-485 */
                Sx[i] = (int)roundf(S_x[i] + Sx_temp[tx_index]) + 1;
            for (int ty_index = 0; ty_index < nty; ty_index++) {             /* :487 */
                for (int i2 = 0; i2 < nbeams; i2++)                          /* :500-503 */
                    Sy[i2] = (int)roundf(S_y[i2] + Sy_temp[ty_index]) + 1;
                int ixy_index = 0;
                float score = 0;                                             /* :507 */
                for (int i3 = 0; i3 < nbeams; i3++) {                        /* :508-521 */
                    int temp_Sx = Sx[i3];
                    int temp_Sy = Sy[i3];
                    if ((temp_Sx > 1) && (temp_Sy > 1) && (temp_Sx < nCols) && (temp_Sy < nRows)) {
                        float v = map->field[(size_t)(temp_Sy - 1) * map->stride + (temp_Sx - 1)];
                        if (last_hit_values) last_hit_values[ixy_index] = v; /* :515 */
                        score = score + v;                                   /* :516 */
                        ixy_index++;
                    }
                }
                int64_t lin = ((int64_t)theta_index * ntx + tx_index) * nty + ty_index;
                if (scores) scores[lin] = score;
                result->last_hits = ixy_index;
                if (score < bestScore) {                                     /* :549-569 */
                    result->best_pose[0] = tx[tx_index];
                    result->best_pose[1] = ty[ty_index];
                    result->best_pose[2] = theta[theta_index];
                    result->best_hits = ixy_index;                           /* :557 */
                    result->best_index = lin;
                    bestScore = score;
                }
            }
        }
    }
    result->best_score = bestScore;

    free(pixelScan_x); free(pixelScan_y); free(S_x); free(S_y); free(Sx); free(Sy);
    free(theta); free(ct); free(st); free(tx); free(ty); free(Sx_temp); free(Sy_temp);
}

void orc_score_poses(const orc_map *map, const float *scan_x, const float *scan_y,
                     int nbeams, const float *poses, const float *ct_in, const float *st_in,
                     int64_t P, float *scores, int32_t *hits, orc_match *result)
{
    const float ipixel = 1 / map->pixel_size;                /* :383 */
    const float minX = map->top_left_x, minY = map->top_left_y;
    const int nRows = map->rows, nCols = map->cols;
    float bestScore = INFINITY;
    if (result) {
        result->best_index = -1; result->best_score = INFINITY;
        result->best_pose[0] = result->best_pose[1] = result->best_pose[2] = 0;
        result->best_hits = 0; result->last_hits = 0;
    }
    for (int64_t p = 0; p < P; ++p) {
        const float px0 = poses[3 * p + 0], py0 = poses[3 * p + 1], th = poses[3 * p + 2];
        const float ct = ct_in ? ct_in[p] : cosf(th);        /* :434 */
        const float st = st_in ? st_in[p] : sinf(th);        /* :435 */
        const float Sx_temp = (px0 - minX) * ipixel;         /* :436 */
        const float Sy_temp = (py0 - minY) * ipixel;         /* :437 */
        float score = 0;
        int nh = 0;
        for (int q = 0; q < nbeams; q++) {
            float psx = scan_x[q] * ipixel;                  /* :418 */
            float psy = scan_y[q] * ipixel;                  /* :419 */
            float S_x = (psx * ct) + (psy * (st));           /* :462 */
            float S_y = (psx * (-st)) + (psy * ct);          /* :463 */
            int temp_Sx = (int)roundf(S_x + Sx_temp) + 1;    /* :483 */
            int temp_Sy = (int)roundf(S_y + Sy_temp) + 1;    /* :501 */
            if ((temp_Sx > 1) && (temp_Sy > 1) && (temp_Sx < nCols) && (temp_Sy < nRows)) {   /* :512 */
                score = score + map->field[(size_t)(temp_Sy - 1) * map->stride + (temp_Sx - 1)];
                nh++;
            }
        }
        scores[p] = score;
        if (hits) hits[p] = nh;
        if (result) {
            result->last_hits = nh;
            if (score < bestScore) {                         /* :549 */
                bestScore = score;
                result->best_index = p;
                result->best_pose[0] = px0; result->best_pose[1] = py0; result->best_pose[2] = th;
                result->best_hits = nh;
            }
        }
    }
    if (result) result->best_score = bestScore;
}

void orc_fastmatch(const orc_map *map, const float *scan_x, const float *scan_y,
                   int nbeams, const float pose[3], const float search_resolution[3],
                   float pose_out[3], float *best_hits, int *best_hits_size)
{
    /* main.c:381-596.  t = searchResolution[0] for both translations (:386),
     * r = searchResolution[2] (:387).  The lattice is built once from the input
     * pose (:422-438) and the refinement at :577-580 is commented out, so every
     * sweep of the while loop (:440) re-scores the same 27 candidates: the first
     * sweep improves from INFINITY, the next four leave noChange set until
     * depth > maxDepth (:576-587).  The outputs after five sweeps equal the
     * outputs after one. */
    const float step[3] = { search_resolution[0], search_resolution[0], search_resolution[2] };
    const int n[3] = { 3, 3, 3 };
    orc_match m;
    memset(&m, 0, sizeof m);
    int sweeps = 0, depth = 0, iter = 0;
    float best = INFINITY;
    while (iter < 50) {                                      /* :392, :440 */
        orc_match cur;
        orc_score_lattice(map, scan_x, scan_y, nbeams, pose, step, n, NULL, best_hits, &cur);
        sweeps++;
        int noChange = !(cur.best_score < best);
        if (!noChange) { best = cur.best_score; m = cur; }
        if (noChange) { depth++; if (depth > 3) break; }     /* :393, :581-585 */
        iter++;
    }
    (void)sweeps;
    pose_out[0] = m.best_pose[0];                            /* :592-594 */
    pose_out[1] = m.best_pose[1];
    pose_out[2] = m.best_pose[2];
    *best_hits_size = m.best_hits;
}

/* ------------------------------------------------------------------------- */
/* Particle weights + systematic resampling (definition; parity unpinned)     */
/* ------------------------------------------------------------------------- */

float orc_exp_det(float x)
{
    /* exp(x), x <= 0, in FMA-free double arithmetic: x = k ln2 + r, |r| <= ln2/2,
     * degree-13 Taylor polynomial by Horner (one multiply, one add per step, each
     * rounded), scaled by 2^k through the exponent field, rounded once to float.
     * Every operation is an IEEE-754 basic operation, so the CUDA restatement
     * (__dmul_rn/__dadd_rn) is bit-identical. */
    double xd = (double)x;
    if (!(xd > -80.0)) return 0.0f;
    if (xd > 0.0) xd = 0.0;
    const double LOG2E  = 1.4426950408889634;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    double kf = nearbyint(xd * LOG2E);
    double r = xd - kf * LN2_HI;
    r = r - kf * LN2_LO;
    static const double c[14] = {
        1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040,
        1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600,
        1.0 / 6227020800.0 };
    double p = c[13];
    for (int i = 12; i >= 0; --i) { p = p * r; p = p + c[i]; }
    int k = (int)kf;
    union { uint64_t u; double d; } two_k;
    two_k.u = (uint64_t)(1023 + k) << 52;
    return (float)(p * two_k.d);
}

void orc_weights_resample(const float *scores, int64_t N, float beta, uint32_t u0_q32,
                          float *weights, uint64_t *q_out, uint64_t *wsum,
                          int32_t *ancestors)
{
    float smin = INFINITY;
    for (int64_t i = 0; i < N; ++i) if (scores[i] < smin) smin = scores[i];
    uint64_t *q = q_out ? q_out : (uint64_t *)malloc(sizeof(uint64_t) * (size_t)N);
    uint64_t W = 0;
    for (int64_t i = 0; i < N; ++i) {
        float d = scores[i] - smin;
        float xarg = -(beta * d);
        float w = orc_exp_det(xarg);
        q[i] = (uint64_t)((double)w * 4294967296.0);
        W += q[i];
    }
    if (wsum) *wsum = W;
    if (weights)
        for (int64_t i = 0; i < N; ++i) weights[i] = (float)((double)q[i] / (double)W);
    if (ancestors) {
        const uint64_t Wd = W / (uint64_t)N, Wm = W % (uint64_t)N;
        const uint64_t U = (uint64_t)(((unsigned __int128)Wd * u0_q32) >> 32);
        uint64_t C = 0;      /* inclusive prefix sum */
        int64_t i = -1;
        for (int64_t k = 0; k < N; ++k) {
            uint64_t T = U + (uint64_t)k * Wd + ((uint64_t)k * Wm) / (uint64_t)N;
            while (!(C > T)) { ++i; C += q[i]; }
            ancestors[k] = (int32_t)i;
        }
    }
    if (!q_out) free(q);
}

/* ------------------------------------------------------------------------- */
/* Multi-resolution match                                                     */
/* ------------------------------------------------------------------------- */

void orc_pyramid_match(const orc_map *maps, int levels, const float *scan_x,
                       const float *scan_y, int nbeams, const float pose0[3],
                       const float *steps, const int *n, orc_match *results)
{
    /* Generalises the reference's own coarse->fine schedule, main.c:901-918:
     * FastMatch on the coarse grid, then FastMatch2 on the fine grid seeded with
     * FastMatchParameters.pose (:918). */
    float seed[3] = { pose0[0], pose0[1], pose0[2] };
    for (int l = 0; l < levels; ++l) {
        orc_score_lattice(&maps[l], scan_x, scan_y, nbeams, seed, steps + 3 * l, n + 3 * l,
                          NULL, NULL, &results[l]);
        seed[0] = results[l].best_pose[0];
        seed[1] = results[l].best_pose[1];
        seed[2] = results[l].best_pose[2];
    }
}

/* ------------------------------------------------------------------------- */
/* Occupancy-grid rasterisation                                               */
/* ------------------------------------------------------------------------- */

int orc_occupational_grid(const float *x, const float *y, int n, float pixel_size, int32_t *grid,
                          int stride, int cap_rows, int cap_cols, int *rows, int *cols,
                          float *min_x, float *min_y)
{
    /* main.c:272-289: bounding box (strict compares, seeded with point 0) */
    float minXY[2] = {x[0], y[0]};
    float maxXY[2] = {x[0], y[0]};
    for (int a = 0; a < n; a++) {
        if (x[a] < minXY[0]) minXY[0] = x[a];
        if (x[a] > maxXY[0]) maxXY[0] = x[a];
        if (y[a] < minXY[1]) minXY[1] = y[a];
        if (y[a] > maxXY[1]) maxXY[1] = y[a];
    }
    /* main.c:296-305: 3-pixel margin, size = round(extent / pixel) + 1 */
    int Sgrid[2];
    for (int a = 0; a < 2; a++) {
        minXY[a] -= (3 * pixel_size);
        maxXY[a] += (3 * pixel_size);
        Sgrid[a] = (int)roundf((maxXY[a] - minXY[a]) / pixel_size) + 1;
    }
    *cols = Sgrid[0];                       /* grid_size[1] = Sgrid[0], main.c:313-314 */
    *rows = Sgrid[1];
    *min_x = minXY[0];                      /* top_left_corner, main.c:359-360 */
    *min_y = minXY[1];
    if (Sgrid[0] > cap_cols || Sgrid[1] > cap_rows) return -1;
    for (int r = 0; r < cap_rows; ++r)      /* main.c:319 memset of the whole array */
        for (int c = 0; c < cap_cols; ++c) grid[(size_t)r * stride + c] = 0;
    /* main.c:332-353 */
    for (int a = 0; a < n; a++) {
        float x_minus_minX = x[a] - minXY[0];
        float y_minus_minY = y[a] - minXY[1];
        int hx = (int)roundf(x_minus_minX / pixel_size) + 1;
        int hy = (int)roundf(y_minus_minY / pixel_size) + 1;
        int idx = (((hy - 1) * Sgrid[0]) + hx) - 1;
        int idx_row = idx / Sgrid[0];
        int idx_col = idx % Sgrid[0];
        grid[(size_t)idx_row * stride + idx_col] = 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Scan front end and map points                                              */
/* ------------------------------------------------------------------------- */

void orc_lidar_angles(float angle_min, float angle_increment, int n, float *angles)
{
    /* main.c:53-57 */
    float angle = angle_min;
    for (int i = 0; i < n; i++) {
        angles[i] = angle;
        angle += angle_increment;
    }
}

int orc_read_scan(const float *ranges, const float *angles, int n, float range_min, int max_range,
                  float *x, float *y)
{
    /* main.c:73-94; maxRange is an int there, converted for the compare */
    int maxRange = max_range;
    int valid_points = 0;
    for (int i = 0; i < n; i++) {
        if ((ranges[i] < range_min) | (ranges[i] > maxRange)) {
            continue;
        } else {
            float test_input = ranges[i];
            float lidar_angle = angles[i];
            x[valid_points] = test_input * cosf(lidar_angle);
            y[valid_points] = test_input * sinf(lidar_angle);
            valid_points++;
        }
    }
    return valid_points;
}

void orc_transform(const float *x, const float *y, int n, const float pose[3], float *tx, float *ty)
{
    /* main.c:98-117 */
    float px = pose[0];
    float py = pose[1];
    float theta = pose[2];
    float ct = cosf(theta);
    float st = sinf(theta);
    for (int i = 0; i < n; i++) {
        float scan_x = x[i];
        float scan_y = y[i];
        tx[i] = (ct * scan_x + st * scan_y) + px;
        ty[i] = (-st * scan_x + ct * scan_y) + py;
    }
}

int orc_extract_local_map(const float *tx, const float *ty, int n, const float *map_x, const float *map_y,
                          int map_size, float border, float *local_x, float *local_y)
{
    /* main.c:156-176: bounding box, strict compares seeded with point 0 */
    float minX = tx[0], minY = ty[0], maxX = tx[0], maxY = ty[0];
    for (int i = 1; i < n; i++) {
        if (tx[i] < minX) minX = tx[i];
        if (tx[i] > maxX) maxX = tx[i];
        if (ty[i] < minY) minY = ty[i];
        if (ty[i] > maxY) maxY = ty[i];
    }
    /* main.c:179-182 */
    minX = minX - border;
    minY = minY - border;
    maxX = maxX + border;
    maxY = maxY + border;
    /* main.c:185-197 */
    int valid_points = 0;
    for (int i = 0; i < map_size; i++) {
        float mx = map_x[i], my = map_y[i];
        if ((mx > minX) && (mx < maxX) && (my > minY) && (my < maxY)) {
            local_x[valid_points] = mx;
            local_y[valid_points] = my;
            valid_points++;
        }
    }
    return valid_points;
}

int orc_grow_map(const float *best_hits, int best_hits_size, const float *tx, const float *ty,
                 float *map_x, float *map_y, int map_size)
{
    /* main.c:941-948 */
    int newPointSize = 0;
    for (int j = 0; j < best_hits_size; j++) {
        if (best_hits[j] > 1.5) {
            map_x[map_size + newPointSize] = tx[j];
            map_y[map_size + newPointSize] = ty[j];
            newPointSize++;
        }
    }
    return newPointSize;
}

/* Subsystem_1/main.c:22-30 (readDatasetLineByLine): the dataset is read with fscanf(filename, "%f,", &value),
 * `column` values per scan.  Same call, until end of file. */
long orc_read_csv(const char *path, float *out, long max_values)
{
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    long n = 0;
    float value;
    while (n < max_values && fscanf(fp, "%f,", &value) == 1) out[n++] = value;     /* main.c:27 */
    fclose(fp);
    return n;
}
