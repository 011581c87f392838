/*
 * slam_replay.c -- the reference's per-scan loop with every data step on the device.
 *
 * Host code in C on top of include/b200slam.h, restating the CONTROL FLOW of the reference's
 * main() (Subsystem_1/main.c:825-990 == Subsystem_1/main_accelerated.c:840-1005): CSV replay,
 * first scan initialises the map, then per scan readAScan -> (after a mini update: Transform,
 * ExtractLocalMap, OccupationalGrid + both distance transforms) -> constant-velocity guess ->
 * FastMatch on the coarse or FastMatch2 on the fine grid -> FastMatch2 refinement -> mini-update
 * test -> map growth.
 *
 * Nothing but poses crosses PCIe in the steady state:
 *   - the dataset is read once, uploaded as raw text and parsed ON THE GPU (b200slam_csv_ingest: the
 *     reference's fscanf("%f,") loop, main.c:22-30, bit for bit); readAScan then reads its ranges from
 *     the resident values;
 *   - a scan is ONE kernel -- readAScan, FastMatch and FastMatch2 fused (no host step between them: scan.size
 *     stays on the device, and the second match centres its lattice on the first one's winner);
 *   - between map updates the loop itself runs ON THE DEVICE (b200slam_scan_chain_*): the motion model
 *     (main.c:875-898), both lattices' tables -- cosf / sinf as glibc computes them -- and the mini-update test
 *     (main.c:928-940) are evaluated by the kernels, which the host queues up to CHAIN_DEPTH scans ahead of the
 *     results it reads from a ring of mapped memory (only to print the poses and to see a mini update coming);
 *     B200SLAM_REPLAY_NO_CHAIN=1 keeps r02's host-driven loop: one kernel and one synchronisation per scan;
 *   - map growth (main.c:942-948) is queued without reading its count back; only a map rebuild (2 % of
 *     the scans) reads sizes, because the grid geometry is computed with the reference's host float
 *     operations (main.c:272-305).
 * Output is the reference's, byte for byte: "scan N" / "pose = ..." lines (main.c:860, 965) and the map
 * dump "%f,%f\n" (main.c:983-985).
 *
 *     b200slam_replay <lidar.csv> <map_out.csv> [nscans = 3480]
 *     B200SLAM_REPLAY_HOST_PARSE=1: parse with the host's fscanf and upload 4 bytes per beam per scan instead
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "b200slam.h"

#define COLUMN 1079                                /* main.c:7 */
#define CHAIN_DEPTH 24                             /* scans queued ahead of the one being fetched (ring: 64) */
#define CHAIN_BATCH 8                              /* scans per kernel launch */

static b200slam_ctx *ctx;

static void must(int rc, const char *what)
{
    if (rc) {
        fprintf(stderr, "b200slam_replay: %s failed (%d): %s\n", what, rc, b200slam_last_error(ctx));
        exit(1);
    }
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int read_row(FILE *fp, float *ranges)
{
    for (int k = 0; k < COLUMN; k++) {             /* main.c:22-30 */
        float value;
        if (fscanf(fp, "%f,", &value) != 1) return k == 0 ? 0 : -1;
        ranges[k] = value;
    }
    return 1;
}

static void build_grids(b200slam_map *coarse, b200slam_map *fine, float pixel, float pixel2)
{
    /* OccupationalGrid, main.c:271-363: both levels rasterised from the resident local map and
     * transformed in place */
    must(b200slam_map_rasterise_local(ctx, coarse, pixel, NULL, NULL, NULL), "rasterise (coarse)");
    must(b200slam_map_edt(ctx, coarse, 10.0f), "EDT (coarse)");
    must(b200slam_map_rasterise_local(ctx, fine, pixel2, NULL, NULL, NULL), "rasterise (fine)");
    must(b200slam_map_edt(ctx, fine, 10.0f), "EDT (fine)");
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <lidar.csv> <map_out.csv> [nscans]\n", argv[0]);
        return 2;
    }
    const int row = argc > 3 ? atoi(argv[3]) : 3480;          /* main_accelerated.c:6 */
    const int host_parse = getenv("B200SLAM_REPLAY_HOST_PARSE") != NULL;
    clock_t start = clock();

    float pose[3] = {0, 0, 0};
    /* main.c:832-840 */
    const float fastResolution[3] = {0.05f, 0.05f, 0.008727f};
    const float fastResolution2[3] = {0.025f, 0.025f, 0.004363f};
    const float borderSize = 1;
    const float pixelSize = 0.2f, pixelSize2 = 0.1f;
    const float miniUpdateDT = 0.3f, miniUpdateDR = 0.0872665f;

    FILE *fp = fopen(argv[1], "r");
    if (!fp) { fprintf(stderr, "Failed to open the file.\n"); return 1; }
    const char *dev = getenv("B200SLAM_DEVICE");
    must(b200slam_create(&ctx, dev ? atoi(dev) : 0), "b200slam_create");

    /* SetLidarParameters, main.c:45-58; cos / sin of the beam angles as readAScan takes them (:90-91) */
    static float cos_a[COLUMN], sin_a[COLUMN], ranges[COLUMN];
    {
        float angle = -2.351831f;
        for (int i = 0; i < COLUMN; i++) {
            cos_a[i] = cosf(angle);
            sin_a[i] = sinf(angle);
            angle += 0.004363f;
        }
    }
    must(b200slam_lidar_set(ctx, cos_a, sin_a, COLUMN, 0.023f), "b200slam_lidar_set");
    b200slam_map *coarse = NULL, *fine = NULL;
    must(b200slam_map_create(ctx, 200, 200, &coarse), "map_create");     /* main.c:201 */
    must(b200slam_map_create(ctx, 400, 400, &fine), "map_create");       /* main.c:207 */

    /* the dataset: raw text -> GPU -> floats (readDatasetLineByLine, main.c:22-30, for every row at once) */
    double t_parse = 0.0;
    int64_t nvalues = 0;
    if (!host_parse) {
        const double tp = now_s();
        fseek(fp, 0, SEEK_END);
        const long nbytes = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        char *text = NULL;
        must(b200slam_host_alloc(ctx, (size_t)(nbytes > 0 ? nbytes : 1), (void **)&text), "host_alloc");
        if (nbytes > 0 && fread(text, 1, (size_t)nbytes, fp) != (size_t)nbytes) { fprintf(stderr, "short read\n"); return 1; }
        must(b200slam_csv_ingest(ctx, text, (size_t)nbytes, NULL, (int64_t)row * COLUMN, &nvalues), "csv_ingest");
        must(b200slam_host_free(ctx, text), "host_free");
        t_parse = now_s() - tp;
        if (nvalues < COLUMN) { fprintf(stderr, "empty dataset\n"); return 1; }
    }

    float (*path)[3] = malloc(sizeof(float[3]) * (size_t)(row > 0 ? row : 1));
    float map_pose[3];
    int size = 0;

    /* scan 0 (main.c:843-853) */
    if (host_parse) {
        if (read_row(fp, ranges) != 1) { fprintf(stderr, "empty dataset\n"); return 1; }
        must(b200slam_scan_read_async(ctx, ranges, 24), "scan_read");
    } else {
        must(b200slam_scan_read_resident_async(ctx, 0, 24), "scan_read");
    }
    must(b200slam_scan_transform(ctx, pose), "scan_transform");
    must(b200slam_mappoints_from_scan(ctx), "mappoints_from_scan");      /* Initialise */
    for (int i = 0; i < 3; i++) { map_pose[i] = pose[i]; path[0][i] = pose[i]; }

    int miniUpdated = 1, path_iter = 1, rebuilds = 0;
    int use_chain = !host_parse && getenv("B200SLAM_REPLAY_NO_CHAIN") == NULL;
    int chain_active = 0, chain_queued = 0, chained_scans = 0, chains = 0;
    double t_read = 0, t_rebuild = 0, t_queue = 0, t_fetch = 0, t_grow = 0;      /* where the host's time goes */
    const double t_loop = now_s();
    for (int scan_iter = 1; scan_iter < row; scan_iter++) {
        printf("scan %d\n", scan_iter + 1);
        double t0 = now_s();
        if (host_parse) {
            const double tp = now_s();
            const int got = read_row(fp, ranges);
            t_parse += now_s() - tp;
            if (got != 1) {
                fprintf(stderr, "b200slam_replay: dataset ends at scan %d (asked for %d)\n", scan_iter, row);
                return 1;
            }
        } else if ((int64_t)(scan_iter + 1) * COLUMN > nvalues) {
            fprintf(stderr, "b200slam_replay: dataset ends at scan %d (asked for %d)\n", scan_iter, row);
            return 1;
        }
        /* After a mini update the scan is needed by Transform / ExtractLocalMap before the matches: readAScan is
         * its own kernel.  Otherwise (98 % of the scans) readAScan + FastMatch2 + FastMatch2 are ONE kernel below. */
        if (miniUpdated) {
            if (host_parse) must(b200slam_scan_read_async(ctx, ranges, 24), "scan_read");
            else must(b200slam_scan_read_resident_async(ctx, (int64_t)scan_iter * COLUMN, 24), "scan_read");
        }
        t_read += now_s() - t0; t0 = now_s();
        int scan_transform_flag = 0;
        if (miniUpdated) {                                                /* main.c:865-872 */
            must(b200slam_scan_transform(ctx, pose), "scan_transform");
            scan_transform_flag = 1;
            must(b200slam_local_map_extract(ctx, borderSize, NULL), "local_map_extract");
            build_grids(coarse, fine, pixelSize, pixelSize2);
            rebuilds++;
        }
        t_rebuild += now_s() - t0; t0 = now_s();
        /* constant-velocity motion model, main.c:875-898 */
        float pose_guess[3];
        if (scan_iter > 1) {
            for (int i = 0; i < 3; i++) {
                const float dp = pose[i] - path[path_iter - 2][i];        /* DiffPose(previous_pose, pose) */
                pose_guess[i] = pose[i] + dp;
            }
        } else {
            for (int i = 0; i < 3; i++) pose_guess[i] = pose[i];
        }
        /* main.c:901-922: FastMatch (coarse grid after a rebuild, else the fine one) then FastMatch2 from its
         * result -- two kernels back to back, one synchronisation */
        int chained = 0;
        if (!miniUpdated && use_chain) {
            /* the device carries the poses from here on: (re)start the chain at this scan, keep it CHAIN_DEPTH ahead.
             * The device's cosf / sinf are glibc's for |theta| <= 16: stay well inside */
            if (chain_active && chain_queued <= scan_iter && fabsf(pose[2]) >= 13.0f) chain_active = 0;   /* host-driven from here */
            if (!chain_active && fabsf(pose[2]) < 13.0f) {
                const int rc = b200slam_scan_chain_begin(ctx, scan_iter, pose, scan_iter > 1 ? path[path_iter - 2] : NULL, map_pose,
                                                         miniUpdateDT, miniUpdateDR);
                if (rc) {
                    fprintf(stderr, "b200slam_replay: device-side loop unavailable (%s); host-driven loop\n", b200slam_last_error(ctx));
                    use_chain = 0;
                } else {
                    chain_active = 1;
                    chain_queued = scan_iter;
                    chains++;
                }
            }
            if (chain_active) {
                while (chain_queued < row && chain_queued + CHAIN_BATCH <= scan_iter + CHAIN_DEPTH + (chain_queued == scan_iter ? CHAIN_BATCH : 0)) {
                    int nb = CHAIN_BATCH;                                 /* one launch = up to CHAIN_BATCH consecutive scans */
                    if (chain_queued + nb > row) nb = row - chain_queued;
                    while (nb > 0 && (int64_t)(chain_queued + nb) * COLUMN > nvalues) nb--;
                    while (nb > 1 && fabsf(pose[2]) + 0.1f * (float)(chain_queued + nb - scan_iter) >= 15.0f) nb--;
                    if (nb <= 0) break;
                    must(b200slam_scan_chain_step_async(ctx, chain_queued, nb, (int64_t)chain_queued * COLUMN, 24, fine, fine,
                                                        fastResolution, fastResolution2), "scan chain step");
                    chain_queued += nb;
                }
                if (chain_queued <= scan_iter) { fprintf(stderr, "b200slam_replay: dataset ends at scan %d\n", scan_iter); return 1; }
                chained = 1;
                chained_scans++;
            }
        }
        if (chained)
            ;
        else if (miniUpdated)
            must(b200slam_fastmatch_pair_async(ctx, coarse, fine, pose_guess, fastResolution, fastResolution2), "fastmatch pair");
        else if (host_parse)
            must(b200slam_scan_step_async(ctx, ranges, 24, fine, fine, pose_guess, fastResolution, fastResolution2), "scan step");
        else
            must(b200slam_scan_step_resident_async(ctx, (int64_t)scan_iter * COLUMN, 24, fine, fine, pose_guess, fastResolution,
                                                   fastResolution2), "scan step");
        t_queue += now_s() - t0; t0 = now_s();
        int stopped = 0;
        if (chained) must(b200slam_scan_chain_fetch(ctx, scan_iter, NULL, pose, &size, NULL, &stopped), "scan chain fetch");
        else must(b200slam_fastmatch_pair_fetch(ctx, NULL, pose, &size, NULL), "fastmatch pair fetch");
        t_fetch += now_s() - t0; t0 = now_s();
        /* mini update, main.c:928-961 */
        float dp[3];
        for (int i = 0; i < 3; i++) dp[i] = fabsf(pose[i] - map_pose[i]);
        const int update = dp[0] > miniUpdateDT || dp[1] > miniUpdateDT || dp[2] > miniUpdateDR;
        if (chained) {
            if (update != stopped) {               /* the device evaluated the same float compares */
                fprintf(stderr, "b200slam_replay: scan %d: device mini-update test %d, host %d\n", scan_iter, stopped, update);
                return 1;
            }
            if (stopped) chain_active = 0;         /* nothing queued behind this scan ran; the scan on the device is this one's */
        }
        if (update) {
            miniUpdated = 1;
            if (!scan_transform_flag) must(b200slam_scan_transform(ctx, pose), "scan_transform");
            must(b200slam_mappoints_grow_async(ctx, 1.5f), "mappoints_grow");
            for (int i = 0; i < 3; i++) map_pose[i] = pose[i];
        } else {
            miniUpdated = 0;
        }
        t_grow += now_s() - t0;
        printf("pose = %f  %f  %f\n", pose[0], pose[1], pose[2]);
        for (int i = 0; i < 3; i++) path[path_iter][i] = pose[i];
        path_iter++;
    }
    must(b200slam_sync(ctx), "sync");
    const double loop_s = now_s() - t_loop;
    printf("time taken = %f\n", (double)(clock() - start) / CLOCKS_PER_SEC);
    fclose(fp);

    int n = 0;
    must(b200slam_mappoints_download(ctx, NULL, NULL, &n), "mappoints_download");
    float *mx = malloc(sizeof(float) * (size_t)(n > 0 ? n : 1)), *my = malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    must(b200slam_mappoints_download(ctx, mx, my, &n), "mappoints_download");
    FILE *fp1 = fopen(argv[2], "w");
    if (!fp1) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    for (int j = 0; j < n; j++) fprintf(fp1, "%f,%f\n", mx[j], my[j]);   /* main.c:983-985 */
    fclose(fp1);
    const double dev_s = loop_s - (host_parse ? t_parse : 0.0);
    fprintf(stderr, "b200slam_replay: %d scans, %d map points, %d map rebuilds, %llu kernel launches; dataset %s in %.4f s; "
                    "loop %.3f s wall -> %.1f us per scan on the device path\n", row, n, rebuilds,
            (unsigned long long)b200slam_launch_count(ctx), host_parse ? "parsed by fscanf" : "read + parsed on the GPU",
            t_parse, loop_s, row > 1 ? 1e6 * dev_s / (row - 1) : 0.0);
    fprintf(stderr, "b200slam_replay: %d of the scans ran in %d device-side chains (%d scans per launch, up to %d queued ahead)\n", chained_scans, chains,
            CHAIN_BATCH, CHAIN_DEPTH);
    fprintf(stderr, "b200slam_replay: host time per scan: queue readAScan %.1f us, map rebuilds %.1f us (%d of them), queue the match pair "
                    "%.1f us, wait for its result %.1f us, mini update / growth %.1f us\n", 1e6 * t_read / (row - 1),
            1e6 * t_rebuild / (row - 1), rebuilds, 1e6 * t_queue / (row - 1), 1e6 * t_fetch / (row - 1), 1e6 * t_grow / (row - 1));
    free(mx); free(my); free(path);
    b200slam_map_destroy(ctx, coarse);
    b200slam_map_destroy(ctx, fine);
    b200slam_destroy(ctx);
    return 0;
}
