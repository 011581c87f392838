/*
 * dropin.c -- the reference's own C function names on top of libb200slam.so.
 *
 * libb200slam_dropin.so exports, with the reference's exact signatures,
 *     void euclidean_distance_transform (int[200][200], float[200][200], int, int)
 *     void euclidean_distance_transform2(int[400][400], float[400][400], int, int)
 *         (Subsystem_1/main_accelerated.c:215,250; Subsystem_1/main.c:223,247 -- 3rd
 *          argument = #columns, 4th = #rows at the only call site, main.c:355-356)
 *     void FastMatch (const float POSE[3], const float searchResolution[3])
 *     void FastMatch2(const float POSE[3], const float searchResolution[3])
 *         (Subsystem_1/main.c:381,598)
 *     void OccupationalGrid(const float PIXELSIZE, const float PIXELSIZE2)
 *         (Subsystem_1/main.c:271 -- the step in front of the transform: map points are
 *          rasterised and transformed on the device, 8 bytes per point cross PCIe instead of
 *          4 bytes per cell)
 *     void readAScan(const int usableRange)            (main.c:71)
 *     void Transform(const float POSE[3])               (main.c:97)
 *     void ExtractLocalMap(const float BORDERSIZE)      (main.c:155)
 *         (the steps in front of OccupationalGrid, on the device; their results are also written
 *          back into the reference's globals scan / local_map, which main() itself reads)
 * so that the unmodified reference program, built as a shared object, calls the GPU path
 * when this library precedes it in the link order (ELF symbol interposition, see
 * INTEGRATION.md).  FastMatch* read the reference's globals `scan` and `occ_grid` and write
 * `FastMatchParameters`, exactly like the functions they replace; the struct layouts below
 * restate main.c:60-66, 200-212, 374-378 and must match them.
 *
 * Errors: the reference functions return void and cannot fail, and there is no CPU
 * fallback by design, so any failure prints the library's message and aborts.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200slam.h"

#define REF_COLUMN 1079                       /* main.c:7 */

typedef struct {                              /* main.c:60-66 */
    float x[REF_COLUMN];
    float y[REF_COLUMN];
    float tx[REF_COLUMN];
    float ty[REF_COLUMN];
    int size;
} ScanData;

typedef struct {                              /* main.c:200-212 */
    int grid[200][200];
    int grid_size[2];
    float metric_grid[200][200];
    float pixel_size;
    float top_left_corner[2];
    int grid2[400][400];
    int grid_size2[2];
    float metric_grid2[400][400];
    float pixel_size2;
    float top_left_corner2[2];
} MyGrid;

typedef struct {                              /* main.c:147-151 */
    float x[25000];
    float y[25000];
    int size;
} myLocalMap;

typedef struct {                              /* main.c:35-42 */
    float angle_min, angle_max, angle_increment, range_min, range_max;
    float angles[REF_COLUMN];
} LidarParameters;

typedef struct {                              /* main.c:121-131 */
    float x[20000];
    float y[20000];
    int size;
    float newPoints_x[4000];
    float newPoints_y[4000];
    int newPointsSize;
    float pose[3];
} MapPoints;

typedef struct {                              /* main.c:374-378 */
    float pose[3];
    float bestHits[2500];
    int bestHits_size;
} MyFastMatchParameters;

/* Defined by the reference translation unit this library is loaded next to. */
extern ScanData scan __attribute__((weak));
extern MyGrid occ_grid __attribute__((weak));
extern MyFastMatchParameters FastMatchParameters __attribute__((weak));
extern myLocalMap local_map __attribute__((weak));
extern LidarParameters lidar __attribute__((weak));
extern MapPoints map __attribute__((weak));
extern float test_input_memory[REF_COLUMN] __attribute__((weak));

static b200slam_ctx *g_ctx;
static float g_scan_x[REF_COLUMN], g_scan_y[REF_COLUMN];   /* the scan the device currently holds */
static int g_scan_size = -1;
/* Device-resident twins of metric_grid / metric_grid2, kept from the last transform so
 * FastMatch does not re-upload the field it was just handed. */
static struct {
    b200slam_map *map;
    int rows, cols;
    const float *host_field;                  /* which host array the device copy mirrors */
} g_maps[2];

static float g_lidar_angles[REF_COLUMN];                    /* the angle table the device's cos / sin came from */
static int g_lidar_set;
static int g_local_on_device;                               /* local_map == the device-resident local map */
static float g_map_x[20000], g_map_y[20000];               /* the map points the device currently holds */
static int g_map_size;

static void die(const char *what, int rc)
{
    fprintf(stderr, "libb200slam_dropin: %s failed (%d): %s\n", what, rc, b200slam_last_error(g_ctx));
    abort();
}

static b200slam_ctx *ctx(void)
{
    if (!g_ctx) {
        const char *dev = getenv("B200SLAM_DEVICE");
        int rc = b200slam_create(&g_ctx, dev ? atoi(dev) : 0);
        if (rc) die("b200slam_create", rc);
    }
    return g_ctx;
}

/* One device map per reference grid, created at the reference's fixed capacity
 * (200 x 200 / 400 x 400, main.c:201-209) and resized to the current grid_size. */
static b200slam_map *slot_map(int slot, int rows, int cols)
{
    const int cap = slot ? 400 : 200;
    if (rows > cap || cols > cap) {
        fprintf(stderr, "libb200slam_dropin: grid %d x %d exceeds the reference capacity %d x %d\n", rows, cols, cap, cap);
        abort();
    }
    if (!g_maps[slot].map) {
        int rc = b200slam_map_create(ctx(), cap, cap, &g_maps[slot].map);
        if (rc) die("b200slam_map_create", rc);
        g_maps[slot].host_field = NULL;
    }
    if (g_maps[slot].rows != rows || g_maps[slot].cols != cols) {
        int rc = b200slam_map_resize(g_maps[slot].map, rows, cols);
        if (rc) die("b200slam_map_resize", rc);
        g_maps[slot].rows = rows;
        g_maps[slot].cols = cols;
        g_maps[slot].host_field = NULL;
    }
    return g_maps[slot].map;
}

static void edt_common(int slot, const int *in, float *out, int stride, int ncols, int nrows)
{
    if (nrows <= 0 || ncols <= 0) return;
    b200slam_map *m = slot_map(slot, nrows, ncols);
    int rc = b200slam_map_upload_occupancy(ctx(), m, in, stride);
    if (rc) die("b200slam_map_upload_occupancy", rc);
    rc = b200slam_map_edt(ctx(), m, 10.0f);                     /* MAX_DIST, main.c:224 */
    if (rc) die("b200slam_map_edt", rc);
    rc = b200slam_map_download_field(ctx(), m, out, stride);    /* writes rows x cols only */
    if (rc) die("b200slam_map_download_field", rc);
    g_maps[slot].host_field = out;
}

void euclidean_distance_transform(int input_map[200][200], float output_distance_map[200][200],
                                  int height, int width)
{
    /* argument names as in main_accelerated.c:215; call site passes (nCols, nRows). */
    edt_common(0, &input_map[0][0], &output_distance_map[0][0], 200, height, width);
}

void euclidean_distance_transform2(int input_map[400][400], float output_distance_map[400][400],
                                   int height, int width)
{
    edt_common(1, &input_map[0][0], &output_distance_map[0][0], 400, height, width);
}

static void sync_device_scan(void);

static void fastmatch_common(int slot, const float POSE[3], const float searchResolution[3])
{
    if (!&scan || !&occ_grid || !&FastMatchParameters) {
        fprintf(stderr, "libb200slam_dropin: FastMatch needs the reference globals scan / occ_grid / "
                        "FastMatchParameters (load next to the reference object)\n");
        abort();
    }
    const int rows = slot ? occ_grid.grid_size2[0] : occ_grid.grid_size[0];      /* main.c:388,605 */
    const int cols = slot ? occ_grid.grid_size2[1] : occ_grid.grid_size[1];      /* main.c:389,606 */
    const float *field = slot ? &occ_grid.metric_grid2[0][0] : &occ_grid.metric_grid[0][0];
    const int stride = slot ? 400 : 200;
    b200slam_map *m = slot_map(slot, rows, cols);
    int rc;
    if (g_maps[slot].host_field != field) {       /* field did not come from our transform */
        rc = b200slam_map_upload_field(ctx(), m, field, stride);
        if (rc) die("b200slam_map_upload_field", rc);
        g_maps[slot].host_field = field;
    }
    rc = b200slam_map_set_geometry(m, slot ? occ_grid.pixel_size2 : occ_grid.pixel_size,
                                   slot ? occ_grid.top_left_corner2[0] : occ_grid.top_left_corner[0],
                                   slot ? occ_grid.top_left_corner2[1] : occ_grid.top_left_corner[1]);
    if (rc) die("b200slam_map_set_geometry", rc);
    /* main.c:417-421.  FastMatch and FastMatch2 of one iteration see the same scan: upload once. */
    sync_device_scan();
    rc = b200slam_fastmatch(ctx(), m, POSE, searchResolution, FastMatchParameters.pose,
                            FastMatchParameters.bestHits, &FastMatchParameters.bestHits_size);
    if (rc) die("b200slam_fastmatch", rc);
}

void FastMatch(const float POSE[3], const float searchResolution[3])
{
    fastmatch_common(0, POSE, searchResolution);
}

void FastMatch2(const float POSE[3], const float searchResolution[3])
{
    fastmatch_common(1, POSE, searchResolution);
}

/* main.c:271-363: both levels rasterised and transformed on the device.  Everything the
 * reference function leaves behind is reproduced: grid / grid2 (zeroed, then the 1s),
 * grid_size, metric grids, pixel sizes, top-left corners. */
static void occgrid_level(int slot, float pixel)
{
    const int S = slot ? 400 : 200;
    int rows = 0, cols = 0;
    float tl[2];
    if (!g_maps[slot].map) slot_map(slot, S, S);
    int rc;
    if (g_local_on_device) {          /* ExtractLocalMap left the very same points on the device */
        rc = b200slam_map_rasterise_local(ctx(), g_maps[slot].map, pixel, &rows, &cols, tl);
        if (rc) die("b200slam_map_rasterise_local", rc);
    } else {
        rc = b200slam_map_rasterise(ctx(), g_maps[slot].map, local_map.x, local_map.y, local_map.size, pixel,
                                    &rows, &cols, tl);
        if (rc) die("b200slam_map_rasterise", rc);
    }
    g_maps[slot].rows = rows;
    g_maps[slot].cols = cols;
    rc = b200slam_map_edt(ctx(), g_maps[slot].map, 10.0f);                  /* main.c:355-356 */
    if (rc) die("b200slam_map_edt", rc);
    int *grid = slot ? &occ_grid.grid2[0][0] : &occ_grid.grid[0][0];
    float *field = slot ? &occ_grid.metric_grid2[0][0] : &occ_grid.metric_grid[0][0];
    memset(grid, 0, sizeof(int) * (size_t)S * S);                            /* main.c:319-320 */
    rc = b200slam_map_download_occupancy(ctx(), g_maps[slot].map, grid, S);
    if (rc) die("b200slam_map_download_occupancy", rc);
    rc = b200slam_map_download_field(ctx(), g_maps[slot].map, field, S);
    if (rc) die("b200slam_map_download_field", rc);
    g_maps[slot].host_field = field;
    if (slot) {
        occ_grid.grid_size2[0] = rows; occ_grid.grid_size2[1] = cols;       /* main.c:316-317 */
        occ_grid.pixel_size2 = pixel;
        occ_grid.top_left_corner2[0] = tl[0]; occ_grid.top_left_corner2[1] = tl[1];
    } else {
        occ_grid.grid_size[0] = rows; occ_grid.grid_size[1] = cols;         /* main.c:313-314 */
        occ_grid.pixel_size = pixel;
        occ_grid.top_left_corner[0] = tl[0]; occ_grid.top_left_corner[1] = tl[1];
    }
}

void OccupationalGrid(const float PIXELSIZE, const float PIXELSIZE2)
{
    if (!&local_map || !&occ_grid) {
        fprintf(stderr, "libb200slam_dropin: OccupationalGrid needs the reference globals local_map / occ_grid\n");
        abort();
    }
    occgrid_level(0, PIXELSIZE);
    occgrid_level(1, PIXELSIZE2);
}

/* ---- the steps in front of OccupationalGrid --------------------------- */

/* main.c:71-95.  Reads test_input_memory / lidar, leaves scan.x / scan.y / scan.size. */
void readAScan(const int usableRange)
{
    float *volatile raw = test_input_memory;                        /* weak: NULL when the reference is not loaded */
    if (!&scan || !&lidar || !raw) {
        fprintf(stderr, "libb200slam_dropin: readAScan needs the reference globals scan / lidar / test_input_memory\n");
        abort();
    }
    int rc;
    if (!g_lidar_set || memcmp(g_lidar_angles, lidar.angles, sizeof g_lidar_angles)) {
        static float ca[REF_COLUMN], sa[REF_COLUMN];
        for (int i = 0; i < REF_COLUMN; i++) {                       /* main.c:90-91 */
            ca[i] = cosf(lidar.angles[i]);
            sa[i] = sinf(lidar.angles[i]);
        }
        rc = b200slam_lidar_set(ctx(), ca, sa, REF_COLUMN, lidar.range_min);
        if (rc) die("b200slam_lidar_set", rc);
        memcpy(g_lidar_angles, lidar.angles, sizeof g_lidar_angles);
        g_lidar_set = 1;
    }
    rc = b200slam_scan_read(ctx(), raw, usableRange, &scan.size);
    if (rc) die("b200slam_scan_read", rc);
    rc = b200slam_scan_download(ctx(), scan.x, scan.y, NULL, NULL, NULL);
    if (rc) die("b200slam_scan_download", rc);
    memcpy(g_scan_x, scan.x, sizeof(float) * (size_t)scan.size);    /* the device holds exactly this scan */
    memcpy(g_scan_y, scan.y, sizeof(float) * (size_t)scan.size);
    g_scan_size = scan.size;
}

static void sync_device_scan(void)
{
    if (g_scan_size != scan.size || memcmp(g_scan_x, scan.x, sizeof(float) * (size_t)scan.size) ||
        memcmp(g_scan_y, scan.y, sizeof(float) * (size_t)scan.size)) {
        int rc = b200slam_scan_upload(ctx(), scan.x, scan.y, scan.size);
        if (rc) die("b200slam_scan_upload", rc);
        memcpy(g_scan_x, scan.x, sizeof(float) * (size_t)scan.size);
        memcpy(g_scan_y, scan.y, sizeof(float) * (size_t)scan.size);
        g_scan_size = scan.size;
    }
}

/* main.c:97-118.  Leaves scan.tx / scan.ty. */
void Transform(const float POSE[3])
{
    if (!&scan) {
        fprintf(stderr, "libb200slam_dropin: Transform needs the reference global scan\n");
        abort();
    }
    sync_device_scan();
    int rc = b200slam_scan_transform(ctx(), POSE);
    if (rc) die("b200slam_scan_transform", rc);
    rc = b200slam_scan_download(ctx(), NULL, NULL, scan.tx, scan.ty, NULL);
    if (rc) die("b200slam_scan_download", rc);
}

/* main.c:155-198.  Reads scan.tx / scan.ty (as the last Transform left them on the device) and map,
 * leaves local_map.  main() appends to map.x / map.y itself (main.c:942-948): only what changed
 * since the last call is uploaded. */
void ExtractLocalMap(const float BORDERSIZE)
{
    if (!&scan || !&map || !&local_map) {
        fprintf(stderr, "libb200slam_dropin: ExtractLocalMap needs the reference globals scan / map / local_map\n");
        abort();
    }
    int same = 0;                                  /* leading points the device already holds */
    const int lim = g_map_size < map.size ? g_map_size : map.size;
    while (same < lim && g_map_x[same] == map.x[same] && g_map_y[same] == map.y[same]) same++;
    int rc = b200slam_mappoints_upload(ctx(), map.x + same, map.y + same, map.size - same, same);
    if (rc) die("b200slam_mappoints_upload", rc);
    memcpy(g_map_x + same, map.x + same, sizeof(float) * (size_t)(map.size - same));
    memcpy(g_map_y + same, map.y + same, sizeof(float) * (size_t)(map.size - same));
    g_map_size = map.size;
    rc = b200slam_local_map_extract(ctx(), BORDERSIZE, &local_map.size);
    if (rc) die("b200slam_local_map_extract", rc);
    rc = b200slam_local_map_download(ctx(), local_map.x, local_map.y, NULL);
    if (rc) die("b200slam_local_map_download", rc);
    g_local_on_device = 1;
}

/* The host field array may be rewritten by someone other than our transform (tests do);
 * callers can force a re-upload on the next FastMatch. */
void b200slam_dropin_invalidate(void)
{
    g_maps[0].host_field = NULL;
    g_maps[1].host_field = NULL;
    g_local_on_device = 0;
}
