/*
 * b200slam.h -- C ABI of libb200slam.so, the B200 (sm_100a) implementation of the
 * lidar-SLAM hot path of circuitpotato/Hardware-Acceleration-of-LIDAR-SLAM:
 * the clamped occupancy-grid Euclidean distance transform and the scan-matching /
 * particle-weighting loop.
 *
 * Plain C: opaque handles, pointers and sizes only; no C++/CUDA/torch types cross this
 * boundary.  Every function returns B200SLAM_OK (0) or a negative error code and never
 * throws; b200slam_last_error() gives the message.  There is NO CPU fallback: without a
 * CUDA device b200slam_create() fails and nothing else can be called.
 *
 * Citations `file:line` are to the reference tree; they name the reference interface the
 * entry point replaces.  The reference-named drop-in wrappers
 *   euclidean_distance_transform / euclidean_distance_transform2 / FastMatch / FastMatch2
 * live in libb200slam_dropin.so (host/dropin.c) on top of this ABI -- see INTEGRATION.md.
 *
 * Conventions (reference): grid[row = y][col = x], row-major;
 * strides are in ELEMENTS; poses are {x, y, theta} in metres / radians, theta applied as
 * R(-theta) (Subsystem_1/main.c:462-463); candidate order is theta slowest, then x, then y
 * (main.c:443,468,487) and ties go to the lowest linear index (strict `<`, main.c:549).
 */
#ifndef B200SLAM_H
#define B200SLAM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SLAM_OK          0
#define B200SLAM_ERR_ARG    (-1)   /* bad argument / unsupported size            */
#define B200SLAM_ERR_CUDA   (-2)   /* CUDA runtime error (message has the cause) */
#define B200SLAM_ERR_NCCL   (-3)   /* NCCL missing or failed                     */
#define B200SLAM_ERR_NOMEM  (-4)
#define B200SLAM_ERR_STATE  (-5)   /* call order (e.g. no scan uploaded)         */

#define B200SLAM_ABI_VERSION 1

typedef struct b200slam_ctx b200slam_ctx;   /* one per process per GPU; not thread-safe,   */
typedef struct b200slam_map b200slam_map;   /* like the reference's global scratch state   */

/* Result of a match.  Mirrors what FastMatch leaves behind in FastMatchParameters
 * (main.c:374-379, 554-557, 592-594). */
typedef struct {
    int64_t best_index;    /* global linear candidate index of the winner             */
    float   best_score;    /* sum of distance-field values over in-bounds beams        */
    float   best_pose[3];  /* {x, y, theta} of the winner                              */
    int32_t best_hits;     /* number of in-bounds beams of the winner (bestHits_size)  */
    int32_t last_hits;     /* number of in-bounds beams of the LAST candidate scored   */
} b200slam_match;

/* ---- context ------------------------------------------------------------------- */

int  b200slam_abi_version(void);
/* Creates the context on CUDA device `device` (one stream, scratch buffers).  Fails with
 * B200SLAM_ERR_CUDA when there is no usable GPU. */
int  b200slam_create(b200slam_ctx **out, int device);
void b200slam_destroy(b200slam_ctx *ctx);
/* Message of the last failure on this context (ctx == NULL: last create failure). */
const char *b200slam_last_error(const b200slam_ctx *ctx);
/* Blocks until everything queued on the context's stream has finished.  Every device-side wait of the
 * library (peer exchange, peer barrier, TMA barrier) is bounded -- B200SLAM_SPIN_TIMEOUT_MS in the
 * environment, default 30 000 -- and one that gave up (a peer died, or the ranks queued their collective
 * calls in different orders) leaves a sticky error: this call and b200slam_match_fetch then return
 * B200SLAM_ERR_STATE until b200slam_comm_init is called again.  The GPU is never left spinning. */
int  b200slam_sync(b200slam_ctx *ctx);
/* The context's cudaStream_t, for callers that time with CUDA events or interoperate. */
void *b200slam_stream(b200slam_ctx *ctx);
/* Number of kernels this library has launched on the context since creation. */
uint64_t b200slam_launch_count(const b200slam_ctx *ctx);
/* name: >= 64 bytes. */
int  b200slam_device_info(const b200slam_ctx *ctx, char *name, int *sm_count, int *cc_major,
                          int *cc_minor, size_t *total_mem);
/* Pinned host memory for callers that want full-speed H2D/D2H through the host calls. */
int  b200slam_host_alloc(b200slam_ctx *ctx, size_t bytes, void **out);
int  b200slam_host_free(b200slam_ctx *ctx, void *p);

/* ---- distance transform ---------------------------------------------------------
 * Replaces euclidean_distance_transform / euclidean_distance_transform2
 * (Subsystem_1/main.c:223-269, Subsystem_1/main_accelerated.c:215-283,
 *  Submodule_2/Accelereated_Euclidean_Distance_Transform.c:1-69) and the FPGA register call
 * euclidean_distance_compute (Submodule_2/Hadrware_acclereated.cpp:223-233):
 *     out[r][c] = d2min < max_dist^2 ? sqrtf(d2min) : max_dist
 * with d2min the integer squared distance to the nearest non-zero cell inside the
 * rows x cols sub-rectangle.  Bit-exact with the reference. */

/* One-shot host call: H2D occupancy, transform on the GPU, D2H field.  Writes only
 * out[0..rows)[0..cols) (main.c:225-226). */
int b200slam_edt(b200slam_ctx *ctx, const int32_t *occ, int occ_stride, float *out,
                 int out_stride, int rows, int cols, float max_dist);

/* Device-resident map (the MyGrid container, main.c:200-213: grid + metric_grid +
 * grid_size + pixel_size + top_left_corner). */
int  b200slam_map_create(b200slam_ctx *ctx, int rows, int cols, b200slam_map **out);
void b200slam_map_destroy(b200slam_ctx *ctx, b200slam_map *map);
/* Changes the current size within the capacity the map was created with (the reference keeps
 * fixed 200 x 200 / 400 x 400 arrays and a varying grid_size, main.c:200-213). */
int  b200slam_map_resize(b200slam_map *map, int rows, int cols);
/* pixel_size / top_left_corner = {minX, minY} (main.c:357-362). */
int  b200slam_map_set_geometry(b200slam_map *map, float pixel_size, float top_left_x,
                               float top_left_y);
int  b200slam_map_upload_occupancy(b200slam_ctx *ctx, b200slam_map *map, const int32_t *occ,
                                   int stride);
/* Rasterises map points into the occupancy grid ON THE DEVICE: one level of OccupationalGrid
 * (Subsystem_1/main.c:271-354) -- bounding box of the points, 3-pixel margin, grid size
 * round(extent / pixel) + 1, a 1 in every cell a point rounds into -- with the reference's float
 * arithmetic (the bounding box and sizes are computed on the host exactly as main.c:272-305,
 * the per-point cell indices on the GPU with IEEE division and roundf, main.c:332-353).
 * The map was created with capacity rows x cols (200 x 200 / 400 x 400 in the reference); its
 * current size, pixel size and top-left corner become those of the rasterised grid
 * (grid_size, pixel_size, top_left_corner: main.c:313-314, 357-362), returned through
 * rows / cols / top_left (each optional).  B200SLAM_ERR_ARG when the grid exceeds the capacity.
 * Only 8 bytes per point cross PCIe instead of 4 bytes per cell.  x, y: host arrays. */
int  b200slam_map_rasterise(b200slam_ctx *ctx, b200slam_map *map, const float *x, const float *y,
                            int npoints, float pixel_size, int *rows, int *cols, float top_left[2]);
/* The same with x and y in PAGE-LOCKED host memory (b200slam_host_alloc): the copies are queued straight
 * from the caller's arrays -- no staging copy, no wait on the stream -- so the call returns as soon as the
 * bounding box (the only host pass over the points) is known; x and y must stay unchanged until the next
 * synchronising call on this context (b200slam_sync, a fetch).  B200SLAM_ERR_ARG for pageable arrays.
 * With b200slam_map_edt and b200slam_score_lattice_async behind it this is the reference's own sequence
 * OccupationalGrid -> euclidean_distance_transform -> FastMatch (main.c:884-918) with 8 bytes per map
 * point crossing PCIe instead of 4 bytes per grid cell. */
int  b200slam_map_rasterise_async(b200slam_ctx *ctx, b200slam_map *map, const float *x, const float *y,
                                  int npoints, float pixel_size, int *rows, int *cols, float top_left[2]);
/* occ -> field entirely on the device (async on the context's stream). */
int  b200slam_map_edt(b200slam_ctx *ctx, b200slam_map *map, float max_dist);
int  b200slam_map_download_field(b200slam_ctx *ctx, b200slam_map *map, float *out, int stride);
/* The int32 occupancy of the current rows x cols (e.g. after b200slam_map_rasterise). */
int  b200slam_map_download_occupancy(b200slam_ctx *ctx, b200slam_map *map, int32_t *out, int stride);
/* Install a precomputed distance field (e.g. one produced by another GPU). */
int  b200slam_map_upload_field(b200slam_ctx *ctx, b200slam_map *map, const float *field,
                               int stride);
/* Raw device pointers and pitches (elements), for zero-copy producers / consumers. */
int  b200slam_map_device_ptrs(b200slam_map *map, int32_t **occ, int *occ_pitch, float **field,
                              int *field_pitch);

/* ---- scan matching ---------------------------------------------------------------
 * Replaces the candidate-scoring loop of FastMatch / FastMatch2 (main.c:381-596, 598-809). */

/* Sensor-frame scan points (ScanData.x/.y/.size, main.c:60-69), kept on the device until
 * replaced. */
int b200slam_scan_upload(b200slam_ctx *ctx, const float *x, const float *y, int nbeams);

/* Axis value k of an n-point lattice axis centred on p: p + (float)(k - n/2) * s
 * (n == 3 gives {p - s, p, p + s}, main.c:424-426, bit for bit). */
float b200slam_lattice_value(float p, float s, int k, int n);

/* Scores every candidate of the lattice {theta} x {tx} x {ty} around pose0 against `map`
 * and returns the arg-min (lowest score, then lowest linear index).
 *   step = {tx step, ty step, theta step}   (searchResolution, main.c:386-387)
 *   n    = {n_theta, n_tx, n_ty}
 *   scores          optional host [n0*n1*n2]: every candidate's score
 *   last_hit_values optional host [nbeams]: in-bounds field values of the LAST candidate
 *                   (what FastMatchParameters.bestHits holds after the call, main.c:515)
 * cos/sin of the lattice angles come from the host libm, exactly as main.c:434-435. */
int b200slam_score_lattice(b200slam_ctx *ctx, b200slam_map *map, const float pose0[3],
                           const float step[3], const int n[3], float *scores,
                           float *last_hit_values, b200slam_match *result);

/* Same, restricted to the theta-major rows [row_begin, row_end) of the n0*n1 (theta, tx)
 * rows -- the unit candidates are sharded by across GPUs.  `result` is the shard-local
 * arg-min with GLOBAL best_index.  With a communicator (b200slam_comm_init) and
 * allreduce != 0 the per-rank bests are all-gathered and every rank returns the global
 * winner. */
int b200slam_score_lattice_rows(b200slam_ctx *ctx, b200slam_map *map, const float pose0[3],
                                const float step[3], const int n[3], int64_t row_begin,
                                int64_t row_end, int allreduce, b200slam_match *result);

/* Queues the same lattice match on the context's stream WITHOUT reading anything back
 * (for device-timed loops and CUDA-graph capture); fetch with b200slam_match_fetch.
 * allreduce: 0 = this rank's shard only; 1 = exchange and merge with the other ranks;
 * 2 = POST this rank's result to its peers (NVLink peer memory, fire and forget) and merge the
 * PREVIOUS post of a sequence of such calls in the same kernel tail -- the ranks drift by up to a
 * step instead of running in lockstep; b200slam_exchange_collect_async after the last call of
 * the sequence merges what is still pending, and b200slam_match_fetch then returns the global
 * result of that last match.  Posts and collects pair up in order on every rank.
 * 3 = DEFERRED: the kernel only records this rank's result on its own GPU, so a burst of
 * independent matches runs at single-GPU speed (no kernel has NVLink stores in flight when it
 * completes, nothing waits for a peer); b200slam_exchange_collect_async must follow after at
 * most 31 such calls: it sends the whole burst to the peers, merges every rank's results in
 * order, and the result of the LAST match is what b200slam_match_fetch returns.  Without NVLink
 * peer memory (NCCL fallback) 2 and 3 act like 1. */
int b200slam_score_lattice_async(b200slam_ctx *ctx, b200slam_map *map, const float pose0[3],
                                 const float step[3], const int n[3], int64_t row_begin,
                                 int64_t row_end, int allreduce);
int b200slam_exchange_collect_async(b200slam_ctx *ctx);
int b200slam_match_fetch(b200slam_ctx *ctx, b200slam_match *result);
/* The device's twin of FastMatchParameters.bestHits[] (main.c:376) as the matches queued so far have left
 * it: hits[0 .. count) (count is clipped to the buffer, >= 2500 floats).  For callers of the *_async
 * entry points; b200slam_fastmatch returns the same data through its best_hits argument. */
int b200slam_match_fetch_hits(b200slam_ctx *ctx, float *hits, int count);

/* Arbitrary pose / particle list: poses[P][3] = {x, y, theta}; ct/st optional [P]
 * (cosf/sinf(theta) from the host libm when NULL).  scores (optional host [P]) and
 * hits (optional host [P]) are copied back when given; the scores also stay on the device
 * for b200slam_weights_resample.  index_base is added to best_index (shard offset). */
int b200slam_score_poses(b200slam_ctx *ctx, b200slam_map *map, const float *poses,
                         const float *ct, const float *st, int64_t P, int64_t index_base,
                         float *scores, int32_t *hits, b200slam_match *result);

/* The reference call itself: 3x3x3 lattice around `pose` with t = search_resolution[0],
 * r = search_resolution[2]; pose_out / best_hits / best_hits_size are exactly what
 * FastMatch leaves in FastMatchParameters (main.c:374-379): the winner's pose, the winner's
 * hit count, and bestHits[] as the reference's loop leaves it.  Every candidate overwrites
 * that array from index 0 in loop order (main.c:515) and nothing ever clears it, so pass the
 * SAME best_hits array (at least as long as the scan, e.g. the reference's own
 * FastMatchParameters.bestHits[2500]) to every call: the call rewrites the last candidate's
 * hit values, behind them those of the most recent candidate that had more, and leaves the rest
 * as earlier calls left it -- which is what main.c:942-948 reads when the winner has more hits
 * than the last candidate.  (The device keeps its own copy of the array across calls;
 * b200slam_mappoints_grow reads that one.)  best_hits / best_hits_size may be NULL. */
int b200slam_fastmatch(b200slam_ctx *ctx, b200slam_map *map, const float pose[3],
                       const float search_resolution[3], float pose_out[3], float *best_hits,
                       int *best_hits_size);

/* ---- scan front end and map points on the device (the steps either side of the hot path) ----
 * With these the per-scan loop of the reference's main() (Subsystem_1/main.c:859-970) keeps the
 * scan, the map points, the local map, both grids and both distance fields on the device: 4
 * bytes per beam go up per scan and a pose comes back.  All arithmetic is the reference's
 * (products and sums rounded separately, strict compares, compaction in input order); cos / sin
 * of the beam angles and of the pose come from the HOST libm, like the lattice tables.
 *
 *   b200slam_lidar_set          SetLidarParameters, main.c:45-58: cos_a[i] = cosf(angles[i]),
 *                               sin_a[i] = sinf(angles[i]) computed by the caller; range_min
 *   b200slam_scan_read          readAScan, main.c:71-95: drop r < range_min | r > max_range,
 *                               x = r * cos, y = r * sin, compacted in beam order into the context's
 *                               scan (what b200slam_scan_upload would have received); *size = scan.size
 *   b200slam_scan_transform     Transform, main.c:97-118: world-frame tx / ty for POSE
 *   b200slam_scan_download      any of x, y, tx, ty (each optional) and the size
 *   b200slam_mappoints_upload   MapPoints map (main.c:121-131): points [offset, offset + n), offset <=
 *                               current size; the size becomes offset + n
 *   b200slam_mappoints_from_scan  Initialise, main.c:136-145: map <- world-frame scan
 *   b200slam_mappoints_grow     main.c:942-948: for j < bestHits_size (of the winner of the LAST
 *                               match): bestHits[j] (hits of the last candidate scored) > threshold ->
 *                               append scan.tx[j], scan.ty[j]; *added = newPointSize
 *   b200slam_mappoints_download
 *   b200slam_local_map_extract  ExtractLocalMap, main.c:155-198: bounding box of the world-frame scan
 *                               +- border, map points strictly inside, order kept; *size = local_map.size
 *   b200slam_local_map_download
 *   b200slam_map_rasterise_local  one level of OccupationalGrid (main.c:271-354) from the resident local
 *                               map: b200slam_map_rasterise without the points crossing PCIe */
int b200slam_lidar_set(b200slam_ctx *ctx, const float *cos_a, const float *sin_a, int nbeams, float range_min);
int b200slam_scan_read(b200slam_ctx *ctx, const float *ranges, int max_range, int *size);
int b200slam_scan_transform(b200slam_ctx *ctx, const float pose[3]);
int b200slam_scan_download(b200slam_ctx *ctx, float *x, float *y, float *tx, float *ty, int *size);
int b200slam_mappoints_upload(b200slam_ctx *ctx, const float *x, const float *y, int n, int offset);
int b200slam_mappoints_from_scan(b200slam_ctx *ctx);
int b200slam_mappoints_grow(b200slam_ctx *ctx, float threshold, int *added);
int b200slam_mappoints_download(b200slam_ctx *ctx, float *x, float *y, int *size);
int b200slam_local_map_extract(b200slam_ctx *ctx, float border, int *size);
int b200slam_local_map_download(b200slam_ctx *ctx, float *x, float *y, int *size);
int b200slam_map_rasterise_local(b200slam_ctx *ctx, b200slam_map *map, float pixel_size, int *rows, int *cols,
                                 float top_left[2]);

/* ---- the per-scan loop without host round trips -------------------------
 * The reference's loop (main.c:859-970) needs the host only where glibc's cosf / sinf define the bits: the lattice
 * tables of the two matches and Transform.  Everything else can stay queued:
 *
 *   b200slam_scan_read_async        readAScan without reading scan.size back: the size stays on the device
 *                                   and every kernel that needs it (Transform, ExtractLocalMap, the matchers)
 *                                   reads it there; the host keeps an upper bound until a call returns it
 *   b200slam_scan_read_resident_async  the same from the values of b200slam_csv_ingest (nothing crosses PCIe)
 *   b200slam_fastmatch_pair_async   FastMatch on map_a around `pose` with res_a, then FastMatch2 on map_b
 *                                   around FastMatch's result with res_b (main.c:902-918), as two kernels
 *                                   with no host step between them: the result of the first is one of
 *                                   3 x 3 x 3 lattice points, so the host sends the second lattice's axis tables
 *                                   for each of the three possible centres per axis (27 cosf / sinf calls,
 *                                   the same float operations as b200slam_fastmatch would do after fetching)
 *                                   and the second kernel picks by the first one's winner.  bestHits[] /
 *                                   bestHits_size are left as after FastMatch2 (what main.c:942-948 reads).
 *   b200slam_fastmatch_pair_fetch   ONE synchronisation per scan: both poses, scan.size, bestHits_size
 *   b200slam_mappoints_grow_async   main.c:942-948 without reading newPointSize back (the map's size stays
 *                                   on the device, like the scan's)
 * b200slam_scan_transform, b200slam_local_map_extract, b200slam_mappoints_download, ... work on either kind
 * of state; the ones that return a size also refresh the host's copy. */
int b200slam_scan_read_async(b200slam_ctx *ctx, const float *ranges, int max_range);
int b200slam_scan_read_resident_async(b200slam_ctx *ctx, int64_t first_value, int max_range);
int b200slam_fastmatch_pair_async(b200slam_ctx *ctx, b200slam_map *map_a, b200slam_map *map_b, const float pose[3],
                                  const float res_a[3], const float res_b[3]);
int b200slam_fastmatch_pair_fetch(b200slam_ctx *ctx, float pose_a[3], float pose_b[3], int *scan_size, int *best_hits_size);
int b200slam_mappoints_grow_async(b200slam_ctx *ctx, float threshold);
/* The whole steady-state scan of the reference's loop as ONE kernel launch: readAScan (main.c:71-95) on `ranges`
 * (host, lidar_n floats; _resident: values [first_value, first_value + lidar_n) of b200slam_csv_ingest), then
 * FastMatch on map_a around `pose` and FastMatch2 on map_b around its result (main.c:902-918), one CTA, the scan
 * and both matches' working set in shared memory.  Equivalent to b200slam_scan_read_async +
 * b200slam_fastmatch_pair_async (same results, same side effects: scan, scan size, bestHits[] twin);
 * b200slam_fastmatch_pair_fetch returns the result.  Scans of up to 1536 beams. */
int b200slam_scan_step_async(b200slam_ctx *ctx, const float *ranges, int max_range, b200slam_map *map_a, b200slam_map *map_b,
                             const float pose[3], const float res_a[3], const float res_b[3]);
int b200slam_scan_step_resident_async(b200slam_ctx *ctx, int64_t first_value, int max_range, b200slam_map *map_a,
                                      b200slam_map *map_b, const float pose[3], const float res_a[3], const float res_b[3]);

/* The per-scan loop ON THE DEVICE (main.c:859-970): scans are queued AHEAD of their results.  The poses the loop
 * carries from one scan to the next -- the current pose, the previous path entry (constant-velocity motion model,
 * main.c:875-898) and map.pose (mini-update test, main.c:928-940) -- live in device memory; the kernel of scan k
 * (readAScan + FastMatch + FastMatch2 fused, as b200slam_scan_step_resident_async) forms pose_guess itself, builds
 * both lattices' axis tables on the device -- cosf / sinf exactly as glibc computes them (csrc/trig.cuh: the libm
 * algorithm restated in double precision, identical on every float of |theta| <= 16; checked against the running
 * libm by b200slam_scan_chain_begin on a sample and by tools/trig_check.cpp exhaustively) -- commits the refined
 * pose in its tail and evaluates the mini-update test.  When the test fires the chain STOPS: the kernels already
 * queued behind that scan return at once, so the scan, bestHits[] and the match state stay those of the scan that
 * needs the map update, and the host does Transform / map growth / the rebuild as before, then begins a new chain.
 *   b200slam_scan_chain_begin       sets the device state; scan_index = the first scan the chain will run;
 *                                   prev_pose NULL: no motion model for that scan (scan_iter == 1, main.c:895-897).
 *                                   B200SLAM_ERR_STATE when the device's cosf / sinf differ from this host's libm
 *                                   (the caller then stays with b200slam_scan_step_resident_async)
 *   b200slam_scan_chain_step_async  queues scans [scan_index, scan_index + nscans) as ONE kernel launch that runs
 *                                   them one after the other while the chain lasts (their ranges: nscans x lidar_n
 *                                   consecutive values of b200slam_csv_ingest from first_value; nscans <= 32);
 *                                   launches may be queued ahead, in order, up to 64 scans in front of the last
 *                                   fetched one; the caller keeps |theta| <= 15 (beyond that: the host-driven calls)
 *   b200slam_scan_chain_fetch       waits for scan scan_index's result in a ring of mapped host memory (64 scans
 *                                   deep): FastMatch's and FastMatch2's poses, scan.size, bestHits_size, and
 *                                   whether the mini-update test fired (*stopped: nothing behind this scan ran). */
int b200slam_scan_chain_begin(b200slam_ctx *ctx, int scan_index, const float pose[3], const float *prev_pose,
                              const float map_pose[3], float mini_update_dt, float mini_update_dr);
int b200slam_scan_chain_step_async(b200slam_ctx *ctx, int scan_index, int nscans, int64_t first_value, int max_range,
                                   b200slam_map *map_a, b200slam_map *map_b, const float res_a[3], const float res_b[3]);
int b200slam_scan_chain_fetch(b200slam_ctx *ctx, int scan_index, float pose_a[3], float pose_b[3], int *scan_size,
                              int *best_hits_size, int *stopped);

/* ---- scan ingest -----------------------------------------------------------
 * Replaces readDatasetLineByLine (Subsystem_1/main.c:22-30): `column` x fscanf(fp, "%f,", &value).  The whole
 * CSV text (values separated by ',' and / or white space) is parsed ON THE GPU: one upload of the raw bytes,
 * then every value is converted in parallel -- bit-identical to glibc's correctly rounded %f conversion (the
 * few tokens a double division cannot settle, and anything that is not a plain decimal, are converted by the
 * host's strtof, the function fscanf itself calls).  values_out (optional, host, max_values floats) receives
 * the values; they also stay on the device for b200slam_scan_read_resident_async.  B200SLAM_ERR_ARG for a
 * token that is not a number. */
int b200slam_csv_ingest(b200slam_ctx *ctx, const char *text, size_t nbytes, float *values_out, int64_t max_values,
                        int64_t *count);
/* Device pointer and count of the ingested values. */
int b200slam_csv_values(b200slam_ctx *ctx, const float **device_values, int64_t *count);

/* ---- scheduling hint -------------------------------------------------------------------
 * (Also read by the distance transform: with many independent transforms in flight it takes taller
 * row chunks per CTA -- less halo work per cell -- where a single transform prefers more, shorter CTAs.)
 * The lattice kernel comes in several tile shapes (candidates per thread x warps per CTA).
 * B200SLAM_MATCH_LATENCY (default): one match at a time, as FastMatch is called per scan
 * (main.c:902-918) -- pick the shape whose busiest SM finishes first.  B200SLAM_MATCH_THROUGHPUT:
 * many independent matches are kept in flight (batched replay, bench.py's pipelined turn) --
 * pick the shape with the fewest L1 wavefronts per evaluation even if one launch alone then
 * occupies only part of the GPU.  Results are identical bit for bit in either mode. */
#define B200SLAM_MATCH_LATENCY    0
#define B200SLAM_MATCH_THROUGHPUT 1
int b200slam_set_match_mode(b200slam_ctx *ctx, int mode);

/* ---- CUDA graphs -----------------------------------------------------------------
 * The replay loop's per-scan work (EDT + two matches, main.c:865-918) is a handful of
 * microsecond-scale kernels; capture the *_async / map_edt calls once and replay them.
 * Between begin and end only b200slam_map_edt and b200slam_score_lattice_async may be
 * called, with scratch already sized by an identical un-captured call. */
typedef struct b200slam_graph b200slam_graph;
int  b200slam_graph_begin(b200slam_ctx *ctx);
int  b200slam_graph_end(b200slam_ctx *ctx, b200slam_graph **out);
int  b200slam_graph_launch(b200slam_ctx *ctx, b200slam_graph *graph);
void b200slam_graph_destroy(b200slam_ctx *ctx, b200slam_graph *graph);

/* ---- device timing ------------------------------------------------------------------
 * CUDA events on the context's stream (the only stream the kernels run on), so callers can
 * time kernels without a CUDA binding of their own.  slot in [0, 4096). */
int b200slam_event_record(b200slam_ctx *ctx, int slot);
/* Makes everything queued on `ctx` from now on wait for event `slot` of `other` (a second
 * context on the same GPU = a second stream): cross-stream dependencies for callers that
 * pipeline independent work, e.g. the transform of the next map under the match on the current
 * one.  Usable inside b200slam_graph_begin/end of `ctx` (the other stream joins the capture). */
int b200slam_event_wait(b200slam_ctx *ctx, b200slam_ctx *other, int slot);
int b200slam_event_elapsed_ms(b200slam_ctx *ctx, int slot_start, int slot_stop, float *ms);

/* ---- particle weights + systematic resampling (extension; not in the reference) ----
 * Uses the scores of the most recent b200slam_score_poses call (device-resident):
 *   w_i = exp_det(-beta * (score_i - score_min)),  q_i = (uint64)(w_i * 2^32),
 *   W = sum q_i,  weights[i] = (float)((double)q_i / W),
 *   T_k = U + floor(k*W/N),  U = ((W div N) * u0_q32) >> 32,
 *   ancestors[k] = first i with inclusive prefix sum of q > T_k.
 * Integer weights make the prefix sum associative, so the indices are identical on the
 * CPU oracle, one GPU and eight GPUs.  With a communicator, score_min and W are global
 * and ancestors holds this rank's slice [k_begin, k_end) of the N_global slots
 * (k_begin/k_count returned), with GLOBAL ancestor indices. */
int b200slam_weights_resample(b200slam_ctx *ctx, float beta, uint32_t u0_q32, float *weights,
                              uint64_t *wsum, int32_t *ancestors, int64_t *k_begin,
                              int64_t *k_count);

/* Device-resident particle set: the particles stay in HBM across filter steps,
 * nothing crosses PCIe per step, and every call below only queues kernels on the context's
 * stream (capturable in a CUDA graph).
 *   upload            poses[P][3] (+ optional ct/st, else host libm) -> device SoA
 *   score_async       scores / hit counts of the resident particles against `map`; the arg-min
 *                     is available through b200slam_match_fetch
 *   resample_async    weights, normalisation, systematic resampling (same definition as
 *                     b200slam_weights_resample) and the gather: the resident set is REPLACED by
 *                     its resampled offspring (particle k <- old particle ancestors[k])
 *   download          current poses and, optionally, the weights / ancestors of the last resample */
int b200slam_particles_upload(b200slam_ctx *ctx, const float *poses, const float *ct, const float *st,
                              int64_t P);
/* SHARDED resident particle set (several GPUs; needs b200slam_comm_init with NVLink peer memory).
 * Collective: rank r hands in its slice poses[P][3] = particles [index_base, index_base + P) of n_global
 * (every rank the same P: a rank addresses its peers' buffers with its own layout).  Afterwards
 * b200slam_particles_score_async / _resample_async run the SAME filter step as on one GPU over the whole
 * set: the score kernel's tail sends {best (score, global index), P} to every rank, the weight kernels
 * wait for them (global minimum score), the prefix-scan kernel exchanges the integer weight sums and
 * derives on the device the global W, this rank's offset and the systematic-resampling slots whose
 * ancestors live here (b200slam_resample_owned_slots' arithmetic), and the resampling kernel stores each
 * offspring (pose + global ancestor index) straight into the buffers of the rank that holds its slot --
 * rank r ends with slots [index_base, index_base + P) of the global offspring, exactly the slice an
 * unsharded run would have there.  Everything travels through NVLink peer memory written by the kernels
 * themselves: no NCCL call, no host synchronisation, CUDA-graph capturable; every wait is bounded
 * (B200SLAM_SPIN_TIMEOUT_MS).  All ranks must queue the same sequence of steps.
 * b200slam_particles_download then returns this rank's slice (ancestors are GLOBAL indices) and
 * b200slam_match_fetch this rank's best particle. */
int b200slam_particles_shard(b200slam_ctx *ctx, const float *poses, const float *ct, const float *st, int64_t P,
                             int64_t index_base, int64_t n_global);
int b200slam_particles_score_async(b200slam_ctx *ctx, b200slam_map *map);
int b200slam_particles_resample_async(b200slam_ctx *ctx, float beta, uint32_t u0_q32);
int b200slam_particles_download(b200slam_ctx *ctx, float *poses, float *scores, float *weights,
                                int32_t *ancestors);

/* ---- multi-resolution match (generalises the 2-level schedule of main.c:901-918) ---- */
int b200slam_pyramid_match(b200slam_ctx *ctx, b200slam_map *const *maps, int levels,
                           const float pose0[3], const float *steps /*[levels][3]*/,
                           const int *n /*[levels][3]*/, b200slam_match *results /*[levels]*/);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ------------------------------
 * Only per-shard bests (8 B/rank) and weight sums (16 B/rank) are exchanged. */
#define B200SLAM_UNIQUE_ID_BYTES 128
int b200slam_comm_unique_id(void *id_out /*[128]*/);            /* rank 0 */
int b200slam_comm_init(b200slam_ctx *ctx, int nranks, int rank, const void *id /*[128]*/);
int b200slam_comm_destroy(b200slam_ctx *ctx);
/* Device-side barrier over the ranks, queued on the context's stream (nothing blocks on the host): the
 * work queued behind it starts only when every rank's stream has reached its own barrier.  Lines the GPUs
 * up before a timed region; the row-sharded transform uses the same kernel internally. */
int b200slam_comm_barrier_async(b200slam_ctx *ctx);
/* Row-sharded transform (BASELINE configs[3]: "8192x8192 grid EDT (row-sharded)").  The
 * transform is clamped at max_dist, so a block of output rows needs only a ceil(max_dist)-1 row
 * halo of the INPUT, which every rank already holds (the occupancy is replicated): nothing is
 * exchanged to compute.  What has to travel is the OUTPUT, because every rank scores against the
 * whole field:
 *   b200slam_map_edt_rows     this GPU only, output rows [row_begin, row_end) (also the cheap way
 *                             to refresh the field after a local change of the occupancy);
 *   b200slam_map_edt_sharded  every rank transforms its block b200slam_shard_range(rows, nranks,
 *                             rank) and the blocks are exchanged so that all ranks end with the
 *                             whole field.  B200SLAM_EDT_GATHER_NCCL: in-place ncclAllGather (equal
 *                             blocks) or grouped ncclBroadcasts.  B200SLAM_EDT_GATHER_P2P: the EDT
 *                             kernel itself stores every row it produces into every peer's field
 *                             over NVLink (fields mapped through CUDA IPC by b200slam_map_share,
 *                             <= 8 ranks), bracketed by two device-side barriers -- compute and
 *                             all-gather are one kernel.
 * Collective calls: every rank must make them in the same order with identically sized maps.
 * Without a communicator b200slam_map_edt_sharded is b200slam_map_edt. */
#define B200SLAM_EDT_GATHER_NCCL 0
#define B200SLAM_EDT_GATHER_P2P  1
int b200slam_map_edt_rows(b200slam_ctx *ctx, b200slam_map *map, float max_dist, int row_begin, int row_end);
int b200slam_map_share(b200slam_ctx *ctx, b200slam_map *map);
int b200slam_map_edt_sharded(b200slam_ctx *ctx, b200slam_map *map, float max_dist, int mode);

/* Even split of `total` units over nranks (pure host arithmetic). */
void b200slam_shard_range(int64_t total, int nranks, int rank, int64_t *begin, int64_t *end);
/* Systematic-resampling slots owned by a rank (pure host): the k in [0, n_global) whose
 * threshold T_k = U + floor(k * w_global / n_global) lies in [rank_offset, rank_offset +
 * w_local), rank_offset being the sum of the lower ranks' integer weight sums. */
void b200slam_resample_owned_slots(uint64_t w_global, int64_t n_global, uint32_t u0_q32,
                                   uint64_t rank_offset, uint64_t w_local, int64_t *k_begin,
                                   int64_t *k_count);
/* Packed (score, index) key used for the arg-min exchange and its merge (pure host). */
uint64_t b200slam_pack_key(float score, uint32_t index);
void     b200slam_unpack_key(uint64_t key, float *score, uint32_t *index);
uint64_t b200slam_merge_keys(const uint64_t *keys, int n);

#ifdef __cplusplus
}
#endif
#endif /* B200SLAM_H */
