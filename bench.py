#!/usr/bin/env python
"""bench.py -- pose x beam evals/s and EDT Mcells/s of the b200slam hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config1|config2|config3|config4|tiny]
    python bench.py --impl reference ...        # the reference's own CPU code, same metric

A "step" is one pass of the hot path over one batch of synthetic input: the clamped EDT of
the occupancy grid followed by the correlative scan match of every lattice candidate against
the fresh distance field, ending in the arg-min.  Default workload: BASELINE.json configs[1]
(2048x2048 grid, 64 x 32 x 32 = 65 536 candidate poses x 360 beams, the single-GPU
configuration); the same line carries under "largest_sweep" the same measurements on
configs[3] (8192x8192 grid, 256 x 128 x 128 = 4 194 304 poses x 1080 beams per GPU), the sweep
north_star states its roofline and scaling targets on.  --workload config2 / config4 are the
particle-filter step (configs[2]) and the 3-level pyramid search (configs[4]).

Own arm, per rank (one process per GPU; torch.distributed only for barrier / max-over-ranks):
  * `value`  : whole-job pose x beam evals / s with inputs resident in HBM: K steps replayed
               back to back as CUDA graphs (the step is a handful of microsecond kernels),
               bracketed by barrier + sync, timed with CUDA events on the library's stream,
               max over ranks.  Steps cycle through a ring of distinct maps larger than L2.
               Many independent steps are in flight, so the matcher runs with the library's
               throughput tile-shape policy (b200slam_set_match_mode); `serial_ms_per_step` is
               the same K steps strictly one kernel after the other with the default policy.
  * per-kernel durations (eager pass, CUDA events around every kernel) -> `roofline`,
               `rooflines` (incl. the matcher under the default, one-match-at-a-time policy)
  * `e2e`    : the same step through the host-buffer C ABI calls: H2D of the int32 grid and
               the scan from pinned memory and D2H of the match result inside the timed region
  * `cpu_baseline` (rank 0, N == 1): the reference's own EDT2 + FastMatch2 (oracle/_ref) or the
               oracle port, 1 core, on a bounded sample of the same workload
N > 1: weak scaling -- every rank keeps the per-GPU workload (map replicated, its own block
of theta rows of an N-times larger lattice) and the per-rank bests are all-gathered with
NCCL inside the step (through NVLink peer memory in the kernel's tail where CUDA IPC works).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "hardware-acceleration-of-lidar-slam_b200"

METRIC = "pose_x_beam_evals_per_s"
UNIT = "evals/s"
L2_BYTES = 126 * 1024 * 1024


# ----------------------------------------------------------------------------------------
def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full summary, or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = the upper half of the samples (idle samples bracket the region)
        load = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(smax), "samples": len(sm),
                "power_w_max": max(power), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
def _cpu_worker(job):
    """One host core: the reference's own euclidean_distance_transform2 + FastMatch2 (or the oracle
    port) on the sample `job` describes.  Runs in its own process: the reference keeps its state in
    globals.  -> (cells, t_edt, evals, t_match, kind)"""
    crop, pixel, tl, sx, sy, pose0, step, n, budget_s, use_ref = job
    from oracle import pyoracle
    S = crop.shape[0]
    nb = len(sx)
    res3 = np.array([step[0], step[1], step[2]], np.float32)
    if use_ref:
        ref = pyoracle.Reference("accel")
        t0 = time.perf_counter()
        field = ref.edt(crop, fine=True)                      # reference EDT2, <= 400 x 400
        t_edt = time.perf_counter() - t0
        ref.set_map(field, pixel, tl, fine=True)
        ref.set_scan(sx, sy)
        calls, t0 = 0, time.perf_counter()
        while True:
            ref.fastmatch(pose0, res3, fine=True)             # 5 sweeps x 27 candidates
            calls += 1
            t_fm = time.perf_counter() - t0
            if t_fm > max(1.0, budget_s - t_edt) or calls >= 20000:
                break
        return S * S, t_edt, calls * 135 * nb, t_fm, "reference"
    orc = pyoracle.Oracle()
    t0 = time.perf_counter()
    field = orc.edt(crop, variant="scatter")                  # the reference's scatter-form loop nest
    t_edt = time.perf_counter() - t0
    om = orc.make_map(field, pixel, tl)
    nn = (min(n[0], 16), n[1], n[2])
    t0 = time.perf_counter()
    orc.score_lattice(om, sx, sy, pose0, step, nn, want_scores=False)
    t_fm = time.perf_counter() - t0
    return S * S, t_edt, nn[0] * nn[1] * nn[2] * nb, t_fm, "port"


def host_cores() -> int:
    return max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))


def cpu_sample(w, synth, budget_s: float = 12.0, want_kind: str | None = None, cores: int | None = None,
               pool=None) -> dict:
    """Times the reference CPU path on a bounded sample of workload `w`.  The reference is a
    single-threaded program, so "all host threads" means one independent copy per core (separate
    processes; the work is embarrassingly parallel over candidates / grid tiles) and the rates add.
    kind "reference": the reference's own euclidean_distance_transform2 and FastMatch2, compiled
    unmodified (oracle/_ref); kind "port": the oracle restatement."""
    import multiprocessing as mp
    from oracle import pyoracle
    rows, cols = w["occ"].shape
    n = w["n"]
    nb_full = len(w["scan_x"])
    evals_full = n[0] * n[1] * n[2] * nb_full
    cells_full = rows * cols
    use_ref = pyoracle.reference_available() and want_kind != "port"
    S = min(400, rows, cols)
    r0, c0 = (rows - S) // 2, (cols - S) // 2
    crop = np.ascontiguousarray(w["occ"][r0:r0 + S, c0:c0 + S])
    pixel = float(w["pixel"])
    tl = (np.float32(w["top_left"][0] + c0 * pixel), np.float32(w["top_left"][1] + r0 * pixel))
    nb = min(nb_full, 1079)
    sx, sy = np.array(w["scan_x"][:nb]), np.array(w["scan_y"][:nb])
    if cores is None:
        cores = host_cores()
    job = (crop, pixel, tl, sx, sy, np.array(w["pose0"]), np.array(w["step"]), tuple(n), budget_s, use_ref)
    if pool is not None:
        results = pool.map(_cpu_worker, [job] * cores)
    elif cores == 1:
        results = [_cpu_worker(job)]
    else:
        with mp.get_context("spawn").Pool(cores) as own:
            results = own.map(_cpu_worker, [job] * cores)
    kind = results[0][4]
    cells_per_s = sum(c / t for c, t, _, _, _ in results)
    evals_per_s = sum(e / t for _, _, e, t, _ in results)
    t_edt = max(r[1] for r in results)
    t_fm = max(r[3] for r in results)
    calls = results[0][2] // (135 * nb) if use_ref else 0
    what = (f"per core: reference euclidean_distance_transform2 on a {S}x{S} crop ({t_edt:.3f} s) + ~{calls} x reference "
            f"FastMatch2 (135 candidate evals x {nb} beams each, {t_fm:.3f} s)") if use_ref else (
            f"per core: oracle scatter-form EDT on a {S}x{S} crop ({t_edt:.3f} s) + oracle lattice x {nb} beams ({t_fm:.3f} s)")
    # one full step on this CPU at the sampled rates (the EDT term is generous to the
    # reference: its loop nest is O(occupied x cells), so it slows down with area)
    t_step = cells_full / cells_per_s + evals_full / evals_per_s
    return {"value": evals_full / t_step, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": what + f"; {cores} independent copies, rates summed; value = full-step evals / "
                             "(cells/rate_edt + evals/rate_match), extrapolated",
            "match_evals_per_s": evals_per_s, "edt_mcells_per_s": cells_per_s / 1e6,
            "host_cpus": os.cpu_count()}


def run_reference_arm(args, synth):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = synth.make_workload(args.workload)
    K, W = args.steps, args.warmup
    import multiprocessing as mp
    per = max(0.4, min(4.0, 150.0 / max(K + W, 1)))
    vals, last = [], None
    cores = host_cores()
    t_all = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        for i in range(W + K):
            s = cpu_sample(w, synth, budget_s=per, cores=cores, pool=pool)
            if i >= W:
                vals.append(s["value"])
            last = s
            if time.perf_counter() - t_all > 240 and len(vals) >= 3:
                break
    v = statistics.median(vals)
    n = w["n"]
    evals_full = n[0] * n[1] * n[2] * len(w["scan_x"])
    last["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(vals), "warmup": W, "ms_per_step": evals_full / v * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w, args, 1), "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(w, args, world) -> dict:
    rows, cols = w["occ"].shape
    n = w["n"]
    return {"workload": f"{args.workload}: synthetic {rows}x{cols} occupancy grid EDT (max_dist 10) + "
                        f"correlative scan match over {n[0] * world}x{n[1]}x{n[2]} = {n[0] * n[1] * n[2] * world} "
                        f"candidate poses x {len(w['scan_x'])} beams",
            "grid": [rows, cols], "lattice_per_gpu": list(n), "beams": len(w["scan_x"]),
            "pixel_m": float(w["pixel"]), "lattice_step": [float(x) for x in w["step"]],
            "parallelism": f"candidate rows sharded over {world} GPU(s), map replicated",
            "l2": "ring of distinct maps larger than the 126 MB L2, one per step"}


# ----------------------------------------------------------------------------------------
def run_b200_arm(args, synth):
    """The default arm.  Prints the bench line for args.workload; when that is config1 (BASELINE
    configs[1], the single-GPU configuration) the line also carries, under "largest_sweep", the same
    measurements on config3 (configs[3]: 8192^2 grid, 4 M poses x 1080 beams per GPU -- the sweep
    north_star's roofline and scaling targets are stated on), taken in the same run."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = measure_workload(args, synth, args.workload, dist, args.steps, want_cpu=not args.no_cpu)
    if args.workload == "config1" and not args.no_extra:
        big = measure_workload(args, synth, "config3", dist, max(4, min(args.steps, 20)), want_cpu=False)
        if rank == 0:
            keep = ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "serial_ms_per_step", "config",
                    "edt_mcells_per_s", "match_evals_per_s_per_gpu", "roofline", "rooflines", "e2e", "gpu_launches",
                    "result")
            line["largest_sweep"] = {k: big[k] for k in keep if k in big}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def measure_workload(args, synth, workload, dist, K, want_cpu):
    """One workload, measured as the module docstring describes -> the JSON line (rank 0; None elsewhere)."""
    mod = importlib.import_module(PKG)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args = argparse.Namespace(**dict(vars(args), workload=workload))
    W = max(args.warmup, 3)
    line = None

    w = synth.make_workload(args.workload)
    rows, cols = w["occ"].shape
    nth, ntx, nty = w["n"]
    nbeams = len(w["scan_x"])
    n_global = (nth * world, ntx, nty)
    row_b, row_e = rank * nth * ntx, (rank + 1) * nth * ntx
    evals_per_rank = nth * ntx * nty * nbeams
    cells = rows * cols

    ctx = mod.Context(local_rank)
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])

    # ring of maps so that consecutive steps never find their inputs in L2
    set_bytes = cells * 8
    ring = max(2, min(16, -(-2 * L2_BYTES // set_bytes) + 1))
    occs, maps = [], []
    for i in range(ring):
        occ = w["occ"] if i == 0 else synth.grid_rooms(rows, cols, synth.SEED_GRID + i)
        pin = ctx.pinned_empty((rows, cols), np.int32)
        pin[...] = occ
        occs.append(pin)
        m = ctx.new_map(rows, cols)
        m.set_geometry(w["pixel"], w["top_left"]).upload_occupancy(pin)
        maps.append(m)
    scan_x = ctx.pinned_empty((nbeams,), np.float32); scan_x[...] = w["scan_x"]
    scan_y = ctx.pinned_empty((nbeams,), np.float32); scan_y[...] = w["scan_y"]
    ctx.scan_upload(scan_x, scan_y)
    allreduce = world > 1 and not args.no_allreduce
    # K independent steps are kept in flight (pipelined graph below): the throughput policy
    ctx.set_match_mode(mod.MATCH_LATENCY if args.latency_mode else mod.MATCH_THROUGHPUT)

    def step_async(i):
        m = maps[i % ring]
        m.edt(10.0)
        ctx.score_lattice_async(m, w["pose0"], w["step"], n_global, row_b, row_e, allreduce)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    # ---- warm-up (sizes every scratch buffer; also checks the result is sane) ----------
    for i in range(W):
        step_async(i)
    first = ctx.match_fetch()
    assert first.best_index >= 0

    # ---- pass A: eager, CUDA events around every kernel -> per-kernel durations ---------
    KA = min(K, 300)
    barrier()
    for i in range(KA):
        ctx.event_record(3 * i)
        maps[i % ring].edt(10.0)
        ctx.event_record(3 * i + 1)
        ctx.score_lattice_async(maps[i % ring], w["pose0"], w["step"], n_global, row_b, row_e)
        ctx.event_record(3 * i + 2)
    ctx.sync()
    edt_ms = [ctx.event_elapsed_ms(3 * i, 3 * i + 1) for i in range(KA)]
    lat_ms = [ctx.event_elapsed_ms(3 * i + 1, 3 * i + 2) for i in range(KA)]
    edt_ms_avg, lat_ms_avg = sum(edt_ms) / KA, sum(lat_ms) / KA
    # the same with the library's DEFAULT policy (latency mode: the tile shape a caller who runs one
    # match at a time gets) -- reported next to the throughput-mode kernel of the timed region
    ctx.set_match_mode(mod.MATCH_LATENCY)
    for i in range(3):
        ctx.score_lattice_async(maps[i % ring], w["pose0"], w["step"], n_global, row_b, row_e)
    for i in range(KA):
        ctx.event_record(3 * i)
        maps[i % ring].edt(10.0)
        ctx.event_record(3 * i + 1)
        ctx.score_lattice_async(maps[i % ring], w["pose0"], w["step"], n_global, row_b, row_e)
        ctx.event_record(3 * i + 2)
    ctx.sync()
    lat_ms_latency_mode = sum(ctx.event_elapsed_ms(3 * i + 1, 3 * i + 2) for i in range(KA)) / KA
    ctx.set_match_mode(mod.MATCH_LATENCY if args.latency_mode else mod.MATCH_THROUGHPUT)

    # ---- pass B: the timed region.  The step is two microsecond-scale kernels, so steps are
    # captured as CUDA graphs: one graph holding a whole turn (`turn_len` consecutive steps) for the
    # bulk and one holding the K % turn_len remaining steps; K steps are replayed exactly.
    # Consecutive steps are independent (each has its own map), so inside a turn the transforms
    # run on a second stream (a second context on the same GPU): the EDT of step i+1 executes
    # under the match of step i.  The un-pipelined graph is timed too (`serial_ms_per_step`).
    turn = turn_serial = rem = rem_serial = None
    ctx_e = None
    use_graph = not args.no_graph
    # a turn = two passes over the ring when that stays a short burst (config 1: 18 steps): the pipeline
    # drain / fill at the graph boundary, and for N > 1 the collect that lines the ranks up, cost per turn
    turn_len = ring * 2 if ring <= 15 else ring
    # N > 1: inside a turn every match only RECORDS its per-rank best (allreduce = 3) and the collect at
    # the end of the turn sends the whole burst to the peers over NVLink and merges every step's results:
    # no scoring kernel waits for a peer or has peer stores in flight.  Turns too long for the exchange
    # ring post from each kernel's tail and merge the previous step's posts there (allreduce = 2).
    post = (3 if turn_len <= 31 else 2) if allreduce else 0

    def capture_serial(nsteps):
        ctx.graph_begin()
        for k in range(nsteps):
            i = k % ring
            maps[i].edt(10.0)
            ctx.score_lattice_async(maps[i], w["pose0"], w["step"], n_global, row_b, row_e, post)
        if allreduce:
            ctx.exchange_collect_async()
        return ctx.graph_end()

    def capture_pipelined(nsteps):
        ctx.graph_begin()
        ctx.event_record(3000)
        ctx_e.event_wait(ctx, 3000)                   # fork: the transform stream joins the capture
        for k in range(nsteps):
            i = k % ring
            ctx_e._check(ctx_e.L.b200slam_map_edt(ctx_e.h, maps[i].h, 10.0))
            ctx_e.event_record(3100 + k)
            ctx.event_wait(ctx_e, 3100 + k)
            ctx.score_lattice_async(maps[i], w["pose0"], w["step"], n_global, row_b, row_e, post)
        if allreduce:
            ctx.exchange_collect_async()
        ctx_e.event_record(3200)
        ctx.event_wait(ctx_e, 3200)                   # join
        return ctx.graph_end()

    if use_graph:
        try:
            if not args.no_pipeline:
                ctx.set_match_mode(mod.MATCH_LATENCY)          # strictly sequential kernels: the default policy
            turn_serial = capture_serial(turn_len)
            rem_serial = capture_serial(K % turn_len) if K % turn_len else None
            turn, rem = turn_serial, rem_serial
            ctx.set_match_mode(mod.MATCH_LATENCY if args.latency_mode else mod.MATCH_THROUGHPUT)
            if not args.no_pipeline:
                ctx_e = mod.Context(local_rank)
                for i in range(ring):                         # warm the second context's kernels
                    ctx_e._check(ctx_e.L.b200slam_map_edt(ctx_e.h, maps[i].h, 10.0))
                ctx_e.sync()
                turn = capture_pipelined(turn_len)
                rem = capture_pipelined(K % turn_len) if K % turn_len else None
        except mod.B200SlamError as e:
            if rank == 0:
                print(f"[bench] graph capture unavailable ({e}); timing eager launches", file=sys.stderr)
            use_graph = False
            turn = turn_serial = rem = rem_serial = None

    def run_steps(n, turn_graph, rem_graph=None):
        """n steps: whole turns, then (n == K only) the graph holding the K % turn_len remaining steps."""
        if not use_graph:
            for i in range(n):
                step_async(i)
            return
        for _ in range(n // turn_len):
            ctx.graph_launch(turn_graph)
        if n % turn_len:
            assert rem_graph is not None and n == K
            ctx.graph_launch(rem_graph)

    serial_ms = None
    if use_graph and turn is not turn_serial:
        run_steps(-(-max(W, turn_len) // turn_len) * turn_len, turn_serial)
        barrier()
        if world > 1 and allreduce:
            run_steps(turn_len, turn_serial)
        ctx.event_record(4002)
        run_steps(K, turn_serial, rem_serial)
        ctx.event_record(4003)
        barrier()
        serial_ms = ctx.event_elapsed_ms(4002, 4003) / K
    run_steps(-(-max(W, turn_len) // turn_len) * turn_len, turn)
    barrier()
    if world > 1 and use_graph and allreduce:
        # The ranks leave the host-side barrier hundreds of microseconds apart -- a quarter of a
        # 100-step timed region at 25 us per step.  One more untimed turn, which ends in the
        # device-side collect, lines the GPUs up; the start event follows it on the stream.
        run_steps(turn_len, turn)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.event_record(4000)
    run_steps(K, turn, rem)
    ctx.event_record(4001)
    barrier()
    launches = 2 * K + (K // turn_len + (1 if K % turn_len else 0) if allreduce else 0)   # EDT + scan-matching kernel per step (+ one collect per turn when N > 1)
    dev_ms = ctx.event_elapsed_ms(4000, 4001)
    clocks = sampler.stop() if sampler else None
    last = ctx.match_fetch()

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -------------
    KE = max(3, min(K, 50 if cells <= (1 << 23) else 10))
    res = mod.Match()

    def step_e2e(i):
        m = maps[i % ring]
        m.upload_occupancy(occs[i % ring])                     # H2D int32 grid (pinned)
        m.edt(10.0)
        ctx.scan_upload(scan_x, scan_y)                        # H2D scan (pinned)
        return ctx.score_lattice_rows(m, w["pose0"], w["step"], n_global, row_b, row_e, allreduce)  # D2H result

    ctx.set_match_mode(mod.MATCH_LATENCY)                      # one synchronous call at a time
    for i in range(2):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(KE):
        res = step_e2e(i)
    ctx.sync()
    e2e_serial_s = time.perf_counter() - t0
    barrier()
    # The same K steps, double buffered the way a caller would with two contexts: step i+1's grid is
    # already crossing PCIe (its own context = its own stream, pinned source) while step i is
    # transformed, matched and its result read back.  Every step still uploads its grid and scan and
    # reads its result inside the timed region.
    ctx_b = mod.Context(local_rank)
    if world > 1:
        uid = [ctx_b.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx_b.comm_init(world, rank, uid[0])
    pair = (ctx, ctx_b)

    def issue(i, c):
        m = maps[i % ring]
        c._check(c.L.b200slam_map_upload_occupancy(c.h, m.h, occs[i % ring].ctypes.data, occs[i % ring].strides[0] // 4))
        c._check(c.L.b200slam_map_edt(c.h, m.h, 10.0))
        c.scan_upload(scan_x, scan_y)
        c.score_lattice_async(m, w["pose0"], w["step"], n_global, row_b, row_e, 1 if allreduce else 0)

    def run_e2e(n):
        issue(0, pair[0])
        r = None
        for i in range(1, n):
            issue(i, pair[i & 1])
            r = pair[(i - 1) & 1].match_fetch()                # D2H result of step i-1
        return pair[(n - 1) & 1].match_fetch()

    run_e2e(4)
    barrier()
    t0 = time.perf_counter()
    res = run_e2e(KE)
    e2e_s = time.perf_counter() - t0
    barrier()
    h2d = cells * 4 + 2 * nbeams * 4 + (2 * n_global[0] + ntx + nty) * 4
    d2h = 16

    # ---- max over ranks ---------------------------------------------------------------
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_s, edt_ms_avg, lat_ms_avg, serial_ms or 0.0, e2e_serial_s], dtype=torch.float64,
                         device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, edt_ms_avg, lat_ms_avg, smax, e2e_serial_s = [float(x) for x in t.tolist()]
        serial_ms = smax if serial_ms is not None else None

    if rank == 0:
        ms_per_step = dev_ms / K
        total_evals = evals_per_rank * world
        value = total_evals / (ms_per_step * 1e-3)
        peak, peak_src = measured_peak_gbs()
        edt_bytes = 8.0 * cells                                              # 8 B / cell
        lat_bytes = 4.0 * evals_per_rank + 4.0 * nth * ntx * nty + 8.0 * nbeams   # 4 B / eval + 4 B / pose
        roofs = {
            "edt_tma_kernel": {"bound": "hbm", "achieved": edt_bytes / (edt_ms_avg * 1e-3) / 1e9, "peak": peak,
                                 "unit": "GB/s", "ms": edt_ms_avg, "algorithmic_bytes": edt_bytes,
                                 "traffic": ncu_traffic(f"edt_tma_kernel:{args.workload}"),
                                 "mcells_per_s": cells / (edt_ms_avg * 1e-3) / 1e6},
            "lattice_kernel": {"bound": "hbm", "achieved": lat_bytes / (lat_ms_avg * 1e-3) / 1e9, "peak": peak,
                               "unit": "GB/s", "ms": lat_ms_avg, "algorithmic_bytes": lat_bytes,
                               "traffic": ncu_traffic(f"lattice_kernel:{args.workload}"),
                               "evals_per_s": evals_per_rank / (lat_ms_avg * 1e-3),
                               "note": "4 B per pose x beam evaluation (SURVEY.md 8d); the gathers are served by "
                                       "L1/L2 (the field is cache resident), so this HBM-equivalent figure may "
                                       "exceed the DRAM peak"},
        }
        roofs["lattice_kernel"]["policy"] = "latency" if args.latency_mode else "throughput"
        roofs["lattice_kernel_latency_mode"] = dict(
            roofs["lattice_kernel"], ms=lat_ms_latency_mode, policy="latency (library default)",
            achieved=lat_bytes / (lat_ms_latency_mode * 1e-3) / 1e9,
            evals_per_s=evals_per_rank / (lat_ms_latency_mode * 1e-3))
        for r in roofs.values():
            r["frac"] = r["achieved"] / r["peak"]
        dom = max(("edt_tma_kernel", "lattice_kernel"), key=lambda k: roofs[k]["ms"])
        roofline = dict(roofs[dom])
        roofline["kernel"] = dom
        roofline["peak_source"] = peak_src
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(w, args, world), ring_maps=ring,
                           timing=(f"cuda-graph replay ({turn_len} steps per graph"
                                   + (", EDT of step i+1 on a second stream under the match of step i)" if ctx_e else ")"))
                           if use_graph else "eager launches"),
            "serial_ms_per_step": serial_ms,
            "edt_mcells_per_s": cells / (edt_ms_avg * 1e-3) / 1e6,
            "match_evals_per_s_per_gpu": evals_per_rank / (lat_ms_avg * 1e-3),
            "roofline": roofline, "rooflines": roofs,
            "e2e": {"value": total_evals / (e2e_s / KE), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / KE * 1e3, "steps": KE,
                    "how": "two contexts alternate: the next step's H2D runs under this step's kernels and result D2H",
                    "one_context_ms_per_step": e2e_serial_s / KE * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "result": {"best_index": int(last.best_index), "best_score": float(last.best_score),
                       "best_hits": int(last.best_hits), "e2e_best_index": int(res.best_index)},
        }
        if world == 1 and want_cpu:
            line["cpu_baseline"] = cpu_sample(w, synth, budget_s=args.cpu_seconds)

    for g in {id(g): g for g in (turn, turn_serial, rem, rem_serial) if g is not None}.values():
        ctx.graph_destroy(g)
    if ctx_e is not None:
        ctx_e.close()
    for m in maps:
        m.close()
    if dist is not None:
        dist.barrier()                    # nobody tears its peer-mapped buffers down while a peer still spins on them
    ctx_b.close()
    ctx.close()
    return line


def run_particles_arm(args, synth):
    """--workload config2 (BASELINE.json configs[2]): FastSLAM-style 100 000 particles x 720 beams;
    a step = score every particle against the distance field + weights + normalise + systematic
    resample + gather, device resident (single GPU; under torchrun every rank runs an independent
    replica -- the resident particle set does not shard)."""
    mod = importlib.import_module(PKG)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps + (args.steps & 1), max(args.warmup, 3)      # even: two steps per graph (buffers swap)
    w = synth.make_workload("config1")
    rows, cols = w["occ"].shape
    P, nbeams, beta, u0 = 100000, 720, 0.05, 0x80000000
    ctx = mod.Context(local_rank)
    m = ctx.new_map(rows, cols)
    m.set_geometry(w["pixel"], w["top_left"]).upload_occupancy(w["occ"]).edt()
    sx, sy = synth.scan_fixed_count(w["occ"], float(w["pixel"]), w["top_left"], w["true_pose"], nbeams)
    ctx.scan_upload(sx, sy)
    poses = synth.particles_gaussian(P, w["true_pose"])
    ctx.particles_upload(poses)

    def step():
        ctx.particles_score_async(m)
        ctx.particles_resample_async(beta, u0)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    for _ in range(2 * ((W + 1) // 2)):
        step()
    ctx.sync()
    # per-kernel split, eager
    KA = 20
    for i in range(KA):
        ctx.event_record(3 * i); ctx.particles_score_async(m)
        ctx.event_record(3 * i + 1); ctx.particles_resample_async(beta, u0)
        ctx.event_record(3 * i + 2)
    ctx.sync()
    sc_ms = sum(ctx.event_elapsed_ms(3 * i, 3 * i + 1) for i in range(KA)) / KA
    rs_ms = sum(ctx.event_elapsed_ms(3 * i + 1, 3 * i + 2) for i in range(KA)) / KA
    ctx.graph_begin(); step(); step(); g = ctx.graph_end()
    launches0 = ctx.launch_count()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.event_record(4000)
    for _ in range(K // 2):
        ctx.graph_launch(g)
    ctx.event_record(4001)
    barrier()
    dev_ms = ctx.event_elapsed_ms(4000, 4001)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    best = ctx.match_fetch()
    # e2e: particles from host memory every step, weights + ancestors back
    KE = 10
    t0 = time.perf_counter()
    for _ in range(KE):
        res, _, _ = ctx.score_poses(m, poses, want_hits=False)
        ctx.weights_resample(P, beta, u0)
    e2e_s = (time.perf_counter() - t0) / KE
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_s, sc_ms, rs_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, sc_ms, rs_ms = [float(x) for x in t.tolist()]
    if rank == 0:
        evals = P * nbeams
        ms_per_step = dev_ms / K
        peak, peak_src = measured_peak_gbs()
        sc_bytes = 4.0 * evals + 16.0 * P + 8.0 * nbeams            # 4 B / eval + 16 B / pose
        line = {
            "metric": METRIC, "value": evals * world / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config2: FastSLAM-style {P} particles x {nbeams} beams on a {rows}x{cols} map: score + "
                                   "weights + normalise + systematic resample + gather, device resident",
                       "particles": P, "beams": nbeams, "beta": beta, "parallelism": f"{world} independent replica(s)",
                       "l2": "the 16.8 MB distance field is L2 resident by design (gather workload)",
                       "timing": "cuda-graph replay (2 steps per graph)"},
            "roofline": {"bound": "hbm", "kernel": "poses_kernel", "achieved": sc_bytes / (sc_ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": sc_bytes / (sc_ms * 1e-3) / 1e9 / peak, "ms": sc_ms,
                         "algorithmic_bytes": sc_bytes, "traffic": ncu_traffic("poses_kernel:config2"),
                         "peak_source": peak_src, "resample_ms": rs_ms,
                         "note": "4 B per pose x beam evaluation; gathers are served by L1/L2"},
            "e2e": {"value": evals * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 12 * P + 8 * nbeams,
                    "d2h_bytes_per_step": 12 * P + 16, "ms_per_step": e2e_s * 1e3, "steps": KE},
            "gpu_launches": int(launches), "clocks": clocks,
            "result": {"best_index": int(best.best_index), "best_score": float(best.best_score)},
        }
        print(json.dumps(line), flush=True)
    ctx.graph_destroy(g)
    m.close()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_pyramid_arm(args, synth):
    """--workload config4 (BASELINE.json configs[4]): multi-resolution correlative search over a
    3-level EDT pyramid (2048^2 @ 4p, 4096^2 @ 2p, 8192^2 @ p; coarsest first), 10 M coarse poses
    over 8 GPUs = 20 x 250 x 250 = 1.25 M per GPU (weak scaling: the theta range grows with N),
    then two 16 x 32 x 32 refinements, each level seeded by the previous winner
    (Subsystem_1/main.c:901-918 generalised).  A step = the three transforms + the three
    matches (b200slam_pyramid_match: candidate rows of every level sharded over the ranks, the
    per-rank bests exchanged inside the kernels' tails)."""
    mod = importlib.import_module(PKG)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, max(args.warmup, 3)
    w = synth.make_workload("config3")
    occ_f = w["occ"]
    nbeams = len(w["scan_x"])
    ctx = mod.Context(local_rank)
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
    maps, pins = [], []
    for f in (4, 2, 1):                                        # coarsest first
        rows, cols = occ_f.shape[0] // f, occ_f.shape[1] // f
        pin = ctx.pinned_empty((rows, cols), np.int32)
        pin[...] = occ_f.reshape(rows, f, cols, f).max(axis=(1, 3))
        pixel, tl = synth.centred_geometry(rows, cols, 0.1 * f)
        m = ctx.new_map(rows, cols)
        m.set_geometry(pixel, tl).upload_occupancy(pin)
        maps.append(m); pins.append(pin)
    scan_x = ctx.pinned_empty((nbeams,), np.float32); scan_x[...] = w["scan_x"]
    scan_y = ctx.pinned_empty((nbeams,), np.float32); scan_y[...] = w["scan_y"]
    ctx.scan_upload(scan_x, scan_y)
    steps = np.array([[0.2, 0.2, 0.034908], [0.1, 0.1, 0.017454], [0.05, 0.05, 0.008727]], np.float32)
    ns = np.array([[20 * world, 250, 250], [16, 32, 32], [16, 32, 32]], np.int32)
    poses_total = int(sum(int(a) * int(b) * int(c) for a, b, c in ns))
    evals_total = poses_total * nbeams
    cells = sum(int(p.size) for p in pins)

    def step():
        for m in maps:
            m.edt(10.0)
        return ctx.pyramid_match(maps, w["pose0"], steps, ns)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    for _ in range(W):
        res = step()
    # dominant kernel: this rank's share of the coarsest level, an event pair around each launch
    KA = 10
    rb, re = rank * 20 * 250, (rank + 1) * 20 * 250
    barrier()
    for i in range(KA):
        ctx.event_record(2 * i)
        ctx.score_lattice_async(maps[0], w["pose0"], steps[0], tuple(int(x) for x in ns[0]), rb, re, False)
        ctx.event_record(2 * i + 1)
    ctx.sync()
    lat_ms = sum(ctx.event_elapsed_ms(2 * i, 2 * i + 1) for i in range(KA)) / KA
    launches0 = ctx.launch_count()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.event_record(4000)
    for _ in range(K):
        res = step()
    ctx.event_record(4001)
    barrier()
    dev_ms = ctx.event_elapsed_ms(4000, 4001)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    # e2e: the three occupancy grids and the scan from pinned host memory every step
    KE = max(3, min(K, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(KE):
        for m, pin in zip(maps, pins):
            m.upload_occupancy(pin)
        ctx.scan_upload(scan_x, scan_y)
        res_e = step()
    ctx.sync()
    e2e_s = (time.perf_counter() - t0) / KE
    barrier()
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_s, lat_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, lat_ms = [float(x) for x in t.tolist()]
    if rank == 0:
        ms_per_step = dev_ms / K
        peak, peak_src = measured_peak_gbs()
        coarse_evals = 20 * 250 * 250 * nbeams
        lat_bytes = 4.0 * coarse_evals + 4.0 * 20 * 250 * 250 + 8.0 * nbeams
        line = {
            "metric": METRIC, "value": evals_total / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config4: 3-level EDT pyramid (2048^2 @ 0.4 m, 4096^2 @ 0.2 m, 8192^2 @ 0.1 m, max_dist 10) + "
                                   f"coarse-to-fine correlative search, {poses_total} candidate poses x {nbeams} beams "
                                   f"({20 * world}x250x250 coarse, then 16x32x32 twice)",
                       "lattices": [[int(x) for x in r] for r in ns], "beams": nbeams,
                       "lattice_steps": [[float(x) for x in r] for r in steps],
                       "parallelism": f"candidate rows of every level sharded over {world} GPU(s), maps replicated",
                       "l2": "the three maps (704 MB of occupancy + field) exceed the 126 MB L2",
                       "timing": "eager launches, one host round trip per level (the next level is seeded by the winner)"},
            "edt_mcells_per_step": cells / 1e6,
            "roofline": {"bound": "hbm", "kernel": "lattice_kernel", "achieved": lat_bytes / (lat_ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": lat_bytes / (lat_ms * 1e-3) / 1e9 / peak, "ms": lat_ms,
                         "algorithmic_bytes": lat_bytes, "traffic": ncu_traffic("lattice_kernel:config4"),
                         "evals_per_s": coarse_evals / (lat_ms * 1e-3), "peak_source": peak_src,
                         "note": "coarsest level, this rank's 20 x 250 x 250 share; 4 B per pose x beam evaluation, gathers "
                                 "served by L1/L2"},
            "e2e": {"value": evals_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": cells * 4 + 8 * nbeams,
                    "d2h_bytes_per_step": 3 * 16, "ms_per_step": e2e_s * 1e3, "steps": KE},
            "gpu_launches": int(launches), "clocks": clocks,
            "result": {"best_index": [int(r.best_index) for r in res], "best_score": [float(r.best_score) for r in res],
                       "e2e_best_index": [int(r.best_index) for r in res_e]},
        }
        print(json.dumps(line), flush=True)
    for m in maps:
        m.close()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config1", choices=["config1", "config2", "config3", "config4", "tiny"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="keep every step's EDT and match strictly back to back (no second stream)")
    ap.add_argument("--latency-mode", action="store_true",
                    help="keep the library's default tile-shape policy (one match at a time) in the timed region too")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="config1 only: skip the extra config3 measurement")
    ap.add_argument("--no-allreduce", action="store_true",
                    help="diagnostic: N > 1 without the per-step exchange of bests (ranks run independently)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    synth = importlib.import_module(PKG + ".synth")
    if args.impl == "reference":
        if args.workload == "config2":
            args.workload = "config1"       # the reference has no particle filter: its matcher on the same map
        if args.workload == "config4":
            args.workload = "config3"       # ... and no pyramid: its matcher on the finest map
        run_reference_arm(args, synth)
    elif args.workload == "config2":
        run_particles_arm(args, synth)
    elif args.workload == "config4":
        run_pyramid_arm(args, synth)
    else:
        run_b200_arm(args, synth)


if __name__ == "__main__":
    main()
