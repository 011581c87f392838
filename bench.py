#!/usr/bin/env python
"""bench.py -- pose x beam evals/s and EDT Mcells/s of the b200slam hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config3|config1|config2|config4|tiny]
    python bench.py --impl reference ...        # the reference's own CPU code, same metric

A "step" is one pass of the hot path over one batch of synthetic input: the clamped EDT of the
occupancy grid followed by the correlative scan match of every lattice candidate against the fresh
distance field, ending in the arg-min.  Default workload: BASELINE.json configs[3], the largest
pose x beam sweep that fits one GPU (8192x8192 grid, 256 x 128 x 128 = 4 194 304 candidate poses x
1080 beams per GPU) -- the sweep north_star states its roofline and scaling targets on.  The same line
carries, measured in the same run, a block per other configuration: "config1" (configs[1]: 2048^2 grid,
64 x 32 x 32 poses x 360 beams), "config2" (configs[2]: 100 000 particles x 720 beams per GPU, weights
+ normalise + systematic resample, sharded over the ranks), "config4" (configs[4]: 3-level EDT pyramid,
10 M poses over 8 GPUs) and, for N > 1, "edt_row_sharded" (configs[3]'s row-sharded transform against
replicated compute).

Own arm, per rank (one process per GPU; torch.distributed only for barrier / max-over-ranks):
  * `value`  : whole-job pose x beam evals / s with inputs resident in HBM.  K steps are replayed back to
               back as CUDA graphs (the step is a handful of kernels); the K-step region is repeated R >= 15
               times, each repetition behind a device-side barrier over the ranks and bracketed by CUDA
               events on the library's stream; per repetition the max over ranks is taken and the MEDIAN
               repetition is reported (`ms_per_step` = that / K; `reps` has the spread).  The whole
               measurement is bracketed by barrier + sync.  Steps cycle through a ring of distinct maps
               larger than L2.  `serial_ms_per_step`: the same K steps strictly one kernel after the other.
  * `result.verified`: after the timed region the last step's winner (index, score bits, hit counts) is
               compared with (a) the strictly sequential graph, (b) the end-to-end host-buffer calls and (c)
               a plain synchronous call that scores the WHOLE lattice of all N ranks on rank 0 alone --
               at N > 1 that proves the exchanged global winner on every rank.
  * `roofline` / `rooflines`: per kernel, from back-to-back launches of that kernel alone inside a CUDA
               graph (its in-pipeline duration) -- HBM roofline for the transform, issue-slot roofline for
               the matcher (its gathers are served by L1/L2; the HBM-equivalent figure stays as a note)
  * `e2e`    : the same step through the host-buffer C ABI calls in the reference's own order (main.c:884-918:
               OccupationalGrid -> euclidean_distance_transform -> FastMatch): H2D of the map's POINTS (what
               OccupationalGrid takes; rasterised on the device) and the scan from pinned memory, D2H of the
               match result, all inside the timed region.  `e2e_int32_grid`: the same with the already
               rasterised int32 grid as the host buffer (4 B per cell: PCIe bound)
  * `cpu_baseline` (rank 0, N == 1): the reference's own EDT2 + FastMatch2 (oracle/_ref) or the oracle
               port on a bounded sample of the same workload
N > 1: weak scaling -- every rank keeps the per-GPU workload (map replicated, its own block of theta rows
of an N-times larger lattice); the per-rank bests travel through NVLink peer memory written by the
scoring kernels / one collect kernel per graph (NCCL all-gather where CUDA IPC is unavailable).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "hardware-acceleration-of-lidar-slam_b200"

METRIC = "pose_x_beam_evals_per_s"
UNIT = "evals/s"
L2_BYTES = 126 * 1024 * 1024
# Issue-slot roofline of the scan matcher: every SM sub-partition issues at most one warp
# instruction per cycle -> peak = 4 x SMs x SM clock.  Algorithmic instructions per pose x beam evaluation of
# the bit-exact formulation (one thread owns a candidate's sequential sum): see DESIGN.md for the derivation.
ISSUE_FLOOR_LANE_INSTR_PER_EVAL = {"lattice_rr2": 34.0 / 16.0, "lattice": 3.0, "poses": 14.0}


def lattice_floor(n, step, pixel):
    """Floor of the kernel variant the launcher picks for this lattice: row reuse (9 gathers + 9 address ops + 16
    adds per 16 candidates) when the ty step is half a pixel and the lattice is large enough for 64 x 64 tiles,
    else one gather + one offset + one address op + one add... per candidate (3 with the loop's shared work)."""
    rr = abs(float(step[1]) / float(pixel) * 2.0 - 1.0) < 1e-3 and n[1] >= 64 and n[2] >= 64
    return ("lattice_rr2" if rr else "lattice"), ISSUE_FLOOR_LANE_INSTR_PER_EVAL["lattice_rr2" if rr else "lattice"]


# ----------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_counter(key: str):
    """Per-launch ncu counter (DRAM bytes, warp instructions) from the committed summary, or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


class ClockSampler:
    """SM clock / throttle reasons / power sampled in-process through NVML every ~2 ms by a thread, from
    before the ranks are lined up until after the timed region (mark() brackets the region itself)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, cuda_index: int):
        self.samples = []            # (t, sm_mhz, reasons mask, power W)
        self.marks = []
        self.err = None
        self._stop = threading.Event()
        self.smax = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = self._handle(cuda_index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:       # noqa: BLE001
            self.err = f"NVML unavailable: {e}"
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _handle(self, cuda_index):
        nv = self.nv
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            tok = [v.strip() for v in vis.split(",") if v.strip()][cuda_index]
            if tok.isdigit():
                return nv.nvmlDeviceGetHandleByIndex(int(tok))
            return nv.nvmlDeviceGetHandleByUUID(tok.encode() if hasattr(tok, "encode") else tok)
        return nv.nvmlDeviceGetHandleByIndex(cuda_index)

    def _run(self):
        if self.nv is None:
            self._run_smi()
            return
        nv, i = self.nv, 0
        power = 0.0
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:    # noqa: BLE001
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                if i % 8 == 0:
                    power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), sm, mask, power))
            except Exception as e:   # noqa: BLE001
                self.err = str(e)
                break
            i += 1
            time.sleep(0.002)

    def _run_smi(self):
        """Fallback without pynvml: nvidia-smi polled at 20 ms."""
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception as e:       # noqa: BLE001
            self.err = f"{self.err}; nvidia-smi unavailable: {e}"
            return
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self._stop.is_set():
            ln = p.stdout.readline()
            if not ln:
                break
            parts = [x.strip() for x in ln.split(",")]
            try:
                mask = sum(b for b, v in zip(bits, parts[3:7]) if v.lower().startswith("active"))
                self.samples.append((time.perf_counter(), float(parts[0]), mask, float(parts[2])))
                self.smax = float(parts[1])
            except (ValueError, IndexError):
                continue
        p.terminate()

    def mark(self):
        self.marks.append(time.perf_counter())

    def stop(self) -> dict:
        self._stop.set()
        self.t.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": [self.err or "no samples"]}
        inside = self.samples
        if len(self.marks) >= 2:
            sel = [s for s in self.samples if self.marks[0] <= s[0] <= self.marks[-1]]
            if len(sel) >= 3:
                inside = sel
        mask = 0
        for s in inside:
            mask |= s[2]
        # "under load": the busy samples are the upper part of the distribution (the thread also sees the
        # idle gaps between host calls)
        sm = sorted(s[1] for s in inside)
        load = sm[len(sm) // 2:]
        return {"sm_mhz": statistics.median(load), "sm_mhz_min": sm[0], "sm_max_mhz": self.smax,
                "samples": len(self.samples), "samples_in_timed_region": len(inside) if inside is not self.samples else 0,
                "power_w_max": max(s[3] for s in self.samples),
                "reasons": sorted(nm for b, nm in self.REASONS.items() if mask & b),
                "how": "NVML in-process, ~2 ms period, started before the ranks are lined up"}


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
            "config": workload_config(w, args.workload, 1), "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(w, name, world) -> dict:
    rows, cols = w["occ"].shape
    n = w["n"]
    return {"workload": f"{name}: synthetic {rows}x{cols} occupancy grid EDT (max_dist 10) + "
                        f"correlative scan match over {n[0] * world}x{n[1]}x{n[2]} = {n[0] * n[1] * n[2] * world} "
                        f"candidate poses x {len(w['scan_x'])} beams",
            "grid": [rows, cols], "lattice_per_gpu": list(n), "beams": len(w["scan_x"]),
            "pixel_m": float(w["pixel"]), "lattice_step": [float(x) for x in w["step"]],
            "parallelism": f"candidate rows sharded over {world} GPU(s), map replicated",
            "l2": "ring of distinct maps larger than the 126 MB L2, one per step"}


# ----------------------------------------------------------------------------------------
class Job:
    """Rank / communicator plumbing shared by the blocks of one bench run."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist, self.torch = dist, torch

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def comm_init(self, ctx):
        if self.world > 1:
            uid = [ctx.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(uid, src=0)
            ctx.comm_init(self.world, self.rank, uid[0])

    def max_over_ranks(self, values):
        """Element-wise max over the ranks of a list of floats."""
        if self.dist is None:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def all_true(self, flag: bool) -> bool:
        return self.max_over_ranks([0.0 if flag else 1.0])[0] == 0.0

    def bcast(self, obj):
        if self.dist is None:
            return obj
        box = [obj if self.rank == 0 else None]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_reps(job: Job, ctx, run_once, est_ms: float, slot0: int = 1000, min_reps: int = 15, max_reps: int = 400,
               target_ms: float = 150.0, sampler: ClockSampler | None = None):
    """R repetitions of `run_once` (the K-step region), each behind a device-side barrier over the ranks and
    bracketed by CUDA events on the library's stream; the whole thing bracketed by sync + host barrier.
    -> (median over repetitions of the max over ranks, all per-repetition maxima) in ms."""
    reps = int(max(min_reps, min(max_reps, -(-target_ms // max(est_ms, 1e-3)))))
    reps = int(job.max_over_ranks([float(reps)])[0])
    ctx.sync()
    job.barrier()
    if sampler:
        sampler.mark()
    for r in range(reps):
        ctx.comm_barrier_async()                 # no-op on one GPU
        ctx.event_record(slot0 + 2 * r)
        run_once()
        ctx.event_record(slot0 + 2 * r + 1)
    ctx.sync()
    if sampler:
        sampler.mark()
    job.barrier()
    ms = [ctx.event_elapsed_ms(slot0 + 2 * r, slot0 + 2 * r + 1) for r in range(reps)]
    ms = job.max_over_ranks(ms)
    return statistics.median(ms), ms


def spread(ms, K):
    s = sorted(ms)
    return {"count": len(s), "median_ms_per_step": statistics.median(s) / K, "min_ms_per_step": s[0] / K,
            "p90_ms_per_step": s[min(len(s) - 1, int(0.9 * len(s)))] / K, "max_ms_per_step": s[-1] / K}


def match_tuple(r):
    return (int(r.best_index), np.float32(r.best_score).tobytes().hex(), int(r.best_hits), int(r.last_hits))


def measure_workload(args, synth, workload, job: Job, K, want_cpu, sample_clocks=True):
    """One EDT + lattice workload, measured as the module docstring describes -> the JSON line (rank 0; None elsewhere)."""
    mod = importlib.import_module(PKG)
    rank, world, local_rank = job.rank, job.world, job.local_rank
    W = max(args.warmup, 3)
    line = None

    w = synth.make_workload(workload)
    rows, cols = w["occ"].shape
    nth, ntx, nty = w["n"]
    nbeams = len(w["scan_x"])
    n_global = (nth * world, ntx, nty)
    row_b, row_e = rank * nth * ntx, (rank + 1) * nth * ntx
    evals_per_rank = nth * ntx * nty * nbeams
    cells = rows * cols

    ctx = mod.Context(local_rank)
    job.comm_init(ctx)

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
    policy = mod.MATCH_LATENCY if args.latency_mode else mod.MATCH_THROUGHPUT
    # K independent steps are kept in flight (pipelined graph below): the throughput policy
    ctx.set_match_mode(policy)

    def step_async(i):
        m = maps[i % ring]
        m.edt(10.0)
        ctx.score_lattice_async(m, w["pose0"], w["step"], n_global, row_b, row_e, allreduce)

    def barrier():
        ctx.sync()
        job.barrier()

    # ---- warm-up (sizes every scratch buffer; also checks the result is sane) ----------
    for i in range(W):
        step_async(i)
    first = ctx.match_fetch()
    assert first.best_index >= 0

    # ---- expected winners: rank 0 ALONE scores the whole lattice of all N ranks on every ring map with plain
    # synchronous calls (default policy, no exchange, nothing pipelined) -- what every other launch sequence
    # below has to reproduce on every rank
    barrier()
    expected = None
    if rank == 0:
        ctx.set_match_mode(mod.MATCH_LATENCY)
        expected = []
        for i in range(ring):
            maps[i].edt(10.0)
            expected.append(match_tuple(ctx.score_lattice_rows(maps[i], w["pose0"], w["step"], n_global, 0,
                                                               n_global[0] * ntx, False)))
        ctx.set_match_mode(policy)
    expected = job.bcast(expected)
    barrier()

    # ---- the timed region.  The step is two kernels, so steps are captured as CUDA graphs: one graph holding
    # a whole turn (`turn_len` consecutive steps) for the bulk and one holding the K % turn_len remaining
    # steps; K steps are replayed exactly.  Consecutive steps are independent (each has its own map), so
    # inside a turn the transforms run on a second stream (a second context on the same GPU): the EDT of
    # step i+1 executes under the match of step i.  The un-pipelined graph is timed too.
    turn = turn_serial = rem = rem_serial = None
    ctx_e = None
    use_graph = not args.no_graph
    # a turn = two passes over the ring when that stays a short burst (config 1: 18 steps): the pipeline
    # drain / fill at the graph boundary, and for N > 1 the collect that merges the ranks, cost per turn
    turn_len = ring * 2 if ring <= 15 else ring
    # N > 1: inside a turn every match only RECORDS its per-rank best (allreduce = 3) and the collect at
    # the end of the turn sends the whole burst to the peers over NVLink and merges every step's results:
    # no scoring kernel waits for a peer or has peer stores in flight.  Turns too long for the exchange
    # ring post from each kernel's tail and merge the previous step's posts there (allreduce = 2).
    post = (3 if turn_len <= 31 else 2) if allreduce else 0
    if allreduce and os.environ.get("B200SLAM_BENCH_POST"):               # diagnostics: force the exchange mode
        post = int(os.environ["B200SLAM_BENCH_POST"])

    def capture(nsteps, pipelined, edt=True, match=True):
        ctx.graph_begin()
        if pipelined:
            ctx.event_record(3000)
            ctx_e.event_wait(ctx, 3000)                   # fork: the transform stream joins the capture
        # N > 1, pipelined turns: the collect that sends and merges a burst runs at the START of the next turn, beside
        # that turn's first transform, so a turn still ends with a scoring kernel and the next turn's transform can
        # start under its tail (a turn that ends with the one-CTA collect exposes a transform and a launch gap:
        # 0.619 instead of 0.584 ms per step on config 3).  run_steps() merges the last burst behind the last turn.
        collect_first = pipelined and allreduce and match and post == 3 and not os.environ.get("B200SLAM_BENCH_NO_COLLECT")
        if collect_first:
            ctx.exchange_collect_async()
        for k in range(nsteps):
            i = k % ring
            if edt and pipelined:
                ctx_e._check(ctx_e.L.b200slam_map_edt(ctx_e.h, maps[i].h, 10.0))
                ctx_e.event_record(3100 + k)
                ctx.event_wait(ctx_e, 3100 + k)
            elif edt:
                maps[i].edt(10.0)
            if match:
                ctx.score_lattice_async(maps[i], w["pose0"], w["step"], n_global, row_b, row_e, post)
        if allreduce and match and not collect_first and not os.environ.get("B200SLAM_BENCH_NO_COLLECT"):   # (diagnostics: timing without the merge)
            ctx.exchange_collect_async()
        if pipelined:
            ctx_e.event_record(3200)
            ctx.event_wait(ctx_e, 3200)                   # join
        return ctx.graph_end()

    g_edt_only = g_match_only = None
    if use_graph:
        try:
            if not args.no_pipeline:
                ctx.set_match_mode(mod.MATCH_LATENCY)          # strictly sequential kernels: the default policy
            turn_serial = capture(turn_len, False)
            rem_serial = capture(K % turn_len, False) if K % turn_len else None
            turn, rem = turn_serial, rem_serial
            ctx.set_match_mode(policy)
            if not args.no_pipeline:
                ctx_e = mod.Context(local_rank)
                ctx_e.set_match_mode(policy)                  # the transform reads the same scheduling hint
                for i in range(ring):                         # warm the second context's kernels
                    ctx_e._check(ctx_e.L.b200slam_map_edt(ctx_e.h, maps[i].h, 10.0))
                ctx_e.sync()
                turn = capture(turn_len, True)
                rem = capture(K % turn_len, True) if K % turn_len else None
            # each kernel alone, back to back inside a graph: its in-pipeline duration (matches chain by PDL)
            g_edt_only = capture(turn_len, False, edt=True, match=False)
            g_match_only = capture(turn_len, False, edt=False, match=True)
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
        if allreduce and post == 3 and turn_graph is turn and turn is not turn_serial:
            ctx.exchange_collect_async()                  # the last pipelined turn's burst (inside the timed region)

    last_map = ((K % turn_len or turn_len) - 1) % ring if use_graph else (K - 1) % ring
    checks = {}

    def est(fn):
        """One untimed + one timed run of fn -> ms (sizes the number of repetitions)."""
        fn(); ctx.sync()
        ctx.event_record(4090); fn(); ctx.event_record(4091); ctx.sync()
        return ctx.event_elapsed_ms(4090, 4091)

    sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None
    serial_ms = serial_reps = None
    if use_graph and turn is not turn_serial:
        e = est(lambda: run_steps(K, turn_serial, rem_serial))
        med, allms = timed_reps(job, ctx, lambda: run_steps(K, turn_serial, rem_serial), e, target_ms=60.0)
        serial_ms, serial_reps = med / K, spread(allms, K)
        checks["serial_graph"] = match_tuple(ctx.match_fetch())
    # every graph of the timed region runs at least once before it (the K % turn_len graph included)
    for _ in range(max(2, -(-W // max(K, 1)))):
        run_steps(K, turn, rem)
    e = est(lambda: run_steps(K, turn, rem))
    med_ms, rep_ms = timed_reps(job, ctx, lambda: run_steps(K, turn, rem), e, sampler=sampler)
    launches = 2 * K + ((K // turn_len + (1 if K % turn_len else 0)) if allreduce else 0)   # EDT + matcher per step (+ one collect per graph when N > 1)
    if allreduce and post == 3 and turn is not turn_serial:
        launches += 1                                          # the last burst's collect behind the last turn
    if not use_graph and allreduce:
        launches = 3 * K
    checks["timed_region"] = match_tuple(ctx.match_fetch())
    clocks = sampler.stop() if sampler else None

    # ---- per-kernel durations: each kernel alone, back to back (graph), plus eager with an event pair per launch
    edt_ms_b2b = lat_ms_b2b = None
    if use_graph and g_edt_only is not None:
        e = est(lambda: ctx.graph_launch(g_edt_only))
        edt_ms_b2b = timed_reps(job, ctx, lambda: ctx.graph_launch(g_edt_only), e, target_ms=40.0)[0] / turn_len
        e = est(lambda: ctx.graph_launch(g_match_only))
        lat_ms_b2b = timed_reps(job, ctx, lambda: ctx.graph_launch(g_match_only), e, target_ms=40.0)[0] / turn_len
    KA = min(K, 300)
    barrier()
    for i in range(KA):
        ctx.event_record(3 * i)
        maps[i % ring].edt(10.0)
        ctx.event_record(3 * i + 1)
        ctx.score_lattice_async(maps[i % ring], w["pose0"], w["step"], n_global, row_b, row_e)
        ctx.event_record(3 * i + 2)
    ctx.sync()
    edt_ms_avg = sum(ctx.event_elapsed_ms(3 * i, 3 * i + 1) for i in range(KA)) / KA
    lat_ms_avg = sum(ctx.event_elapsed_ms(3 * i + 1, 3 * i + 2) for i in range(KA)) / KA
    # the same with the library's DEFAULT policy (latency mode: the tile shape a caller who runs one
    # match at a time gets)
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

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -------------
    KE = max(3, min(K, 50 if cells <= (1 << 23) else 10))
    res = mod.Match()

    def step_e2e(i):
        m = maps[i % ring]
        m.upload_occupancy(occs[i % ring])                     # H2D int32 grid (pinned)
        m.edt(10.0)
        ctx.scan_upload(scan_x, scan_y)                        # H2D scan (pinned)
        return ctx.score_lattice_rows(m, w["pose0"], w["step"], n_global, row_b, row_e, allreduce)  # D2H result

    for i in range(2):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(KE):
        res = step_e2e(i)
    ctx.sync()
    e2e_serial_s = time.perf_counter() - t0
    checks["e2e_one_context"] = match_tuple(res)
    barrier()
    # The same steps, double buffered the way a caller would with two contexts: step i+1's grid is
    # already crossing PCIe (its own context = its own stream, pinned source) while step i is
    # transformed, matched and its result read back.  Every step still uploads its grid and scan and
    # reads its result inside the timed region.
    ctx_b = mod.Context(local_rank)
    job.comm_init(ctx_b)
    pair = (ctx, ctx_b)

    def issue(i, c):
        m = maps[i % ring]
        c._check(c.L.b200slam_map_upload_occupancy(c.h, m.h, occs[i % ring].ctypes.data, occs[i % ring].strides[0] // 4))
        c._check(c.L.b200slam_map_edt(c.h, m.h, 10.0))
        c.scan_upload(scan_x, scan_y)
        c.score_lattice_async(m, w["pose0"], w["step"], n_global, row_b, row_e, 1 if allreduce else 0)

    def run_e2e(n):
        issue(0, pair[0])
        for i in range(1, n):
            issue(i, pair[i & 1])
            pair[(i - 1) & 1].match_fetch()                    # D2H result of step i-1
        return pair[(n - 1) & 1].match_fetch()

    run_e2e(4)
    barrier()
    t0 = time.perf_counter()
    res = run_e2e(KE)
    e2e_s = time.perf_counter() - t0
    checks["e2e"] = match_tuple(res)
    barrier()
    h2d = cells * 4 + 2 * nbeams * 4 + (2 * n_global[0] + ntx + nty) * 4
    d2h = 16

    # ---- e2e, the reference's own sequence (main.c:884-918): OccupationalGrid from the map's POINTS ->
    # euclidean_distance_transform -> FastMatch.  The host buffers are the map points (x, y: what the reference's
    # OccupationalGrid takes) and the scan; 8 bytes per point cross PCIe instead of 4 bytes per cell, the grid is
    # rasterised on the device.  Two cells at the 3-pixel margin pin the bounding box so that the rasterised
    # grid is rows x cols again.
    ptx, pty, occ_pts = [], [], []
    for i in range(ring):
        o = np.array(occs[i], copy=True)
        o[3, 3] = 1
        o[rows - 4, cols - 4] = 1
        r_, c_ = np.nonzero(o)
        px = ctx.pinned_empty((len(r_),), np.float32)
        py = ctx.pinned_empty((len(r_),), np.float32)
        px[...] = (float(w["top_left"][0]) + c_ * float(w["pixel"])).astype(np.float32)
        py[...] = (float(w["top_left"][1]) + r_ * float(w["pixel"])).astype(np.float32)
        ptx.append(px); pty.append(py)
        occ_pts.append(o if i == 0 else None)
    npts = max(len(a) for a in ptx)
    barrier()
    expected_pts = None
    raster_ok = True
    if rank == 0:
        ctx.set_match_mode(mod.MATCH_LATENCY)
        expected_pts = []
        for i in range(ring):
            got = maps[i].rasterise(np.array(ptx[i]), np.array(pty[i]), float(w["pixel"]))[:2]
            raster_ok = raster_ok and got == (rows, cols)
            if i == 0:
                raster_ok = raster_ok and bool(np.array_equal(maps[i].download_occupancy(), occ_pts[0]))
            maps[i].edt(10.0)
            expected_pts.append(match_tuple(ctx.score_lattice_rows(maps[i], w["pose0"], w["step"], n_global, 0,
                                                                   n_global[0] * ntx, False)))
        ctx.set_match_mode(policy)
    expected_pts = job.bcast(expected_pts)
    barrier()

    # DEPTH steps in flight, one context each (the contexts take turns): step i is issued once step i - DEPTH has been
    # read back, so the steps in flight need DEPTH different maps (a map object more, fed the first map's points,
    # when the ring is shorter).
    DEPTH = max(2, int(os.environ.get("B200SLAM_BENCH_E2E_DEPTH", "3")))
    pipe = list(pair)
    while len(pipe) < DEPTH:
        c_new = mod.Context(local_rank)
        job.comm_init(c_new)
        pipe.append(c_new)
    emaps, esrc = list(maps), list(range(ring))
    while len(emaps) < DEPTH:
        m_new = ctx.new_map(rows, cols)
        m_new.set_geometry(w["pixel"], w["top_left"])
        emaps.append(m_new); esrc.append(0)
    ER = len(emaps)

    def issue_pts(i, c):
        m = emaps[i % ER]
        m.rasterise_async(ptx[esrc[i % ER]], pty[esrc[i % ER]], float(w["pixel"]), ctx=c)   # H2D points (pinned) + rasterise
        c._check(c.L.b200slam_map_edt(c.h, m.h, 10.0))
        c.scan_upload(scan_x, scan_y)                                                  # H2D scan (pinned)
        c.score_lattice_async(m, w["pose0"], w["step"], n_global, row_b, row_e, 1 if allreduce else 0)

    def run_e2e_pts(n):
        res_ = None
        for i in range(n):
            if i >= DEPTH:
                res_ = pipe[i % DEPTH].match_fetch()           # D2H result of step i - DEPTH: its context and map are free again
            issue_pts(i, pipe[i % DEPTH])
        for i in range(max(n - DEPTH, 0), n):
            res_ = pipe[i % DEPTH].match_fetch()
        return res_

    KP = max(3, min(K, 50))
    run_e2e_pts(4)
    barrier()
    t0 = time.perf_counter()
    res = run_e2e_pts(KP)
    e2e_pts_s = time.perf_counter() - t0
    checks["e2e_points"] = match_tuple(res)
    barrier()
    t0 = time.perf_counter()
    for i in range(KP):
        issue_pts(i, ctx)
        res = ctx.match_fetch()
    e2e_pts_serial_s = time.perf_counter() - t0
    checks["e2e_points_one_context"] = match_tuple(res)
    barrier()
    h2d_pts = 8 * npts + 2 * nbeams * 4 + (2 * n_global[0] + ntx + nty) * 4

    # ---- verification: every launch sequence, on every rank, against rank 0's stand-alone winners ----------
    want = {"serial_graph": expected[last_map], "timed_region": expected[last_map],
            "e2e_one_context": expected[(KE - 1) % ring], "e2e": expected[(KE - 1) % ring],
            "e2e_points": expected_pts[esrc[(KP - 1) % ER]], "e2e_points_one_context": expected_pts[esrc[(KP - 1) % ER]]}
    mism = {k: (v, want[k]) for k, v in checks.items() if tuple(v) != tuple(want[k])}
    if args.no_allreduce and world > 1:
        mism = {}                                              # diagnostic mode: ranks hold shard-local winners
    verified = job.all_true(not mism and raster_ok)
    if mism:
        print(f"[bench] rank {rank}: result mismatch {mism}", file=sys.stderr, flush=True)

    # ---- max over ranks ---------------------------------------------------------------
    e2e_s, edt_ms_avg, lat_ms_avg, e2e_serial_s, lat_ms_latency_mode, e2e_pts_s, e2e_pts_serial_s = job.max_over_ranks(
        [e2e_s, edt_ms_avg, lat_ms_avg, e2e_serial_s, lat_ms_latency_mode, e2e_pts_s, e2e_pts_serial_s])

    if rank == 0:
        ms_per_step = med_ms / K
        total_evals = evals_per_rank * world
        value = total_evals / (ms_per_step * 1e-3)
        peak, sm_mhz, peak_src = measured_peaks()
        info = ctx.device_info()
        edt_ms = edt_ms_b2b if edt_ms_b2b else edt_ms_avg
        lat_ms = lat_ms_b2b if lat_ms_b2b else lat_ms_avg
        how = ("that kernel alone, back to back inside a CUDA graph (its duration inside the pipelined step), "
               "CUDA events on its stream, median of >= 15 repetitions") if edt_ms_b2b else "eager, event pair per launch"
        edt_bytes = 8.0 * cells                                              # 8 B / cell
        lat_bytes = 4.0 * evals_per_rank + 4.0 * nth * ntx * nty + 8.0 * nbeams   # 4 B / eval + 4 B / pose
        issue_peak = 4.0 * info["sm_count"] * sm_mhz * 1e6 / 1e9                 # G warp-instructions / s
        floor_kind, floor = lattice_floor(w["n"], w["step"], w["pixel"])
        alg_winstr = floor * evals_per_rank / 32.0
        meas_winstr = ncu_counter(f"lattice_kernel:{workload}:warp_instructions")
        meas_l1wf = ncu_counter(f"lattice_kernel:{workload}:l1_wavefronts")
        l1_peak = info["sm_count"] * sm_mhz * 1e6 / 1e9                          # G wavefronts / s: one per cycle per SM
        roofs = {
            "edt_tma_kernel": {"bound": "hbm", "achieved": edt_bytes / (edt_ms * 1e-3) / 1e9, "peak": peak,
                               "unit": "GB/s", "ms": edt_ms, "ms_eager_alone": edt_ms_avg, "algorithmic_bytes": edt_bytes,
                               "traffic": ncu_counter(f"edt_tma_kernel:{workload}"),
                               "occupancy_bytes_per_cell_streamed": 1 if cells >= (1 << 20) else 4,
                               "note": "algorithmic bytes are SURVEY.md 8d's 8 B / cell (int32 occupancy read + f32 written).  Maps of "
                                       ">= 2^20 cells keep a byte shadow of the occupancy (packed at upload / written by the "
                                       "rasteriser) and the kernel streams that: 5 B / cell actually move (`traffic`), so `frac` "
                                       "can approach 8 / 5 of what an int32-reading kernel could reach",
                               "mcells_per_s": cells / (edt_ms * 1e-3) / 1e6, "how": how},
            "lattice_kernel": {"bound": "issue", "achieved": alg_winstr / (lat_ms * 1e-3) / 1e9, "peak": issue_peak,
                               "unit": "Gwarp-inst/s", "ms": lat_ms, "ms_eager_alone": lat_ms_avg,
                               "algorithmic_warp_instructions": alg_winstr,
                               "floor_lane_instr_per_eval": floor, "floor_kind": floor_kind,
                               "measured_warp_instructions": meas_winstr,
                               "measured_lane_instr_per_eval": (meas_winstr * 32.0 / evals_per_rank) if meas_winstr else None,
                               "issue_utilisation": (meas_winstr / (lat_ms * 1e-3) / 1e9 / issue_peak) if meas_winstr else None,
                               "l1_wavefronts": {"measured_per_launch": meas_l1wf, "peak_gwf_per_s": l1_peak,
                                                 "utilisation": (meas_l1wf / (lat_ms * 1e-3) / 1e9 / l1_peak) if meas_l1wf else None,
                                                 "note": "second ceiling of the gathers: l1tex__data_pipe_lsu_wavefronts (global + "
                                                         "shared) against one wavefront per cycle per SM"},
                               "traffic": ncu_counter(f"lattice_kernel:{workload}"),
                               "evals_per_s": evals_per_rank / (lat_ms * 1e-3),
                               "peak_how": f"4 warp schedulers x {info['sm_count']} SMs x {sm_mhz:.0f} MHz",
                               "hbm_equivalent": {"algorithmic_bytes": lat_bytes, "gbs": lat_bytes / (lat_ms * 1e-3) / 1e9,
                                                  "note": "4 B per pose x beam evaluation (SURVEY.md 8d); the gathers are "
                                                          "served by L1/L2 (the field is cache resident: `traffic` is the real "
                                                          "DRAM bytes per launch), so this is not a memory roofline"},
                               "policy": "latency" if args.latency_mode else "throughput", "how": how},
        }
        roofs["lattice_kernel_latency_mode"] = {"ms_eager_alone": lat_ms_latency_mode, "policy": "latency (library default)",
                                                "evals_per_s": evals_per_rank / (lat_ms_latency_mode * 1e-3)}
        for r in ("edt_tma_kernel", "lattice_kernel"):
            roofs[r]["frac"] = roofs[r]["achieved"] / roofs[r]["peak"]
            roofs[r]["share_of_step"] = roofs[r]["ms"] / ms_per_step
        dom = max(("edt_tma_kernel", "lattice_kernel"), key=lambda k: roofs[k]["ms"])
        roofline = dict(roofs[dom])
        roofline["kernel"] = dom
        roofline["peak_source"] = peak_src if dom == "edt_tma_kernel" else roofs[dom]["peak_how"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(w, workload, world),          # identical in the reference arm's line
            "timing": {"ring_maps": ring,
                       "how": (f"cuda-graph replay ({turn_len} steps per graph"
                               + (", EDT of step i+1 on a second stream under the match of step i)" if ctx_e else ")")
                               + f"; median of {len(rep_ms)} repetitions of the {K}-step region, each behind a device-side barrier")
                       if use_graph else "eager launches"},
            "reps": spread(rep_ms, K),
            "serial_ms_per_step": serial_ms, "serial_reps": serial_reps,
            "edt_mcells_per_s": cells / (edt_ms * 1e-3) / 1e6,
            "match_evals_per_s_per_gpu": evals_per_rank / (lat_ms * 1e-3),
            "roofline": roofline, "rooflines": roofs,
            "e2e": {"value": total_evals / (e2e_pts_s / KP), "unit": UNIT, "h2d_bytes_per_step": h2d_pts,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_pts_s / KP * 1e3, "steps": KP,
                    "host_buffers": f"the map's {npts} points (x, y f32: what the reference's OccupationalGrid takes, "
                                    "main.c:271) + the scan, page-locked; result read back every step",
                    "calls": "b200slam_map_rasterise_async -> b200slam_map_edt -> b200slam_scan_upload -> "
                             "b200slam_score_lattice_async -> b200slam_match_fetch  (the reference's OccupationalGrid -> "
                             "euclidean_distance_transform -> FastMatch, main.c:884-918)",
                    "how": f"{DEPTH} contexts take turns: the next steps' H2D and rasterisation run under this step's kernels and result D2H",
                    "one_context_ms_per_step": e2e_pts_serial_s / KP * 1e3},
            "e2e_int32_grid": {"value": total_evals / (e2e_s / KE), "unit": UNIT, "h2d_bytes_per_step": h2d,
                               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / KE * 1e3, "steps": KE,
                               "host_buffers": "the rasterised int32 occupancy grid (4 B per cell: PCIe bound) + the scan",
                               "how": "two contexts alternate: the next step's H2D runs under this step's kernels and result D2H",
                               "one_context_ms_per_step": e2e_serial_s / KE * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "result": {"best_index": checks["timed_region"][0], "best_score_bits": checks["timed_region"][1],
                       "best_hits": checks["timed_region"][2], "e2e_best_index": checks["e2e"][0],
                       "verified": bool(verified),
                       "e2e_points_rasterised_grid_identical": bool(raster_ok),
                       "verified_how": f"timed region == strictly sequential graph == e2e from the int32 grid (both forms) == rank 0 "
                                       f"scoring alone; e2e from points (both forms) == rank 0 alone on the grid rasterised from the same "
                                       f"points (== the int32 grid + 2 margin cells, compared cell for cell); all: rank 0 scoring all "
                                       f"{n_global[0] * ntx * nty} candidates alone with synchronous calls, on every one of the "
                                       f"{world} rank(s): index, score bits, winner's and last candidate's hit counts"},
        }
        if world == 1 and want_cpu:
            line["cpu_baseline"] = cpu_sample(w, synth, budget_s=args.cpu_seconds)

    for g in {id(g): g for g in (turn, turn_serial, rem, rem_serial, g_edt_only, g_match_only) if g is not None}.values():
        ctx.graph_destroy(g)
    if ctx_e is not None:
        ctx_e.close()
    for m in maps:
        m.close()
    job.barrier()                         # nobody tears its peer-mapped buffers down while a peer still spins on them
    for m_x in emaps[ring:]:
        m_x.close()
    for c_x in pipe[2:]:
        c_x.close()
    ctx_b.close()
    ctx.close()
    if not verified and not os.environ.get("B200SLAM_BENCH_NO_COLLECT"):
        raise SystemExit(f"[bench] {workload}: the timed region's result differs from the stand-alone result")
    return line


def particles_block(args, synth, job: Job, K):
    """BASELINE.json configs[2]: FastSLAM-style 100 000 particles x 720 beams PER GPU; a step = score every
    particle against the distance field + weights + normalise + systematic resample + offspring gather,
    device resident.  N > 1: the particle set is sharded over the ranks (map replicated); {score min,
    integer weight sum, count} and the offspring travel through NVLink peer memory inside the step, no
    host synchronisation."""
    mod = importlib.import_module(PKG)
    rank, world, local_rank = job.rank, job.world, job.local_rank
    K, W = K + (K & 1), max(args.warmup, 3)      # even: two steps per graph (buffers swap)
    w = synth.make_workload("config1")
    rows, cols = w["occ"].shape
    P, nbeams, beta, u0 = 100000, 720, 0.05, 0x80000000
    ctx = mod.Context(local_rank)
    sharded = world > 1 and hasattr(mod.Context, "particles_shard")
    if sharded:
        job.comm_init(ctx)
    m = ctx.new_map(rows, cols)
    m.set_geometry(w["pixel"], w["top_left"]).upload_occupancy(w["occ"]).edt()
    sx, sy = synth.scan_fixed_count(w["occ"], float(w["pixel"]), w["top_left"], w["true_pose"], nbeams)
    ctx.scan_upload(sx, sy)
    poses_all = synth.particles_gaussian(P * world, w["true_pose"])
    poses = poses_all[rank * P:(rank + 1) * P]
    if sharded:
        ctx.particles_shard(poses, rank * P, P * world)
    else:
        ctx.particles_upload(poses)

    def step():
        ctx.particles_score_async(m)
        ctx.particles_resample_async(beta, u0)

    for _ in range(2 * ((W + 1) // 2)):
        step()
    ctx.sync()
    job.barrier()
    # per-kernel split, eager
    KA = 20
    for i in range(KA):
        ctx.event_record(3 * i); ctx.particles_score_async(m)
        ctx.event_record(3 * i + 1); ctx.particles_resample_async(beta, u0)
        ctx.event_record(3 * i + 2)
    ctx.sync()
    sc_ms = sum(ctx.event_elapsed_ms(3 * i, 3 * i + 1) for i in range(KA)) / KA
    rs_ms = sum(ctx.event_elapsed_ms(3 * i + 1, 3 * i + 2) for i in range(KA)) / KA
    job.barrier()
    ctx.graph_begin(); step(); step(); g = ctx.graph_end()
    launches0 = ctx.launch_count()

    def run_once():
        for _ in range(K // 2):
            ctx.graph_launch(g)

    run_once(); ctx.sync()
    ctx.event_record(4090); run_once(); ctx.event_record(4091); ctx.sync()
    launches_per_rep = (ctx.launch_count() - launches0) // 2
    med_ms, rep_ms = timed_reps(job, ctx, run_once, ctx.event_elapsed_ms(4090, 4091), target_ms=60.0)
    best = ctx.match_fetch()
    # verification: one filter step from a fixed particle set against the unsharded computation on rank 0
    verified = True
    if sharded:
        ctx.particles_shard(poses, rank * P, P * world)
        step()
        got_poses, _, _, got_anc = ctx.particles_download()
        if rank == 0:
            c1 = mod.Context(local_rank)
            m1 = c1.new_map(rows, cols)
            m1.set_geometry(w["pixel"], w["top_left"]).upload_occupancy(w["occ"]).edt()
            c1.scan_upload(sx, sy)
            c1.particles_upload(poses_all)
            c1.particles_score_async(m1); c1.particles_resample_async(beta, u0)
            want_poses, _, _, want_anc = c1.particles_download()
            m1.close(); c1.close()
        else:
            want_poses = want_anc = None
        want_poses, want_anc = job.bcast((want_poses, want_anc))
        sl = slice(rank * P, (rank + 1) * P)
        verified = bool(np.array_equal(got_anc, want_anc[sl]) and
                        np.array_equal(got_poses.view(np.uint32), want_poses[sl].view(np.uint32)))
        verified = job.all_true(verified)
    # e2e: particles from host memory every step, weights + ancestors back
    KE = 10
    job.barrier()
    t0 = time.perf_counter()
    for _ in range(KE):
        ctx.score_poses(m, poses, index_base=rank * P if sharded else 0, want_hits=False)
        ctx.weights_resample(P * world if sharded else P, beta, u0)
    e2e_s = (time.perf_counter() - t0) / KE
    e2e_s, sc_ms, rs_ms = job.max_over_ranks([e2e_s, sc_ms, rs_ms])
    line = None
    if rank == 0:
        evals = P * nbeams
        ms_per_step = med_ms / K
        peak, sm_mhz, peak_src = measured_peaks()
        info = ctx.device_info()
        issue_peak = 4.0 * info["sm_count"] * sm_mhz * 1e6 / 1e9
        floor = ISSUE_FLOOR_LANE_INSTR_PER_EVAL["poses"]
        sc_bytes = 4.0 * evals + 16.0 * P + 8.0 * nbeams            # 4 B / eval + 16 B / pose
        line = {
            "metric": METRIC, "value": evals * world / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config2: FastSLAM-style {P * world} particles x {nbeams} beams on a {rows}x{cols} map: score + "
                                   "weights + normalise + systematic resample + offspring gather, device resident",
                       "particles_per_gpu": P, "beams": nbeams, "beta": beta,
                       "parallelism": (f"particles sharded over {world} GPUs, map replicated; sums and offspring through NVLink "
                                       "peer memory, no host sync per step") if sharded else f"{world} GPU(s)",
                       "l2": "the 16.8 MB distance field is L2 resident by design (gather workload)",
                       "timing": f"cuda-graph replay (2 steps per graph); median of {len(rep_ms)} repetitions of the {K}-step region"},
            "reps": spread(rep_ms, K),
            "roofline": {"bound": "issue", "kernel": "poses_kernel", "achieved": floor * evals / 32.0 / (sc_ms * 1e-3) / 1e9,
                         "peak": issue_peak, "unit": "Gwarp-inst/s",
                         "frac": floor * evals / 32.0 / (sc_ms * 1e-3) / 1e9 / issue_peak, "ms": sc_ms,
                         "floor_lane_instr_per_eval": floor, "traffic": ncu_counter("poses_kernel:config2"),
                         "measured_warp_instructions": ncu_counter("poses_kernel:config2:warp_instructions"),
                         "peak_how": f"4 warp schedulers x {info['sm_count']} SMs x {sm_mhz:.0f} MHz", "resample_ms": rs_ms,
                         "hbm_equivalent_gbs": sc_bytes / (sc_ms * 1e-3) / 1e9,
                         "note": "every particle has its own theta, so the rotation is paid per evaluation: instruction bound"},
            "e2e": {"value": evals * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 12 * P + 8 * nbeams,
                    "d2h_bytes_per_step": 12 * P + 16, "ms_per_step": e2e_s * 1e3, "steps": KE},
            "gpu_launches": int(launches_per_rep),
            "result": {"best_index": int(best.best_index), "best_score": float(best.best_score), "verified": bool(verified),
                       "verified_how": "one filter step: this rank's offspring (ancestors and poses, bit for bit) == the same "
                                       "slice of an unsharded run of all particles on rank 0" if sharded else "single GPU"},
        }
    ctx.graph_destroy(g)
    m.close()
    job.barrier()
    ctx.close()
    if not verified:
        raise SystemExit("[bench] config2: sharded particle filter differs from the unsharded run")
    return line


def pyramid_block(args, synth, job: Job, K):
    """BASELINE.json configs[4]: multi-resolution correlative search over a 3-level EDT pyramid (2048^2 @ 4p,
    4096^2 @ 2p, 8192^2 @ p; coarsest first), 10 M coarse poses over 8 GPUs = 20 x 250 x 250 = 1.25 M per
    GPU (weak scaling: the theta range grows with N), then two 16 x 32 x 32 refinements, each level seeded
    by the previous winner (Subsystem_1/main.c:901-918 generalised).  A step = the three transforms + the
    three matches (b200slam_pyramid_match: candidate rows of every level sharded over the ranks, the
    per-rank bests exchanged inside the kernels' tails)."""
    mod = importlib.import_module(PKG)
    rank, world, local_rank = job.rank, job.world, job.local_rank
    W = max(args.warmup, 3)
    w = synth.make_workload("config3")
    occ_f = w["occ"]
    nbeams = len(w["scan_x"])
    ctx = mod.Context(local_rank)
    job.comm_init(ctx)
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
    holder = {}

    def step():
        for m in maps:
            m.edt(10.0)
        holder["res"] = ctx.pyramid_match(maps, w["pose0"], steps, ns)

    def run_once():
        for _ in range(K):
            step()

    for _ in range(W):
        step()
    # dominant kernel: this rank's share of the coarsest level, an event pair around each launch
    KA = 10
    rb, re = rank * 20 * 250, (rank + 1) * 20 * 250
    ctx.sync(); job.barrier()
    for i in range(KA):
        ctx.event_record(2 * i)
        ctx.score_lattice_async(maps[0], w["pose0"], steps[0], tuple(int(x) for x in ns[0]), rb, re, False)
        ctx.event_record(2 * i + 1)
    ctx.sync()
    lat_ms = sum(ctx.event_elapsed_ms(2 * i, 2 * i + 1) for i in range(KA)) / KA
    launches0 = ctx.launch_count()
    ctx.event_record(4090); run_once(); ctx.event_record(4091); ctx.sync()
    launches = ctx.launch_count() - launches0
    med_ms, rep_ms = timed_reps(job, ctx, run_once, ctx.event_elapsed_ms(4090, 4091), target_ms=60.0, min_reps=5)
    res = holder["res"]
    # verification: rank 0 alone, the whole lattice of every level, plain synchronous calls
    got = [match_tuple(r) for r in res]
    want = None
    if rank == 0:
        want, seed = [], np.array(w["pose0"], np.float32)
        for l in range(3):
            nl = tuple(int(x) for x in ns[l])
            r = ctx.score_lattice_rows(maps[l], seed, steps[l], nl, 0, nl[0] * nl[1], False)
            want.append(match_tuple(r))
            seed = r.pose()
    want = job.bcast(want)
    verified = job.all_true(got == want)
    # e2e: the three occupancy grids and the scan from pinned host memory every step
    KE = max(3, min(K, 5))
    ctx.sync(); job.barrier()
    t0 = time.perf_counter()
    for _ in range(KE):
        for m, pin in zip(maps, pins):
            m.upload_occupancy(pin)
        ctx.scan_upload(scan_x, scan_y)
        step()
    ctx.sync()
    e2e_s = (time.perf_counter() - t0) / KE
    job.barrier()
    e2e_s, lat_ms = job.max_over_ranks([e2e_s, lat_ms])
    line = None
    if rank == 0:
        ms_per_step = med_ms / K
        peak, sm_mhz, peak_src = measured_peaks()
        info = ctx.device_info()
        issue_peak = 4.0 * info["sm_count"] * sm_mhz * 1e6 / 1e9
        floor = ISSUE_FLOOR_LANE_INSTR_PER_EVAL["lattice_rr2"]
        coarse_evals = 20 * 250 * 250 * nbeams
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
                       "timing": f"eager launches, one host round trip per level (the next level is seeded by the winner); "
                                 f"median of {len(rep_ms)} repetitions of the {K}-step region"},
            "reps": spread(rep_ms, K),
            "edt_mcells_per_step": cells / 1e6,
            "roofline": {"bound": "issue", "kernel": "lattice_kernel", "achieved": floor * coarse_evals / 32.0 / (lat_ms * 1e-3) / 1e9,
                         "peak": issue_peak, "unit": "Gwarp-inst/s",
                         "frac": floor * coarse_evals / 32.0 / (lat_ms * 1e-3) / 1e9 / issue_peak, "ms": lat_ms,
                         "floor_lane_instr_per_eval": floor, "evals_per_s": coarse_evals / (lat_ms * 1e-3),
                         "peak_how": f"4 warp schedulers x {info['sm_count']} SMs x {sm_mhz:.0f} MHz",
                         "note": "coarsest level, this rank's 20 x 250 x 250 share"},
            "e2e": {"value": evals_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": cells * 4 + 8 * nbeams,
                    "d2h_bytes_per_step": 3 * 16, "ms_per_step": e2e_s * 1e3, "steps": KE},
            "gpu_launches": int(launches),
            "result": {"best_index": [g[0] for g in got], "verified": bool(verified),
                       "verified_how": "every level's exchanged winner on every rank == rank 0 scoring the whole level alone"},
        }
    for m in maps:
        m.close()
    job.barrier()
    ctx.close()
    if not verified:
        raise SystemExit("[bench] config4: sharded pyramid result differs from the stand-alone result")
    return line


def edt_sharded_block(args, synth, job: Job, size=8192, iters=20):
    """configs[3]'s "8192x8192 grid EDT (row-sharded)" against replicated compute, N > 1: (a) the
    whole transform on every GPU, (b) own row block + in-place ncclAllGather, (c) own row block with the EDT
    kernel itself storing every row into every peer's field over NVLink.  Every variant is compared with
    (a) bit for bit first; CUDA events on the library's stream, max over ranks."""
    mod = importlib.import_module(PKG)
    rank, world, local = job.rank, job.world, job.local_rank
    ctx = mod.Context(local)
    job.comm_init(ctx)
    S = size
    maps = []
    for i in range(2):                       # two maps: 2 x 512 MiB > L2, alternate between launches
        m = ctx.new_map(S, S)
        m.upload_occupancy(synth.grid_rooms(S, S, synth.SEED_GRID + i))
        maps.append(m)
    want = [m.edt().download_field() for m in maps]
    shared = True
    try:
        for m in maps:
            m.share()
    except mod.B200SlamError as e:
        shared = False
        if rank == 0:
            print(f"[bench] peer-shared maps unavailable: {e}", file=sys.stderr, flush=True)

    def timed(fn):
        for i in range(3):
            fn(maps[i % 2])
        ctx.sync(); job.barrier()
        ctx.comm_barrier_async()
        ctx.event_record(0)
        for i in range(iters):
            fn(maps[i % 2])
        ctx.event_record(1)
        ctx.sync(); job.barrier()
        return job.max_over_ranks([ctx.event_elapsed_ms(0, 1) / iters])[0]

    out = {"size": S, "n_gpus": world, "iters": iters, "verified": True}
    out["replicated_ms"] = timed(lambda m: m.edt())
    variants = [("sharded_nccl_ms", mod.EDT_GATHER_NCCL)] + ([("sharded_p2p_fused_ms", mod.EDT_GATHER_P2P)] if shared else [])
    ok = True
    for name, mode in variants:
        for m, wf in zip(maps, want):
            m.upload_field(np.full((S, S), -1.0, np.float32))
            m.edt_sharded(mode)
            ok = ok and np.array_equal(m.download_field().view(np.uint32), wf.view(np.uint32))
        out[name] = timed(lambda m, mode=mode: m.edt_sharded(mode))
    rb, re = mod.shard_range(S, world, rank)
    out["own_block_only_ms"] = timed(lambda m: m.edt_rows(rb, re))
    out["verified"] = job.all_true(ok)
    out["field_bytes"] = S * S * 4
    out["received_bytes_per_gpu"] = S * S * 4 * (world - 1) // world
    job.barrier()
    for m in maps:
        m.close()
    ctx.close()
    if not out["verified"]:
        raise SystemExit("[bench] row-sharded EDT differs from the replicated transform")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config1", "config2", "config3", "config4", "tiny"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="keep every step's EDT and match strictly back to back (no second stream)")
    ap.add_argument("--latency-mode", action="store_true",
                    help="keep the library's default tile-shape policy (one match at a time) in the timed region too")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the blocks of the other configurations")
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
        return
    job = Job()
    K = args.steps
    keep = ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "reps", "serial_ms_per_step", "config", "timing",
            "edt_mcells_per_s", "match_evals_per_s_per_gpu", "roofline", "rooflines", "e2e", "e2e_int32_grid", "gpu_launches", "result")
    if args.workload == "config2":
        line = particles_block(args, synth, job, K)
    elif args.workload == "config4":
        line = pyramid_block(args, synth, job, K)
    else:
        line = measure_workload(args, synth, args.workload, job, K, want_cpu=not args.no_cpu)
        if args.workload == "config3" and not args.no_extra:
            # the other configurations, same run (driver-visible at every N)
            extra = {}
            c1 = measure_workload(args, synth, "config1", job, max(K, 100), want_cpu=False, sample_clocks=False)
            c2 = particles_block(args, synth, job, max(K, 40))
            c4 = pyramid_block(args, synth, job, max(4, min(K, 10)))
            if job.rank == 0:
                extra["config1"] = {k: c1[k] for k in keep if k in c1}
                extra["config2"] = {k: c2[k] for k in keep if k in c2}
                extra["config4"] = {k: c4[k] for k in keep if k in c4}
            if job.world > 1:
                es = edt_sharded_block(args, synth, job)
                if job.rank == 0:
                    extra["edt_row_sharded"] = es
            if job.rank == 0:
                line.update(extra)
                line["gpu_launches_all_blocks"] = int(line["gpu_launches"] + sum(extra[k].get("gpu_launches", 0) for k in
                                                                                   ("config1", "config2", "config4")))
    if job.rank == 0:
        print(json.dumps(line), flush=True)
    job.close()


if __name__ == "__main__":
    main()
