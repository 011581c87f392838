#!/usr/bin/env python
"""tools/kbench.py -- per-kernel timings on the GPU box (development aid, not the bench contract).

    python tools/kbench.py edt      # EDT on 400^2 / 2048^2 / 8192^2, rooms and bernoulli grids
    python tools/kbench.py lattice  # lattice matcher at the BASELINE.json shapes
    python tools/kbench.py poses    # pose-list scorer + weights/resample (config 2)

Times with CUDA events on the library's stream; inputs rotate through a ring larger than L2.
"""
from __future__ import annotations

import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "hardware-acceleration-of-lidar-slam_b200"
PEAK = 6542.7
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def time_loop(ctx, fn, iters, warm=3):
    for i in range(warm):
        fn(i)
    ctx.sync()
    ctx.event_record(0)
    for i in range(iters):
        fn(i)
    ctx.event_record(1)
    ctx.sync()
    return ctx.event_elapsed_ms(0, 1) / iters


def bench_edt(mod, synth, ctx, sizes):
    for rows, cols in sizes:
        cells = rows * cols
        ring = max(2, min(12, (2 * 126 * 2 ** 20) // (cells * 8) + 1))
        for kind in ("rooms", "bern0.01", "bern0.05", "empty"):
            maps = []
            for i in range(ring):
                if kind == "rooms":
                    occ = synth.grid_rooms(rows, cols, synth.SEED_GRID + i)
                elif kind == "empty":
                    occ = np.zeros((rows, cols), np.int32)
                else:
                    occ = synth.grid_bernoulli(rows, cols, float(kind[4:]), synth.SEED_GRID + i)
                m = ctx.new_map(rows, cols)
                m.upload_occupancy(occ)
                maps.append(m)
                if kind != "rooms" and i >= 1 and cells >= 2 ** 26:
                    break
            n = len(maps)
            iters = 50 if cells <= 2 ** 22 else 20
            ms = time_loop(ctx, lambda i: maps[i % n].edt(10.0), iters)
            gbs = cells * 8 / (ms * 1e-3) / 1e9
            print(f"edt {rows}x{cols} {kind:9s} ring={n:2d}  {ms * 1e3:9.2f} us  {cells / ms / 1e3:10.1f} Mcells/s  "
                  f"{gbs:8.1f} GB/s  frac={gbs / PEAK:.3f}", flush=True)
            for m in maps:
                m.close()


def bench_edt_cb(mod, synth, ctx):
    """Chunk-height sweep (B200SLAM_EDT_CB, read at every launch) of the 8192^2 / 2048^2 transform, back to back
    (consecutive transforms overlap through programmatic dependent launch) and isolated (event pair per launch)."""
    for rows in (2048, 8192):
        cells = rows * rows
        ring = 8 if rows == 2048 else 2
        for kind in ("rooms", "bern0.05"):
            maps = []
            for i in range(ring):
                occ = synth.grid_rooms(rows, rows, synth.SEED_GRID + i) if kind == "rooms" else \
                    synth.grid_bernoulli(rows, rows, 0.05, synth.SEED_GRID + i)
                m = ctx.new_map(rows, rows)
                m.upload_occupancy(occ)
                maps.append(m)
            for cb in os.environ.get("SWEEP", "0 3 4 5 7 9 14 20").split():
                if cb == "0":
                    os.environ.pop("B200SLAM_EDT_CB", None)
                else:
                    os.environ["B200SLAM_EDT_CB"] = cb
                ms = time_loop(ctx, lambda i: maps[i % ring].edt(10.0), 40)
                K = 30
                for i in range(K):
                    ctx.event_record(100 + 2 * i); maps[i % ring].edt(10.0); ctx.event_record(101 + 2 * i)
                ctx.sync()
                iso = sorted(ctx.event_elapsed_ms(100 + 2 * i, 101 + 2 * i) for i in range(K))[K // 2]
                print(f"edt {rows}^2 {kind:8s} cb={cb:>2s}: back-to-back {ms * 1e3:8.2f} us (frac {cells * 8 / ms / 1e6 / PEAK:.3f})   "
                      f"isolated median {iso * 1e3:8.2f} us (frac {cells * 8 / iso / 1e6 / PEAK:.3f})", flush=True)
            os.environ.pop("B200SLAM_EDT_CB", None)
            for m in maps:
                m.close()


def bench_lattice(mod, synth, ctx, names):
    for name in names:
        w = synth.make_workload(name)
        rows, cols = w["occ"].shape
        m = ctx.new_map(rows, cols)
        m.set_geometry(w["pixel"], w["top_left"]).upload_occupancy(w["occ"]).edt()
        ctx.scan_upload(w["scan_x"], w["scan_y"])
        n = w["n"]
        evals = n[0] * n[1] * n[2] * len(w["scan_x"])
        ms = time_loop(ctx, lambda i: ctx.score_lattice_async(m, w["pose0"], w["step"], n), 20 if evals > 1e9 else 100)
        r = ctx.match_fetch()
        print(f"lattice {name} {n} x {len(w['scan_x'])} beams: {ms * 1e3:9.2f} us  {evals / ms / 1e9:8.3f} Tevals/s "
              f"best={r.best_index} score={r.best_score:.3f}", flush=True)
        m.close()


def bench_lattice_sweep(mod, synth, ctx, name="config1"):
    """Tile-shape sweep of the lattice matcher through B200SLAM_LATTICE_CFG (read at every launch)."""
    w = synth.make_workload(name)
    rows, cols = w["occ"].shape
    m = ctx.new_map(rows, cols)
    m.set_geometry(w["pixel"], w["top_left"]).upload_occupancy(w["occ"]).edt()
    ctx.scan_upload(w["scan_x"], w["scan_y"])
    n = w["n"]
    evals = n[0] * n[1] * n[2] * len(w["scan_x"])
    ref = None
    for cfg in os.environ.get("SWEEP", "1,1,8 2,1,8,2 2,1,8 4,1,8 8,1,8 16,2,4 1,1,4").split():
        os.environ["B200SLAM_LATTICE_CFG"] = cfg
        ms = time_loop(ctx, lambda i: ctx.score_lattice_async(m, w["pose0"], w["step"], n), 200)
        # isolated: an event pair around every launch (no overlap with the neighbours)
        K = 50
        for i in range(K):
            ctx.event_record(100 + 2 * i); ctx.score_lattice_async(m, w["pose0"], w["step"], n); ctx.event_record(101 + 2 * i)
        ctx.sync()
        iso = sorted(ctx.event_elapsed_ms(100 + 2 * i, 101 + 2 * i) for i in range(K))[K // 2]
        r = ctx.match_fetch()
        ref = ref or (r.best_index, r.best_score)
        print(f"lattice {name} cfg={cfg:7s}: back-to-back {ms * 1e3:8.2f} us  isolated median {iso * 1e3:8.2f} us  "
              f"{evals / ms / 1e9:7.3f} Tevals/s  same_result={(r.best_index, r.best_score) == ref}", flush=True)
    os.environ.pop("B200SLAM_LATTICE_CFG", None)
    m.close()


def bench_pipeline_ab(mod, synth, ctx, name="config1"):
    """A/B of tile shapes inside bench.py's pipelined step (EDT of step i+1 on a second stream under the
    match of step i, one CUDA graph per turn of the map ring), all in one process, interleaved twice."""
    w = synth.make_workload(name)
    rows, cols = w["occ"].shape
    n = w["n"]
    ring = 9 if rows <= 2048 else 2
    maps = []
    for i in range(ring):
        m = ctx.new_map(rows, cols)
        m.set_geometry(w["pixel"], w["top_left"]).upload_occupancy(
            w["occ"] if i == 0 else synth.grid_rooms(rows, cols, synth.SEED_GRID + i))
        maps.append(m)
    ctx.scan_upload(w["scan_x"], w["scan_y"])
    ctx_e = mod.Context(0)
    evals = n[0] * n[1] * n[2] * len(w["scan_x"])
    cfgs = os.environ.get("SWEEP", "1,1,8 2,1,8,2 4,1,8 2,1,8").split()
    graphs = {}
    for cfg in cfgs:
        os.environ["B200SLAM_LATTICE_CFG"] = cfg.split("/")[0]
        os.environ.pop("B200SLAM_EDT_CB", None)
        if "/" in cfg:
            os.environ["B200SLAM_EDT_CB"] = cfg.split("/")[1]
        for i in range(ring):
            maps[i].edt(10.0); ctx.score_lattice_async(maps[i], w["pose0"], w["step"], n)
            ctx_e._check(ctx_e.L.b200slam_map_edt(ctx_e.h, maps[i].h, 10.0))
        ctx.sync(); ctx_e.sync()
        ctx.graph_begin()
        for i in range(ring):
            maps[i].edt(10.0); ctx.score_lattice_async(maps[i], w["pose0"], w["step"], n)
        gs = ctx.graph_end()
        ctx.graph_begin()
        ctx.event_record(3000); ctx_e.event_wait(ctx, 3000)
        for i in range(ring):
            ctx_e._check(ctx_e.L.b200slam_map_edt(ctx_e.h, maps[i].h, 10.0))
            ctx_e.event_record(3100 + i); ctx.event_wait(ctx_e, 3100 + i)
            ctx.score_lattice_async(maps[i], w["pose0"], w["step"], n)
        ctx_e.event_record(3200); ctx.event_wait(ctx_e, 3200)
        gp = ctx.graph_end()
        graphs[cfg] = (gs, gp)
    os.environ.pop("B200SLAM_LATTICE_CFG", None)
    os.environ.pop("B200SLAM_EDT_CB", None)
    for rep in range(2):
        for cfg in cfgs:
            gs, gp = graphs[cfg]
            ser = time_loop(ctx, lambda i: ctx.graph_launch(gs), 30) / ring
            pip = time_loop(ctx, lambda i: ctx.graph_launch(gp), 30) / ring
            print(f"pipeline {name} rep {rep} cfg={cfg:11s}: serial {ser * 1e3:7.2f} us/step  pipelined {pip * 1e3:7.2f} us/step "
                  f"({evals / pip / 1e9:6.3f} Tevals/s)", flush=True)
    ctx_e.close()
    for m in maps:
        m.close()


def bench_poses(mod, synth, ctx):
    w = synth.make_workload("config1")
    rows, cols = w["occ"].shape
    m = ctx.new_map(rows, cols)
    m.set_geometry(w["pixel"], w["top_left"]).upload_occupancy(w["occ"]).edt()
    x, y = synth.scan_fixed_count(w["occ"], float(w["pixel"]), w["top_left"], w["true_pose"], 720)
    ctx.scan_upload(x, y)
    P = 100000
    poses = synth.particles_gaussian(P, w["true_pose"])
    import time
    for _ in range(3):
        ctx.score_poses(m, poses, want_hits=False)
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.score_poses(m, poses, want_hits=False)
    t1 = time.perf_counter()
    for _ in range(10):
        ctx.weights_resample(P, 0.05, 0x80000000, want_weights=True)
    t2 = time.perf_counter()
    # device-resident filter step: score + weights + normalise + resample + gather, two steps per graph
    ctx.particles_upload(poses)
    for _ in range(2):
        ctx.particles_score_async(m); ctx.particles_resample_async(0.05, 0x80000000)
    ctx.sync()
    ctx.graph_begin()
    for _ in range(2):
        ctx.particles_score_async(m); ctx.particles_resample_async(0.05, 0x80000000)
    g = ctx.graph_end()
    ms = time_loop(ctx, lambda i: ctx.graph_launch(g), 20) / 2
    ctx.event_record(10); ctx.particles_score_async(m); ctx.event_record(11); ctx.particles_resample_async(0.05, 0x80000000)
    ctx.event_record(12); ctx.sync()
    print(f"particles 100k x 720 device-resident step: {ms * 1e3:.1f} us  {P * 720 / ms / 1e9:.3f} Tevals/s "
          f"(eager: score {ctx.event_elapsed_ms(10, 11) * 1e3:.1f} us, weights+resample+gather {ctx.event_elapsed_ms(11, 12) * 1e3:.1f} us)", flush=True)
    ctx.graph_destroy(g)
    print(f"poses 100k x 720 (host call incl. H2D/D2H): {(t1 - t0) / 10 * 1e3:.3f} ms  "
          f"{P * 720 / ((t1 - t0) / 10) / 1e9:.2f} Gevals/s; weights+resample {(t2 - t1) / 10 * 1e3:.3f} ms", flush=True)
    m.close()


def main():
    what = sys.argv[1:] or ["edt", "lattice", "poses"]
    mod = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    with mod.Context(0) as ctx:
        print(ctx.device_info(), flush=True)
        if "edt" in what:
            bench_edt(mod, synth, ctx, [(400, 400), (2048, 2048), (8192, 8192)])
        if "edtcb" in what:
            bench_edt_cb(mod, synth, ctx)
        if "edtbig" in what:
            bench_edt(mod, synth, ctx, [(8192, 8192)])
        if "lattice" in what:
            bench_lattice(mod, synth, ctx, ["tiny", "config1", "config3"])
        if "latcfg" in what:                 # config 3 under each B200SLAM_LATTICE_CFG in $SWEEP
            for cfg in os.environ.get("SWEEP", "16,2,4,1,2 8,1,8,1,2 4,1,8,1,2 16,2,4").split():
                os.environ["B200SLAM_LATTICE_CFG"] = cfg
                print("cfg", cfg, flush=True)
                bench_lattice(mod, synth, ctx, ["config3"])
            os.environ.pop("B200SLAM_LATTICE_CFG", None)
        if "latbig" in what:
            for q in ("1", ""):
                if q:
                    os.environ["B200SLAM_LATTICE_NO_RR"] = q
                else:
                    os.environ.pop("B200SLAM_LATTICE_NO_RR", None)
                print("row reuse", "off" if q else "on", flush=True)
                bench_lattice(mod, synth, ctx, ["config1", "config3"])
        if "pipe" in what:
            bench_pipeline_ab(mod, synth, ctx)
        if "latsweep" in what:
            bench_lattice_sweep(mod, synth, ctx)
        if "poses" in what:
            bench_poses(mod, synth, ctx)


if __name__ == "__main__":
    main()
