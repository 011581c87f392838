"""Per-shard lattice time on ONE GPU: is a rank's step time a function of which theta rows it owns?"""
import importlib, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mod = importlib.import_module("hardware-acceleration-of-lidar-slam_b200")
synth = importlib.import_module("hardware-acceleration-of-lidar-slam_b200.synth")
from tools.kbench import time_loop
with mod.Context(0) as ctx:
    w = synth.make_workload("config1")
    rows, cols = w["occ"].shape
    m = ctx.new_map(rows, cols)
    m.set_geometry(w["pixel"], w["top_left"]).upload_occupancy(w["occ"]).edt()
    ctx.scan_upload(w["scan_x"], w["scan_y"])
    for label, n in (("theta x8", (512, 32, 32)), ("tx x8", (64, 256, 32))):
        nrows = n[0] * n[1]
        for r in range(8):
            rb, re = mod.shard_range(nrows, 8, r)
            ms = time_loop(ctx, lambda i: ctx.score_lattice_async(m, w["pose0"], w["step"], n, rb, re), 50)
            res = ctx.match_fetch()
            print(f"{label}: rank {r} rows [{rb},{re}) {ms * 1e3:7.2f} us  best={res.best_index} hits={res.best_hits}", flush=True)
    m.close()
