#!/usr/bin/env python
"""tools/xchg_diag.py -- where the multi-GPU exchange costs time (development aid; torchrun, N >= 2)."""
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "hardware-acceleration-of-lidar-slam_b200"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mod = importlib.import_module(PKG); synth = importlib.import_module(PKG + ".synth")
ctx = mod.Context(local)
uid = [ctx.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0); ctx.comm_init(world, rank, uid[0])
w = synth.make_workload("config1"); rows, cols = w["occ"].shape; nth, ntx, nty = w["n"]
ng = (nth * world, ntx, nty); rb, re = rank * nth * ntx, (rank + 1) * nth * ntx
ring = 9
maps = []
for i in range(ring):
    m = ctx.new_map(rows, cols); m.set_geometry(w["pixel"], w["top_left"]).upload_occupancy(w["occ"] if i == 0 else synth.grid_rooms(rows, cols, synth.SEED_GRID + i)); maps.append(m)
ctx.scan_upload(w["scan_x"], w["scan_y"]); ctx.set_match_mode(mod.MATCH_THROUGHPUT)
ctx_e = mod.Context(local)
for i in range(ring):
    maps[i].edt(10.0); ctx.score_lattice_async(maps[i], w["pose0"], w["step"], ng, rb, re, 3)
    ctx_e._check(ctx_e.L.b200slam_map_edt(ctx_e.h, maps[i].h, 10.0))
ctx.exchange_collect_async(); ctx.sync(); ctx_e.sync()

def capture(post, collect, stamp=False):
    ctx.graph_begin(); ctx.event_record(3000); ctx_e.event_wait(ctx, 3000)
    for i in range(ring):
        ctx_e._check(ctx_e.L.b200slam_map_edt(ctx_e.h, maps[i].h, 10.0)); ctx_e.event_record(3100 + i); ctx.event_wait(ctx_e, 3100 + i)
        ctx.score_lattice_async(maps[i], w["pose0"], w["step"], ng, rb, re, post)
    if collect: ctx.exchange_collect_async()
    ctx_e.event_record(3200); ctx.event_wait(ctx_e, 3200)
    return ctx.graph_end()

def timeit(g, n=30, barrier_each=False):
    for _ in range(3): ctx.graph_launch(g)
    ctx.sync(); dist.barrier()
    ctx.event_record(0)
    for _ in range(n):
        ctx.graph_launch(g)
        if barrier_each: ctx.sync()
    ctx.event_record(1); ctx.sync(); dist.barrier()
    return ctx.event_elapsed_ms(0, 1) / n / ring * 1e3

g0 = capture(0, False); t0 = timeit(g0)
g3 = capture(3, True); t3 = timeit(g3)
# posts recorded but the collect outside the graph, once per TWO turns would overflow: collect eagerly after each launch
g3n = capture(3, False)
for _ in range(3): ctx.graph_launch(g3n); ctx.exchange_collect_async()
ctx.sync(); dist.barrier(); ctx.event_record(0)
for _ in range(30): ctx.graph_launch(g3n); ctx.exchange_collect_async()
ctx.event_record(1); ctx.sync(); dist.barrier(); t3n = ctx.event_elapsed_ms(0, 1) / 30 / ring * 1e3
# the collect kernel alone (nothing pending)
ctx.sync(); dist.barrier(); ctx.event_record(0)
for _ in range(50): ctx.exchange_collect_async()
ctx.event_record(1); ctx.sync(); tc = ctx.event_elapsed_ms(0, 1) / 50 * 1e3
# one burst eagerly, events around the collect
for rep in range(3):
    dist.barrier()
    for i in range(ring): ctx.score_lattice_async(maps[i], w["pose0"], w["step"], ng, rb, re, 3)
    ctx.event_record(10); ctx.exchange_collect_async(); ctx.event_record(11); ctx.sync()
    tcb = ctx.event_elapsed_ms(10, 11) * 1e3
print(f"[rank {rank}] us/step: no exchange {t0:.2f} | deferred posts + collect in graph {t3:.2f} | collect outside graph {t3n:.2f} | "
      f"empty collect kernel {tc:.2f} us | collect after a burst of {ring}: {tcb:.2f} us", flush=True)
dist.barrier(); ctx_e.close()
for m in maps: m.close()
ctx.close(); dist.destroy_process_group()
