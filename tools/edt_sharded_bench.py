#!/usr/bin/env python
"""tools/edt_sharded_bench.py -- config 3's "row-sharded EDT" measured against replicated compute.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29611 tools/edt_sharded_bench.py [--size 8192]

Per rank, same 8192^2 map: (a) the whole transform on every GPU (no exchange), (b) each rank's row block
+ in-place ncclAllGather of the blocks, (c) each rank's row block with the EDT kernel itself storing
every row into every peer's field over NVLink, bracketed by two device barriers.  CUDA events on the
library's stream, max over ranks; every variant is checked against (a) bit for bit first.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "hardware-acceleration-of-lidar-slam_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mod = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    ctx = mod.Context(local)
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    S = args.size
    occ = synth.grid_rooms(S, S, synth.SEED_GRID)
    maps = []
    for i in range(2):                       # two maps: 2 x 512 MiB > L2, alternate between launches
        m = ctx.new_map(S, S)
        m.upload_occupancy(occ if i == 0 else synth.grid_rooms(S, S, synth.SEED_GRID + 1))
        maps.append(m)
    want = [m.edt().download_field() for m in maps]
    shared = True
    try:
        for m in maps:
            m.share()
    except mod.B200SlamError as e:
        shared = False
        if rank == 0:
            print(f"peer-shared maps unavailable: {e}", flush=True)

    def timed(fn):
        for i in range(3):
            fn(maps[i % 2])
        ctx.sync(); dist.barrier()
        ctx.event_record(0)
        for i in range(args.iters):
            fn(maps[i % 2])
        ctx.event_record(1)
        ctx.sync(); dist.barrier()
        t = torch.tensor([ctx.event_elapsed_ms(0, 1) / args.iters], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"size": S, "n_gpus": world, "iters": args.iters}
    out["replicated_ms"] = timed(lambda m: m.edt())
    variants = [("sharded_nccl_ms", mod.EDT_GATHER_NCCL)] + ([("sharded_p2p_fused_ms", mod.EDT_GATHER_P2P)] if shared else [])
    for name, mode in variants:
        for m, w in zip(maps, want):
            m.upload_field(np.full((S, S), -1.0, np.float32))
            m.edt_sharded(mode)
            ok = np.array_equal(m.download_field().view(np.uint32), w.view(np.uint32))
            assert ok, f"rank {rank}: {name} differs from the replicated transform"
        out[name] = timed(lambda m, mode=mode: m.edt_sharded(mode))
    rb, re = mod.shard_range(S, world, rank)
    out["own_block_only_ms"] = timed(lambda m: m.edt_rows(rb, re))
    out["field_bytes"] = S * S * 4
    out["received_bytes_per_gpu"] = S * S * 4 * (world - 1) // world
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    for m in maps:
        m.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
