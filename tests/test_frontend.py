"""Scan front end and map points: readAScan / Transform / ExtractLocalMap /
map growth, Subsystem_1/main.c:71-118, 155-198, 942-948.

CPU: the oracle restatement against fixtures produced by the reference's own functions
(tests/golden/make_frontend_golden.py) and, where oracle/_ref exists, against those functions live.
GPU: the device front end through the C ABI against the same fixtures and the oracle -- bit-exact,
order included (indices j of the compacted scan carry meaning downstream, main.c:944-945)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits


@pytest.fixture(scope="module")
def fg():
    return np.load(os.path.join(GOLDEN, "frontend_golden.npz"))


def _same(a, b):
    return len(a) == len(b) and np.array_equal(bits(a), bits(b))


def test_oracle_front_end_matches_reference_golden(oracle, fg):
    assert _same(oracle.lidar_angles(), fg["angles"])
    for k in range(int(fg["count"])):
        x, y = oracle.read_scan(fg[f"ranges_{k}"], fg["angles"], float(fg["range_min"]), 24)
        assert _same(x, fg[f"x_{k}"]) and _same(y, fg[f"y_{k}"]), f"case {k} readAScan"
        tx, ty = oracle.transform(x, y, fg[f"pose_{k}"])
        assert _same(tx, fg[f"tx_{k}"]) and _same(ty, fg[f"ty_{k}"]), f"case {k} Transform"
        lx, ly = oracle.extract_local_map(tx, ty, fg[f"map_x_{k}"], fg[f"map_y_{k}"], float(fg[f"border_{k}"]))
        assert _same(lx, fg[f"local_x_{k}"]) and _same(ly, fg[f"local_y_{k}"]), f"case {k} ExtractLocalMap"


def test_oracle_front_end_matches_live_reference(oracle, synth):
    from oracle import pyoracle
    if not pyoracle.reference_available():
        pytest.skip("oracle/_ref not built")
    for which in ("accel", "main"):
        ref = pyoracle.Reference(which)
        angles, rmin = ref.lidar_angles()
        assert _same(oracle.lidar_angles(), angles)
        ranges = synth.lidar_dataset(12, seed=synth.SEED_SCAN + 9)
        for s in (0, 5, 11):
            r = ranges[s].copy()
            r[s::17] = 31.0
            x, y = ref.read_a_scan(r, 24)
            ox, oy = oracle.read_scan(r, angles, rmin, 24)
            assert _same(x, ox) and _same(y, oy)
            pose = np.array([0.1 * s, -0.2 * s, 0.05 * s - 0.2], np.float32)
            tx, ty = ref.transform(pose)
            otx, oty = oracle.transform(ox, oy, pose)
            assert _same(tx, otx) and _same(ty, oty)
            u = synth.hash_uniform(0xBEE + s, np.arange(2 * 15000)).reshape(2, 15000)
            mx, my = ((u[0] - 0.5) * 50).astype(np.float32), ((u[1] - 0.5) * 50).astype(np.float32)
            lx, ly = ref.extract_local_map(mx, my, 1.0)
            olx, oly = oracle.extract_local_map(otx, oty, mx, my, 1.0)
            assert _same(lx, olx) and _same(ly, oly)


def test_oracle_grow_map_definition(oracle):
    # main.c:942-948: strict `> 1.5` on the (double-promoted) hit value, j below bestHits_size only
    hits = np.array([0.0, 1.5, 1.5000001, 10.0, 2.0, 9.0], np.float32)
    tx = np.arange(6, dtype=np.float32); ty = -tx
    mx, my = oracle.grow_map(hits, 5, tx, ty, np.array([7.0], np.float32), np.array([8.0], np.float32))
    assert mx.tolist() == [7.0, 2.0, 3.0, 4.0] and my.tolist() == [8.0, -2.0, -3.0, -4.0]


@pytest.mark.gpu
def test_device_front_end_matches_golden(ctx, fg):
    ctx.lidar_set(fg["angles"], float(fg["range_min"]))
    for k in range(int(fg["count"])):
        n = ctx.scan_read(fg[f"ranges_{k}"], 24)
        assert n == len(fg[f"x_{k}"])
        ctx.scan_transform(fg[f"pose_{k}"])
        x, y, tx, ty = ctx.scan_download()
        assert _same(x, fg[f"x_{k}"]) and _same(y, fg[f"y_{k}"]), f"case {k} readAScan"
        assert _same(tx, fg[f"tx_{k}"]) and _same(ty, fg[f"ty_{k}"]), f"case {k} Transform"
        ctx.mappoints_upload(fg[f"map_x_{k}"], fg[f"map_y_{k}"])
        m = ctx.local_map_extract(float(fg[f"border_{k}"]))
        lx, ly = ctx.local_map_download()
        assert m == len(lx) and _same(lx, fg[f"local_x_{k}"]) and _same(ly, fg[f"local_y_{k}"]), f"case {k} ExtractLocalMap"


@pytest.mark.gpu
def test_device_chain_scan_to_pose_and_growth(ctx, oracle, synth, fg):
    """ranges -> scan -> Transform -> Initialise -> ExtractLocalMap -> OccupationalGrid (both levels) -> EDT ->
    FastMatch / FastMatch2 -> map growth, all device resident, against the oracle chain (each stage of
    which is pinned to the reference)."""
    angles, rmin = fg["angles"], float(fg["range_min"])
    ranges = synth.lidar_dataset(30, seed=synth.SEED_SCAN + 21)
    ctx.lidar_set(angles, rmin)
    pose0 = np.zeros(3, np.float32)
    # scan 0 initialises the map (main.c:846-849)
    ctx.scan_read(ranges[0], 24); ctx.scan_transform(pose0); ctx.mappoints_from_scan()
    ox, oy = oracle.read_scan(ranges[0], angles, rmin, 24)
    omx, omy = oracle.transform(ox, oy, pose0)
    gm = [ctx.new_map(200, 200), ctx.new_map(400, 400)]
    obest = np.zeros(2500, np.float32)          # FastMatchParameters.bestHits: global, never cleared (main.c:376)
    try:
        for s in (7, 15):
            ctx.scan_read(ranges[s], 24); ctx.scan_transform(pose0)
            ox, oy = oracle.read_scan(ranges[s], angles, rmin, 24)
            otx, oty = oracle.transform(ox, oy, pose0)
            n = ctx.local_map_extract(1.0)
            olx, oly = oracle.extract_local_map(otx, oty, omx, omy, 1.0)
            assert n == len(olx)
            guess = np.array([0.02, -0.03, 0.004], np.float32)
            res1, res2 = np.array([0.05, 0.05, 0.008727], np.float32), np.array([0.025, 0.025, 0.004363], np.float32)
            oms = []
            for m, pix, cap in ((gm[0], 0.2, 200), (gm[1], 0.1, 400)):
                rows, cols, tl = m.rasterise_local(pix)
                ogrid, otl = oracle.occupational_grid(olx, oly, pix, cap, cap)
                assert (rows, cols) == ogrid.shape and np.array_equal(m.download_occupancy(), ogrid)
                m.edt()
                oms.append(oracle.make_map(oracle.edt(ogrid), np.float32(pix), otl))
            # FastMatch on the coarse grid, FastMatch2 on the fine one seeded by it (main.c:902-918)
            p1, _, _ = ctx.fastmatch(gm[0], guess, res1)
            p2, hits, nbest = ctx.fastmatch(gm[1], p1, res2)
            o1 = oracle.fastmatch(oms[0], ox, oy, guess, res1, hits_buf=obest)
            o2 = oracle.fastmatch(oms[1], ox, oy, o1[0], res2, hits_buf=obest)
            assert np.array_equal(bits(p2), bits(o2[0])) and nbest == o2[2]
            # growth at the matched pose (main.c:936-948)
            ctx.scan_transform(p2)
            added = ctx.mappoints_grow(1.5)
            gtx, gty = oracle.transform(ox, oy, p2)
            omx2, omy2 = oracle.grow_map(o2[1], o2[2], gtx, gty, omx, omy)
            assert added == len(omx2) - len(omx)
            mx, my = ctx.mappoints_download()
            assert _same(mx, omx2) and _same(my, omy2)
            omx, omy = omx2, omy2
    finally:
        for m in gm:
            m.close()
