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


@pytest.mark.gpu
def test_fastmatch_pair_equals_two_fastmatch_calls(b200slam, oracle, synth, small_lattice_kernel):
    """b200slam_fastmatch_pair_async: FastMatch then FastMatch2 from its result (main.c:902-918) as two kernels
    with no host step between them -- the second picks its lattice by the first one's winner on the device --
    and the scan's size kept on the device (b200slam_scan_read_async).  Poses, bestHits_size and the bestHits[]
    twin must equal two synchronous b200slam_fastmatch calls and the oracle, also where the 27 candidates'
    hit counts differ (poses at the grid's edge) and with beams that readAScan drops."""
    w = synth.make_workload("tiny")
    field = oracle.edt(w["occ"])
    coarse_occ = w["occ"].reshape(w["occ"].shape[0] // 2, 2, w["occ"].shape[1] // 2, 2).max(axis=(1, 3)).astype(np.int32)
    cfield = oracle.edt(coarse_occ)
    pixel_c, tl_c = synth.centred_geometry(coarse_occ.shape[0], coarse_occ.shape[1], 0.2)
    angles = oracle.lidar_angles()
    rng = np.random.default_rng(11)
    ranges = rng.uniform(0.5, 9.0, len(angles)).astype(np.float32)
    ranges[rng.choice(len(angles), 90, replace=False)] = 30.0           # dropped by readAScan (main.c:78)
    ranges[5] = 0.01
    sx, sy = oracle.read_scan(ranges, angles)
    om_f = oracle.make_map(field, w["pixel"], w["top_left"])
    om_c = oracle.make_map(cfield, pixel_c, tl_c)
    res_a = np.array([0.05, 0.05, 0.008727], np.float32)
    res_b = np.array([0.025, 0.025, 0.004363], np.float32)
    poses = [np.array([w["pose0"][0] + 1.7 * k, w["pose0"][1] + 1.1 * k, w["pose0"][2] + 0.3 * k], np.float32) for k in range(9)]
    with b200slam.Context(0) as ca, b200slam.Context(0) as cb, b200slam.Context(0) as cc:
        maps = {}
        for c in (ca, cb, cc):
            mf = c.new_map(*field.shape); mf.set_geometry(w["pixel"], w["top_left"]).upload_field(field)
            mc = c.new_map(*cfield.shape); mc.set_geometry(pixel_c, tl_c).upload_field(cfield)
            maps[c] = (mc, mf)
            c.lidar_set(angles, 0.023)
        obuf = np.zeros(2500, np.float32)
        sbuf = np.zeros(2500, np.float32)
        for k, p in enumerate(poses):
            first_c = k % 2 == 0                                       # coarse grid after a rebuild, else the fine one
            # oracle
            o1, _, _ = oracle.fastmatch(om_c if first_c else om_f, sx, sy, p, res_a, hits_buf=obuf)
            o2, _, on = oracle.fastmatch(om_f, sx, sy, o1, res_b, hits_buf=obuf)
            # two synchronous calls, host-known scan size
            n = ca.scan_read(ranges)
            assert n == len(sx)
            s1, _, _ = ca.fastmatch(maps[ca][0] if first_c else maps[ca][1], p, res_a, hits_buf=sbuf)
            s2, _, sn = ca.fastmatch(maps[ca][1], s1, res_b, hits_buf=sbuf)
            # the pair, scan size on the device
            cb.scan_read_async(ranges)
            cb.fastmatch_pair_async(maps[cb][0] if first_c else maps[cb][1], maps[cb][1], p, res_a, res_b)
            pa, pb, pn, pbh = cb.fastmatch_pair_fetch()
            assert pn == n
            assert np.array_equal(bits(pa), bits(o1)) and np.array_equal(bits(pb), bits(o2)) and pbh == on, k
            assert np.array_equal(bits(s1), bits(o1)) and np.array_equal(bits(s2), bits(o2)) and sn == on, k
            assert np.array_equal(bits(cb.match_fetch_hits(2500)), bits(obuf)), k
            assert np.array_equal(bits(sbuf), bits(obuf)), k
            # readAScan + both matches as ONE kernel (b200slam_scan_step_async): same poses, counts, bestHits[] twin,
            # and the scan it leaves on the device is readAScan's
            cc.scan_step_async(ranges, maps[cc][0] if first_c else maps[cc][1], maps[cc][1], p, res_a, res_b)
            qa, qb, qn, qbh = cc.fastmatch_pair_fetch()
            assert qn == n and qbh == on
            assert np.array_equal(bits(qa), bits(o1)) and np.array_equal(bits(qb), bits(o2)), k
            assert np.array_equal(bits(cc.match_fetch_hits(2500)), bits(obuf)), k
            gx, gy = cc.scan_download(transformed=False)
            assert np.array_equal(bits(gx), bits(sx)) and np.array_equal(bits(gy), bits(sy))
        for c in (ca, cb, cc):
            for m in maps[c]:
                m.close()
