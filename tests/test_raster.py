"""OccupationalGrid rasterisation (Subsystem_1/main.c:271-354).

CPU: the oracle restatement against the reference's own OccupationalGrid (both pixel sizes, grids,
sizes, corners, and the distance fields it produces).  GPU: b200slam_map_rasterise against the
oracle, then the whole device-resident chain points -> grid -> EDT -> FastMatch against the
reference flow."""
import numpy as np
import pytest

from conftest import bits


def _points(synth, n=6000, seed=3, extent=(14.0, 9.0)):
    """Wall-like map points: a noisy rectangle outline plus a few clusters."""
    u = synth.hash_uniform(seed, np.arange(4 * n)).reshape(4, n)
    side = (u[0] * 4).astype(int)
    t = u[1]
    hw, hh = extent[0] / 2, extent[1] / 2
    x = np.where(side == 0, -hw, np.where(side == 1, hw, (2 * t - 1) * hw))
    y = np.where(side == 2, -hh, np.where(side == 3, hh, (2 * t - 1) * hh))
    x = x + (u[2] - 0.5) * 0.06 + 1.37
    y = y + (u[3] - 0.5) * 0.06 - 0.61
    return x.astype(np.float32), y.astype(np.float32)


@pytest.mark.parametrize("seed,extent", [(3, (14.0, 9.0)), (4, (30.0, 18.5)), (5, (2.0, 35.0))])
def test_oracle_matches_reference_occupational_grid(oracle, synth, seed, extent):
    from oracle import pyoracle
    if not pyoracle.reference_available():
        pytest.skip("oracle/_ref not built")
    ref = pyoracle.Reference("accel")
    x, y = _points(synth, 5000 + 7 * seed, seed, extent)
    levels = ref.occupational_grid(x, y, 0.2, 0.1)
    for (rgrid, rfield, rtl), (pix, cap) in zip(levels, ((0.2, 200), (0.1, 400))):
        ogrid, otl = oracle.occupational_grid(x, y, pix, cap, cap)
        assert ogrid.shape == rgrid.shape and np.array_equal(ogrid, rgrid)
        assert bits(np.array(otl)).tolist() == bits(np.array(rtl)).tolist()
        assert np.array_equal(bits(oracle.edt(ogrid)), bits(rfield))      # and the EDT of it


def test_oracle_grid_too_large_is_reported(oracle, synth):
    x, y = _points(synth, 100, 9, (60.0, 5.0))
    with pytest.raises(ValueError):
        oracle.occupational_grid(x, y, 0.1, 400, 400)


@pytest.mark.gpu
@pytest.mark.parametrize("pix,cap", [(0.2, 200), (0.1, 400), (0.05, 1024)])
def test_rasterise_matches_oracle(ctx, oracle, synth, pix, cap):
    for seed, extent in ((3, (14.0, 9.0)), (4, (30.0, 18.5)), (6, (7.3, 31.0))):
        x, y = _points(synth, 4000 + seed, seed, extent)
        try:
            ogrid, otl = oracle.occupational_grid(x, y, pix, cap, cap)
        except ValueError:
            m = ctx.new_map(cap, cap)
            with pytest.raises(Exception):
                m.rasterise(x, y, pix)
            m.close()
            continue
        m = ctx.new_map(cap, cap)
        try:
            rows, cols, tl = m.rasterise(x, y, pix)
            assert (rows, cols) == ogrid.shape
            assert bits(np.array(tl)).tolist() == bits(np.array(otl)).tolist()
            assert np.array_equal(m.download_occupancy(), ogrid)
            m.edt()
            assert np.array_equal(bits(m.download_field()), bits(oracle.edt(ogrid)))
            # a second, smaller rasterisation into the same map must not see stale cells
            x2, y2 = _points(synth, 500, seed + 20, (extent[0] * 0.5, extent[1] * 0.5))
            ogrid2, _ = oracle.occupational_grid(x2, y2, pix, cap, cap)
            m.rasterise(x2, y2, pix)
            assert np.array_equal(m.download_occupancy(), ogrid2)
        finally:
            m.close()


@pytest.mark.gpu
def test_points_to_pose_chain_matches_reference_flow(ctx, oracle, synth):
    """points -> grid -> EDT -> FastMatch entirely on the device == the oracle's chain."""
    x, y = _points(synth, 8000, 11, (16.0, 10.0))
    sx, sy = synth.scan_random(1079, seed=5, rmin=1.0, rspan=6.0)
    pose = np.array([1.2, -0.5, 0.03], np.float32)
    for pix, cap, res3 in ((0.2, 200, (0.05, 0.05, 0.008727)), (0.1, 400, (0.025, 0.025, 0.004363))):
        m = ctx.new_map(cap, cap)
        try:
            m.rasterise(x, y, pix)
            m.edt()
            ctx.scan_upload(sx, sy)
            gpose, ghits, gn = ctx.fastmatch(m, pose, np.array(res3, np.float32))
            ogrid, otl = oracle.occupational_grid(x, y, pix, cap, cap)
            om = oracle.make_map(oracle.edt(ogrid), pix, otl)
            opose, ohits, on = oracle.fastmatch(om, sx, sy, pose, np.array(res3, np.float32))
            assert gn == on and np.array_equal(bits(gpose), bits(opose))
        finally:
            m.close()


@pytest.mark.gpu
def test_rasterise_sequence_erases_exactly_the_previous_cells(ctx, oracle, synth):
    """The map remembers the cells a rasterisation set and the next one erases those instead of clearing the
    grid (csrc/raster.cu).  A sequence on ONE map -- growing and shrinking point sets (the cell list is
    reallocated), both pixel sizes, a host upload of a full occupancy in between, pageable and page-locked
    (b200slam_map_rasterise_async) points -- must leave exactly the oracle's grid every time."""
    cap = 400
    m = ctx.new_map(cap, cap)
    try:
        steps = [(300, 31, (9.0, 6.0), 0.1, False), (9000, 32, (30.0, 18.5), 0.1, True), (40, 33, (3.0, 2.0), 0.2, True),
                 (6000, 34, (14.0, 30.0), 0.1, False), (6000, 35, (14.0, 9.0), 0.2, True)]
        for k, (n, seed, extent, pix, pinned) in enumerate(steps):
            x, y = _points(synth, n, seed, extent)
            ogrid, otl = oracle.occupational_grid(x, y, pix, cap, cap)
            if k == 3:                                   # someone uploads a dense grid: the cell list is void
                m.rows, m.cols = cap, cap
                ctx._check(ctx.L.b200slam_map_resize(m.h, cap, cap))
                m.upload_occupancy(np.ones((cap, cap), np.int32))
            if pinned:
                px = ctx.pinned_empty((n,), np.float32); px[...] = x
                py = ctx.pinned_empty((n,), np.float32); py[...] = y
                rows, cols, tl = m.rasterise_async(px, py, pix)
            else:
                rows, cols, tl = m.rasterise(x, y, pix)
            assert (rows, cols) == ogrid.shape
            assert bits(np.array(tl)).tolist() == bits(np.array(otl)).tolist()
            assert np.array_equal(m.download_occupancy(), ogrid), f"step {k}"
            # nothing stale anywhere in the allocation either: the whole capacity, not just the region in use
            ctx._check(ctx.L.b200slam_map_resize(m.h, cap, cap))
            m.rows, m.cols = cap, cap
            full = m.download_occupancy()
            assert int(full.sum()) == int(ogrid.sum()), f"step {k}: stale cells outside the grid in use"
            ctx._check(ctx.L.b200slam_map_resize(m.h, rows, cols))
            m.rows, m.cols = rows, cols
        with pytest.raises(Exception):                   # pageable arrays are refused by the async form
            m.rasterise_async(np.zeros(8, np.float32), np.zeros(8, np.float32), 0.1)
    finally:
        m.close()
