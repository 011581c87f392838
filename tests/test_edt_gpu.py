"""GPU parity: CUDA distance transform (through the C ABI) vs the CPU oracle and the
reference's golden vectors.  Bar: bit-exact (integer d2min -> IEEE sqrtf)."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


def test_edt_golden_vectors(ctx, edt_golden):
    g = edt_golden
    for k in range(int(g["count"])):
        occ = g[f"occ_{k}"].astype(np.int32)
        out = ctx.edt(occ)
        assert np.array_equal(bits(out), bits(g[f"out_{k}"])), f"golden EDT case {k}"


@pytest.mark.parametrize("rows,cols,p", [
    (1, 1, 1.0), (1, 1, 0.0), (1, 97, 0.1), (83, 1, 0.1), (7, 9, 0.3), (31, 33, 0.05), (64, 64, 0.01),
    (65, 63, 0.01), (200, 200, 0.01), (400, 400, 0.05), (157, 127, 0.02), (300, 1000, 0.001),
    (1000, 300, 0.003), (513, 767, 0.0), (90, 2111, 0.02),
])
def test_edt_matches_oracle_bernoulli(ctx, oracle, synth, rows, cols, p):
    occ = synth.grid_bernoulli(rows, cols, p, seed=rows * 7919 + cols)
    assert np.array_equal(bits(ctx.edt(occ)), bits(oracle.edt(occ)))


def test_edt_rooms_and_nonbinary_occupancy(ctx, oracle, synth):
    occ = synth.grid_rooms(700, 900, seed=4)
    assert np.array_equal(bits(ctx.edt(occ)), bits(oracle.edt(occ)))
    # any non-zero cell is an obstacle (Subsystem_1/main.c:227 tests truthiness)
    occ2 = occ * np.int32(-7) + (occ * synth.grid_bernoulli(700, 900, 0.5, seed=9)) * np.int32(1 << 20)
    assert np.array_equal(bits(ctx.edt(occ2)), bits(oracle.edt((occ2 != 0).astype(np.int32))))


def test_edt_reference_strides_and_stale_border(ctx, oracle, synth):
    # the reference keeps a fixed 400-wide array and only rows x cols is written
    # (main.c:225-226); everything outside must be left untouched
    rows, cols = 157, 127
    inp = np.zeros((400, 400), np.int32)
    inp[:rows, :cols] = synth.grid_bernoulli(rows, cols, 0.02, seed=31)
    inp[rows:, :] = 1          # stale junk outside the sub-rectangle must not influence it
    inp[:, cols:] = 1
    out = np.full((400, 400), np.float32(-3.0))
    ctx.edt(inp[:rows, :cols], out=out[:rows, :cols])
    assert np.array_equal(bits(out[:rows, :cols]), bits(oracle.edt(np.ascontiguousarray(inp[:rows, :cols]))))
    assert np.all(out[rows:, :] == -3.0) and np.all(out[:, cols:] == -3.0)


@pytest.mark.parametrize("max_dist", [0.5, 1.0, 1.5, 2.0, 3.0, 4.5, 7.0, 10.0, 12.0, 15.0, 16.0, 20.0])
def test_edt_other_max_dist(ctx, oracle, synth, max_dist):
    occ = synth.grid_bernoulli(140, 190, 0.004, seed=int(max_dist * 10))
    assert np.array_equal(bits(ctx.edt(occ, max_dist)), bits(oracle.edt(occ, max_dist)))


def test_edt_device_resident_map(ctx, oracle, synth):
    occ = synth.grid_rooms(333, 450, seed=8)
    m = ctx.new_map(333, 450)
    try:
        m.upload_occupancy(occ).edt()
        assert np.array_equal(bits(m.download_field()), bits(oracle.edt(occ)))
        # idempotent re-run on resident data
        m.edt()
        assert np.array_equal(bits(m.download_field()), bits(oracle.edt(occ)))
    finally:
        m.close()


def test_edt_2048_full_size(ctx, oracle, synth):
    # BASELINE configs[1] grid size, compared cell for cell with the oracle
    occ = synth.grid_bernoulli(2048, 2048, 0.01)
    out = ctx.edt(occ)
    assert np.array_equal(bits(out), bits(oracle.edt(occ)))


def test_edt_8192_properties_and_tiles(ctx, oracle, synth):
    # BASELINE configs[3] grid size: size-independent properties + oracle on sampled tiles
    n = 8192
    occ = synth.grid_bernoulli(n, n, 0.002)
    out = ctx.edt(occ)
    assert np.all(out[occ != 0] == 0.0)
    assert float(out.max()) == 10.0 and float(out.min()) == 0.0
    allowed = np.unique(np.concatenate([np.sqrt(np.arange(100, dtype=np.float32)), [np.float32(10.0)]]))
    assert np.all(np.isin(np.unique(out), allowed))
    # an interior tile depends only on the tile plus a 9-cell halo
    for (r0, c0) in [(0, 0), (4000, 4100), (8192 - 300, 8192 - 300), (0, 8192 - 280), (5000, 0)]:
        r1, c1 = min(r0 + 300, n), min(c0 + 300, n)
        rr0, cc0, rr1, cc1 = max(r0 - 9, 0), max(c0 - 9, 0), min(r1 + 9, n), min(c1 + 9, n)
        ref = oracle.edt(np.ascontiguousarray(occ[rr0:rr1, cc0:cc1]))
        want = ref[r0 - rr0:r0 - rr0 + (r1 - r0), c0 - cc0:c0 - cc0 + (c1 - c0)]
        assert np.array_equal(bits(out[r0:r1, c0:c1]), bits(want)), (r0, c0)
    # monotone under adding obstacles
    occ2 = occ.copy()
    occ2[::97, ::89] = 1
    out2 = ctx.edt(occ2)
    assert np.all(out2 <= out)


def test_edt_empty_shapes(ctx):
    out = ctx.edt(np.zeros((0, 5), np.int32))
    assert out.shape == (0, 5)


@pytest.mark.gpu
def test_bad_arguments_are_reported_not_fatal(ctx, b200slam, synth):
    """Every entry point answers bad input with an error code and a message (the new
    generic API returns int status and never throws or crashes)."""
    L = ctx.L
    occ = np.zeros((8, 8), np.int32)
    out = np.zeros((8, 8), np.float32)
    assert L.b200slam_edt(ctx.h, None, 8, out.ctypes.data, 8, 8, 8, 10.0) == b200slam.ERR_ARG
    assert L.b200slam_edt(ctx.h, occ.ctypes.data, 4, out.ctypes.data, 8, 8, 8, 10.0) == b200slam.ERR_ARG   # stride < cols
    assert L.b200slam_edt(ctx.h, occ.ctypes.data, 8, out.ctypes.data, 8, 8, 8, -1.0) == b200slam.ERR_ARG
    assert L.b200slam_edt(ctx.h, occ.ctypes.data, 8, out.ctypes.data, 8, 8, 8, 300.0) == b200slam.ERR_ARG
    assert b"max_dist" in L.b200slam_last_error(ctx.h)
    with pytest.raises(b200slam.B200SlamError):
        ctx.new_map(0, 5)
    m = ctx.new_map(16, 16)
    try:
        with pytest.raises(b200slam.B200SlamError) as e:            # geometry not set
            ctx.score_lattice(m, (0, 0, 0), (0.05, 0.05, 0.01), (3, 3, 3))
        assert e.value.code == b200slam.ERR_STATE
        m.set_geometry(0.1, (0.0, 0.0))
        with pytest.raises(b200slam.B200SlamError):                 # zero-sized lattice
            ctx.score_lattice(m, (0, 0, 0), (0.05, 0.05, 0.01), (0, 3, 3))
        with pytest.raises(b200slam.B200SlamError):                 # row range outside the lattice
            ctx.score_lattice_rows(m, (0, 0, 0), (0.05, 0.05, 0.01), (3, 3, 3), 0, 10)
        assert L.b200slam_map_resize(m.h, 17, 4) == b200slam.ERR_ARG
        x = np.array([0.0, 100.0], np.float32)
        with pytest.raises(b200slam.B200SlamError):                 # rasterised grid exceeds the capacity
            m.rasterise(x, x, 0.1)
        # the context is still usable afterwards
        assert np.array_equal(ctx.edt(occ), np.full((8, 8), 10.0, np.float32))
    finally:
        m.close()


def test_edt_row_ranges_compose_and_touch_nothing_else(ctx, oracle, synth):
    """b200slam_map_edt_rows (one rank's block of the row-sharded transform): ranges that
    are not multiples of the 9-row batch, a one-row and an empty range; rows outside a range keep the
    sentinel; the union equals the oracle bit for bit (the halo comes from the full occupancy)."""
    rows, cols = 1000, 700
    occ = synth.grid_bernoulli(rows, cols, 0.01, synth.SEED_GRID + 77)
    want = oracle.edt(occ)
    m = ctx.new_map(rows, cols)
    try:
        m.upload_occupancy(occ)
        m.upload_field(np.full((rows, cols), -1.0, np.float32))
        m.edt_rows(0, 0)
        m.edt_rows(331, 332)
        got = m.download_field()
        assert np.array_equal(bits(got[331]), bits(want[331]))
        assert (np.delete(got, 331, axis=0) == -1.0).all()
        for rb, re in [(0, 331), (332, 640), (640, 993), (993, 1000)]:
            m.edt_rows(rb, re)
        assert np.array_equal(bits(m.download_field()), bits(want))
        with pytest.raises(Exception):
            m.edt_rows(10, rows + 1)
    finally:
        m.close()


@pytest.mark.parametrize("max_dist", [10.0, 2.0, 15.0, 3.5])
def test_edt_byte_shadow_and_int32_paths_agree(ctx, oracle, synth, monkeypatch, max_dist):
    """Maps of >= 2^20 cells keep a byte shadow of the occupancy (uploads pack it, rasterisations write both) and
    the transform streams that instead of the int32 grid (csrc/edt.cu, OT = uint8_t).  Both paths against the
    oracle on the same grids -- ragged size (partial strips and batches), non-binary cells, a sub-rectangle
    after b200slam_map_resize -- and the fall-back once the raw occupancy pointer has been handed out."""
    import ctypes as C
    rows, cols = 1111, 1203
    occ = synth.grid_rooms(rows, cols, seed=21) * np.int32(-5) + synth.grid_bernoulli(rows, cols, 0.002, seed=22) * np.int32(1 << 9)
    want = oracle.edt((occ != 0).astype(np.int32), max_dist)
    sub = (700, 1000)
    want_sub = oracle.edt((occ[:sub[0], :sub[1]] != 0).astype(np.int32), max_dist)
    for no_bytes in (False, True):
        if no_bytes:
            monkeypatch.setenv("B200SLAM_EDT_NO_BYTES", "1")        # read when the map is created
        m = ctx.new_map(rows, cols)
        try:
            m.upload_occupancy(occ).edt(max_dist)
            assert np.array_equal(bits(m.download_field()), bits(want)), ("int32" if no_bytes else "bytes")
            ctx._check(ctx.L.b200slam_map_resize(m.h, sub[0], sub[1]))
            m.rows, m.cols = sub
            m.edt(max_dist)                                          # same memory, smaller grid in use
            assert np.array_equal(bits(m.download_field()), bits(want_sub))
            if not no_bytes:
                # the raw pointer leaves the library: the shadow can no longer be trusted, the int32 grid is read
                p, pitch = C.c_void_p(), C.c_int32(0)
                ctx._check(ctx.L.b200slam_map_device_ptrs(m.h, C.byref(p), C.byref(pitch), None, None))
                ctx._check(ctx.L.b200slam_map_resize(m.h, rows, cols))
                m.rows, m.cols = rows, cols
                m.upload_occupancy(occ).edt(max_dist)
                assert np.array_equal(bits(m.download_field()), bits(want))
        finally:
            monkeypatch.delenv("B200SLAM_EDT_NO_BYTES", raising=False)
            m.close()
