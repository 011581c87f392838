"""CPU: pins the oracle (oracle/slam_oracle.c) to the reference.

* against the committed golden vectors (tests/golden, generated from the unmodified
  reference by tests/golden/make_golden.py) -- always;
* against the reference's own compiled functions (oracle/_ref/*.so) on fresh random
  inputs -- whenever those objects exist (build container, or shipped to the GPU box).
"""
import numpy as np
import pytest

from conftest import bits
from oracle import pyoracle

needs_ref = pytest.mark.skipif(not pyoracle.reference_available(),
                               reason="oracle/_ref not built (needs /root/reference)")


def test_edt_oracle_matches_golden(oracle, edt_golden):
    g = edt_golden
    for k in range(int(g["count"])):
        occ = g[f"occ_{k}"].astype(np.int32)
        assert np.array_equal(bits(oracle.edt(occ)), bits(g[f"out_{k}"])), f"golden EDT case {k}"


def test_edt_variants_agree_small(oracle, synth):
    # the three reference formulations (per-cell search, scatter, separable closed form)
    for k, (rows, cols, p) in enumerate([(23, 31, 0.05), (40, 17, 0.2), (12, 12, 0.0), (9, 30, 1.0)]):
        occ = synth.grid_bernoulli(rows, cols, p, seed=100 + k)
        a = oracle.edt(occ)
        assert np.array_equal(bits(a), bits(oracle.edt(occ, variant="percell")))
        assert np.array_equal(bits(a), bits(oracle.edt(occ, variant="scatter")))


def test_edt_closed_form_properties(oracle, synth):
    occ = synth.grid_bernoulli(90, 110, 0.01, seed=5)
    d = oracle.edt(occ)
    assert np.all(d[occ != 0] == 0.0)
    assert d.max() <= 10.0 and d.min() >= 0.0
    # strict '<' against 10^2: a lone obstacle 10 cells away (d2 = 100, also 6-8-10) clamps
    occ = np.zeros((30, 30), np.int32)
    occ[5, 5] = 1
    d = oracle.edt(occ)
    assert d[5, 15] == 10.0 and d[13, 11] == 10.0 and d[5, 14] == 9.0
    assert d[12, 12] == np.sqrt(np.float32(98.0))


@pytest.mark.parametrize("max_dist", [1.0, 2.5, 4.0, 10.0, 13.0, 17.0])
def test_edt_other_max_dist(oracle, synth, max_dist):
    occ = synth.grid_bernoulli(40, 45, 0.01, seed=77)
    a = oracle.edt(occ, max_dist)
    b = oracle.edt(occ, max_dist, variant="percell")
    assert np.array_equal(bits(a), bits(b))


def test_fastmatch_oracle_matches_golden(oracle, fastmatch_golden):
    g = fastmatch_golden
    for k in range(int(g["count"])):
        pixel, tlx, tly = g[f"geom_{k}"]
        om = oracle.make_map(g[f"field_{k}"], pixel, (tlx, tly))
        pose, hits, n = oracle.fastmatch(om, g[f"scan_x_{k}"], g[f"scan_y_{k}"], g[f"pose_{k}"], g[f"res_{k}"])
        last = int(g[f"out_last_{k}"])
        assert np.array_equal(bits(pose), bits(g[f"out_pose_{k}"])), f"case {k} pose"
        assert n == int(g[f"out_size_{k}"]), f"case {k} bestHits_size"
        assert np.array_equal(bits(hits[:last]), bits(g[f"out_hits_{k}"])), f"case {k} bestHits"


def test_zero_hit_candidate_wins(oracle, fastmatch_golden):
    # a candidate whose beams all miss the grid scores 0 and wins
    g = fastmatch_golden
    k = int(g["count"]) - 1
    assert int(g[f"out_size_{k}"]) == 0
    assert np.array_equal(bits(g[f"out_pose_{k}"]),
                          bits(np.array([g[f"pose_{k}"][0] - g[f"res_{k}"][0], g[f"pose_{k}"][1] - g[f"res_{k}"][0],
                                         g[f"pose_{k}"][2] - g[f"res_{k}"][2]], np.float32)))


def test_lattice_equals_pose_list(oracle, synth):
    # the pose-list scorer restates the same per-beam arithmetic: a lattice expanded to an
    # explicit pose list (with the lattice's own cos/sin) gives identical scores
    w = synth.make_workload("tiny")
    field = oracle.edt(w["occ"])
    om = oracle.make_map(field, w["pixel"], w["top_left"])
    n = (5, 4, 6)
    res, scores, _ = oracle.score_lattice(om, w["scan_x"], w["scan_y"], w["pose0"], w["step"], n)
    L = oracle.lib
    poses = np.array([[L.orc_lattice_value(w["pose0"][0], w["step"][0], ix, n[1]),
                       L.orc_lattice_value(w["pose0"][1], w["step"][1], iy, n[2]),
                       L.orc_lattice_value(w["pose0"][2], w["step"][2], it, n[0])]
                      for it in range(n[0]) for ix in range(n[1]) for iy in range(n[2])], np.float32)
    res2, scores2, hits2 = oracle.score_poses(om, w["scan_x"], w["scan_y"], poses)
    assert np.array_equal(bits(scores), bits(scores2))
    assert res.best_index == res2.best_index and res.best_hits == res2.best_hits


def test_weights_resample_definition(oracle):
    rng = np.random.default_rng(3)
    scores = (rng.random(5000) * 40 + 100).astype(np.float32)
    w, q, W, anc = oracle.weights_resample(scores, 0.5, 0x80000000)
    assert W == int(q.sum()) and q.max() == 2 ** 32            # best particle has weight exactly 1
    assert abs(float(w.sum()) - 1.0) < 1e-4
    assert np.all(np.diff(anc) >= 0) and anc.min() >= 0 and anc.max() < len(scores)
    # offspring counts follow the weights to within one
    counts = np.bincount(anc, minlength=len(scores))
    expect = q.astype(np.float64) / W * len(scores)
    assert np.all(np.abs(counts - expect) <= 1.0 + 1e-9)
    # exp_det is a faithful exp
    xs = -np.linspace(0, 30, 301)
    got = np.array([oracle.exp_det(float(x)) for x in xs])
    assert np.allclose(got, np.exp(xs.astype(np.float32).astype(np.float64)), rtol=2e-7, atol=0)


# ---- live against the compiled reference -------------------------------------------
@needs_ref
@pytest.mark.parametrize("which", ["accel", "main", "edtfrag"])
def test_edt_oracle_vs_compiled_reference(oracle, synth, which):
    ref = pyoracle.Reference(which)
    rng = np.random.default_rng(11)
    for k in range(6):
        fine = bool(k % 2)
        S = 400 if fine else 200
        rows = int(rng.integers(1, 90))
        cols = rows if which == "edtfrag" else int(rng.integers(1, 90))
        assert rows <= S
        occ = synth.grid_bernoulli(rows, cols, float(rng.choice([0.0, 0.003, 0.02, 0.2])), seed=900 + k)
        assert np.array_equal(bits(ref.edt(occ, fine)), bits(oracle.edt(occ))), (which, k, rows, cols)


@needs_ref
@pytest.mark.parametrize("which", ["accel", "main"])
def test_fastmatch_oracle_vs_compiled_reference(oracle, synth, which):
    ref = pyoracle.Reference(which)
    rng = np.random.default_rng(21)
    for k in range(8):
        fine = bool(k % 2)
        S = 400 if fine else 200
        rows, cols = int(rng.integers(60, S)), int(rng.integers(60, S))
        px = 0.1 if fine else 0.2
        occ = synth.grid_rooms(rows, cols, seed=300 + k, n_segments=8, n_pillars=6)
        field = oracle.edt(occ)
        pixel, tl = synth.centred_geometry(rows, cols, px)
        pose_true = (rng.normal() * 0.3, rng.normal() * 0.3, rng.uniform(-3, 3))
        sx, sy = synth.scan_raycast(occ, px, tl, pose_true, 1079, reference_lidar=True, noise_seed=k)
        if len(sx) == 0:
            continue
        pose = np.array(pose_true, np.float32) + np.array([0.04, -0.03, 0.005], np.float32)
        res = np.array([0.05, 0.05, 0.008727], np.float32)
        ref.set_map(field, float(pixel), tl, fine)
        ref.set_scan(sx, sy)
        p1, h1, n1 = ref.fastmatch(pose, res, fine)
        om = oracle.make_map(field, pixel, tl)
        p2, h2, n2 = oracle.fastmatch(om, sx, sy, pose, res)
        m, _, _ = oracle.score_lattice(om, sx, sy, pose, [res[0], res[0], res[2]], [3, 3, 3])
        assert np.array_equal(bits(p1), bits(p2)) and n1 == n2
        assert np.array_equal(bits(h1[:m.last_hits]), bits(h2[:m.last_hits]))


@pytest.mark.parametrize("which", ["accel", "main"])
def test_besthits_global_is_left_exactly_as_the_reference_leaves_it(oracle, synth, which):
    """FastMatchParameters.bestHits is a global that EVERY candidate overwrites from index 0 (main.c:515) and
    nothing ever clears; main() reads it up to the WINNER's count (main.c:942-948), i.e. possibly past the last
    candidate's hits.  The oracle (and the CUDA path, tests/test_score_gpu.py) must leave the whole array as the
    reference does -- checked here over a sequence of calls against the compiled reference's own global,
    with poses walking off the grid so that the 27 candidates' hit counts differ."""
    from oracle import pyoracle
    if not pyoracle.reference_available():
        pytest.skip("oracle/_ref not built")
    ref = pyoracle.Reference(which)
    rows, cols = 150, 190
    pixel, tl = synth.centred_geometry(rows, cols, 0.2)
    # two vertical walls, one of them two cells from the right edge, and a scan that lies exactly on them
    # when seen from `true`: the aligned candidate scores 0 with every beam in bounds and wins, while the
    # last candidate (+t in x and y, main.c:422-438) pushes the outer wall's beams off the grid -- so the
    # winner has MORE hits than the last candidate and main.c:942-948 reads the array's tail
    occ = np.zeros((rows, cols), np.int32)
    wall_rows = np.arange(20, 131)
    occ[wall_rows, cols - 3] = 1
    occ[wall_rows, 60] = 1
    true = np.array([1.0, -0.6, 0.0], np.float32)
    wx = np.concatenate([np.full(len(wall_rows), float(tl[0]) + (cols - 3) * 0.2), np.full(len(wall_rows), float(tl[0]) + 60 * 0.2)])
    wy = np.concatenate([float(tl[1]) + wall_rows * 0.2] * 2)
    sx, sy = (wx - true[0]).astype(np.float32), (wy - true[1]).astype(np.float32)
    field = ref.edt(occ, fine=False)
    ref.set_map(field, float(pixel), tl, fine=False)
    ref.set_scan(sx, sy)
    om = oracle.make_map(field, float(pixel), tl)
    # the global lives as long as the loaded library: start from whatever earlier tests of this process left in it
    obuf = np.ctypeslib.as_array(ref.fmp.bestHits).astype(np.float32).copy()
    res = np.array([0.3, 0.3, 0.05], np.float32)
    saw_tail = False
    for k in range(7):
        pose = np.array([true[0] + 0.37 * (k % 3) - 0.2 * (k // 3), true[1] + 0.29 * (k // 2), 0.013 * k], np.float32)
        rp, rhits, rn = ref.fastmatch(pose, res, fine=False)          # rhits: the reference's whole global array
        op, _, on = oracle.fastmatch(om, sx, sy, pose, res, hits_buf=obuf)
        assert np.array_equal(bits(rp), bits(op)) and rn == on
        assert np.array_equal(bits(rhits), bits(obuf)), f"call {k}: bestHits differs at {np.flatnonzero(rhits != obuf)[:5]}"
        lat, _, _ = oracle.score_lattice(om, sx, sy, pose, [res[0], res[0], res[2]], (3, 3, 3))
        saw_tail |= lat.best_hits > lat.last_hits
    assert saw_tail, "the sequence never had a winner with more hits than the last candidate"


def test_clamped_edt_of_a_row_block_needs_only_a_nine_row_halo(oracle, synth):
    """The claim behind the row-sharded transform: with max_dist = 10 a
    block of output rows depends on the occupancy of those rows plus 9 above and below, nothing else."""
    rows, cols = 300, 170
    occ = synth.grid_bernoulli(rows, cols, 0.004, seed=0x5A4D)
    full = oracle.edt(occ)
    for rb, re in [(0, 97), (97, 98), (98, 211), (211, 300)]:
        lo, hi = max(0, rb - 9), min(rows, re + 9)
        part = oracle.edt(np.ascontiguousarray(occ[lo:hi]))
        assert np.array_equal(bits(part[rb - lo:re - lo]), bits(full[rb:re]))
        if rb >= 9 and re + 9 <= rows:                 # ... and 8 rows would not do
            lo8, hi8 = rb - 8, re + 8
            part8 = oracle.edt(np.ascontiguousarray(occ[lo8:hi8]))
            if not np.array_equal(bits(part8[rb - lo8:re - lo8]), bits(full[rb:re])):
                break
    else:
        pytest.skip("no block in this grid was sensitive to the 9th halo row")
