"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/*.so, built
from /root/reference by oracle/Makefile) on seeded inputs.  Run in the build container:

    make -C oracle && python tests/golden/make_golden.py

The fixtures pin both the CPU oracle (tests/test_oracle_vs_reference.py, no GPU) and the
CUDA path (tests/test_*_gpu.py) to the reference's own outputs on machines where
/root/reference does not exist.

  edt_golden.npz       reference euclidean_distance_transform{,2} on 8 grids
                       (Subsystem_1/main_accelerated.c:215-283; cross-checked here against
                        Subsystem_1/main.c:223-269 and the Submodule_2 fragment)
  fastmatch_golden.npz reference FastMatch / FastMatch2 (Subsystem_1/main.c:381-809):
                       inputs (field, geometry, scan, pose, resolution) and outputs
                       (FastMatchParameters.pose, bestHits_size, bestHits[0:last_hits])
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle, Reference  # noqa: E402

synth = importlib.import_module("hardware-acceleration-of-lidar-slam_b200.synth")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    accel, main_ref, frag = Reference("accel"), Reference("main"), Reference("edtfrag")
    orc = Oracle()

    # ---- EDT ---------------------------------------------------------------------------
    edt = {}
    cases = [  # (rows, cols, fine, kind, param)
        (37, 53, False, "bernoulli", 0.02), (200, 200, False, "bernoulli", 0.01),
        (64, 64, False, "empty", 0.0), (50, 70, False, "full", 1.0),
        (123, 87, True, "bernoulli", 0.05), (157, 127, True, "rooms", 0.0),
        (1, 40, False, "bernoulli", 0.1), (33, 1, True, "bernoulli", 0.2),
    ]
    for k, (rows, cols, fine, kind, p) in enumerate(cases):
        if kind == "bernoulli":
            occ = synth.grid_bernoulli(rows, cols, p, seed=0xED70000 + k)
        elif kind == "empty":
            occ = np.zeros((rows, cols), np.int32)
        elif kind == "full":
            occ = np.ones((rows, cols), np.int32)
        else:
            occ = synth.grid_rooms(rows, cols, seed=0xED70000 + k)
        out = accel.edt(occ, fine)
        assert np.array_equal(out.view(np.uint32), main_ref.edt(occ, fine).view(np.uint32))
        if rows == cols:
            assert np.array_equal(out.view(np.uint32), frag.edt(occ, fine).view(np.uint32))
        edt[f"occ_{k}"] = occ.astype(np.uint8)
        edt[f"out_{k}"] = out
        edt[f"fine_{k}"] = np.array(fine)
    edt["count"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "edt_golden.npz"), **edt)

    # ---- FastMatch ---------------------------------------------------------------------
    fm = {}
    n = 0
    for k in range(10):
        fine = bool(k % 2)
        rows, cols = (110 + 9 * k, 190 - 7 * k) if not fine else (260 + 11 * k, 300 - 9 * k)
        px = 0.1 if fine else 0.2
        occ = synth.grid_rooms(rows, cols, seed=0xFA570000 + k, n_segments=10, n_pillars=8)
        field = accel.edt(occ, fine)
        pixel, tl = synth.centred_geometry(rows, cols, px)
        true_pose = (0.3 * np.cos(k), 0.25 * np.sin(2 * k), 0.4 * k - 1.5)
        if k < 8:
            sx, sy = synth.scan_raycast(occ, px, tl, true_pose, 1079, max_range=24.0, reference_lidar=True,
                                        noise_seed=0x5CA70000 + k)
        else:   # all beams outside the grid: the zero-hit quirk (score 0 wins)
            sx = np.full(50, 500.0, np.float32)
            sy = np.full(50, -500.0, np.float32)
        pose = np.array([true_pose[0] + 0.03, true_pose[1] - 0.04, true_pose[2] + 0.006], np.float32)
        res = np.array([0.05, 0.05, 0.008727], np.float32) if k % 4 < 2 else np.array([0.025, 0.025, 0.004363], np.float32)
        for ref in (accel, main_ref):
            ref.set_map(field, float(pixel), tl, fine)
            ref.set_scan(sx, sy)
        p1, h1, n1 = accel.fastmatch(pose, res, fine)
        p2, h2, n2 = main_ref.fastmatch(pose, res, fine)
        assert np.array_equal(p1.view(np.uint32), p2.view(np.uint32)) and n1 == n2
        om = orc.make_map(field, float(pixel), tl)
        m, _, _ = orc.score_lattice(om, sx, sy, pose, [res[0], res[0], res[2]], [3, 3, 3])
        last = m.last_hits
        assert np.array_equal(h1[:last], h2[:last])
        fm[f"field_{n}"] = field
        fm[f"geom_{n}"] = np.array([pixel, tl[0], tl[1]], np.float32)
        fm[f"scan_x_{n}"], fm[f"scan_y_{n}"] = sx, sy
        fm[f"pose_{n}"], fm[f"res_{n}"] = pose, res
        fm[f"out_pose_{n}"] = p1
        fm[f"out_size_{n}"] = np.array(n1, np.int32)
        fm[f"out_last_{n}"] = np.array(last, np.int32)
        fm[f"out_hits_{n}"] = h1[:last].copy()
        n += 1
    fm["count"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "fastmatch_golden.npz"), **fm)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
