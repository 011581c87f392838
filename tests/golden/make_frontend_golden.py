"""Generates tests/golden/frontend_golden.npz by running the UNMODIFIED reference's own
SetLidarParameters / readAScan / Transform / ExtractLocalMap (Subsystem_1/main_accelerated.c, same
text as Subsystem_1/main.c:45-198) on its own globals, on seeded inputs:

    make -C oracle && python tests/golden/make_frontend_golden.py

Pins oracle.read_scan / transform / extract_local_map and the device front end
(b200slam_scan_read / scan_transform / local_map_extract) where /root/reference does not exist."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Reference  # noqa: E402

synth = importlib.import_module("hardware-acceleration-of-lidar-slam_b200.synth")
HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    """(ranges[1079], pose, map points, border): synthetic scans of the replay room, plus ranges on
    and around the validity limits (0.023 m, 24 m), and map points on the box edges."""
    ranges_all = synth.lidar_dataset(40, seed=synth.SEED_SCAN + 5)
    out = []
    for k in range(6):
        r = ranges_all[5 * k].astype(np.float32).copy()
        if k == 1:                                   # limits: exactly min / max, just inside / outside
            r[:8] = [0.023, np.nextafter(np.float32(0.023), np.float32(0)), 24.0, np.nextafter(np.float32(24), np.float32(25)),
                     0.0, 60.0, 23.999, 0.0231]
        if k == 2:
            r[::3] = 30.0                            # a third of the beams out of range
        if k == 3:
            r[1:] = 100.0                            # a single survivor
        pose = np.array([0.4 * k - 1.0, 0.3 - 0.2 * k, 0.37 * k - 0.9], np.float32)
        u = synth.hash_uniform(0xF00D + k, np.arange(2 * 9000)).reshape(2, 9000)
        mx = ((u[0] - 0.5) * 40.0).astype(np.float32)
        my = ((u[1] - 0.5) * 30.0).astype(np.float32)
        out.append((r, pose, mx, my, np.float32(1.0 if k % 2 == 0 else 0.35)))
    return out


def main():
    ref = Reference("accel")
    angles, range_min = ref.lidar_angles()
    g = {"angles": angles, "range_min": np.float32(range_min)}
    cs = cases()
    for k, (r, pose, mx, my, border) in enumerate(cs):
        x, y = ref.read_a_scan(r, 24)
        tx, ty = ref.transform(pose)
        if k == 4:                                   # map points exactly on the box: strict compares drop them
            mx[:4] = [tx.min() - border, tx.max() + border, tx[0], tx[0]]
            my[:4] = [ty[0], ty[0], ty.min() - border, ty.max() + border]
        lx, ly = ref.extract_local_map(mx, my, float(border))
        g.update({f"ranges_{k}": r, f"pose_{k}": pose, f"map_x_{k}": mx, f"map_y_{k}": my, f"border_{k}": border,
                  f"x_{k}": x, f"y_{k}": y, f"tx_{k}": tx, f"ty_{k}": ty, f"local_x_{k}": lx, f"local_y_{k}": ly})
    g["count"] = np.array(len(cs))
    np.savez_compressed(os.path.join(HERE, "frontend_golden.npz"), **g)
    print("wrote frontend_golden.npz:", {k: (len(g[f'x_{k}']), len(g[f'local_x_{k}'])) for k in range(len(cs))})


if __name__ == "__main__":
    main()
