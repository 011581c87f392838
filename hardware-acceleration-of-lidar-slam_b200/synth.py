"""Synthetic occupancy grids, lidar scans, candidate lattices and particle sets
(SURVEY.md section 8d).  Everything derives from a counter-based integer hash
(splitmix64 of seed ^ counter), so the same inputs can be regenerated anywhere -- tests,
bench.py on the GPU box, the golden-fixture script -- without shipping data.

Conventions follow the reference: grid[row = y][col = x]; a world point W is seen from
pose (px, py, theta) at scan = R(theta) (W - p), because Transform applies R(-theta)
(Subsystem_1/main.c:115-116).
"""
from __future__ import annotations

import numpy as np

SEED_GRID = 0x5EED0001
SEED_SCAN = 0x5EED0002
SEED_PARTICLES = 0x5EED0003

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def hash_uniform(seed: int, counter: np.ndarray) -> np.ndarray:
    """float64 in [0, 1) from (seed, counter)."""
    h = splitmix64(np.uint64(seed) ^ np.asarray(counter, dtype=np.uint64))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def grid_bernoulli(rows: int, cols: int, p: float, seed: int = SEED_GRID) -> np.ndarray:
    """int32 occupancy, each cell occupied with probability p (hash of seed ^ (r<<32|c))."""
    out = np.empty((rows, cols), np.int32)
    thr = np.uint64(min(int(p * 2.0 ** 64), 2 ** 64 - 1))
    c = np.arange(cols, dtype=np.uint64)
    step = max(1, (1 << 22) // max(cols, 1))
    for r0 in range(0, rows, step):
        r = np.arange(r0, min(r0 + step, rows), dtype=np.uint64)
        ctr = (r[:, None] << np.uint64(32)) | c[None, :]
        out[r0:r0 + len(r)] = (splitmix64(np.uint64(seed) ^ ctr) < thr).astype(np.int32)
    return out


def grid_rooms(rows: int, cols: int, seed: int = SEED_GRID, n_segments: int | None = None,
               n_pillars: int | None = None) -> np.ndarray:
    """Outer wall rectangle + hashed axis-aligned wall segments + square pillars
    (a few % occupancy, like a rasterised indoor map)."""
    occ = np.zeros((rows, cols), np.int32)
    m = max(4, min(rows, cols) // 64)
    occ[m, m:cols - m] = 1
    occ[rows - 1 - m, m:cols - m] = 1
    occ[m:rows - m, m] = 1
    occ[m:rows - m, cols - 1 - m] = 1
    if n_segments is None:
        n_segments = max(8, (rows * cols) // 60000)
    if n_pillars is None:
        n_pillars = max(6, (rows * cols) // 90000)
    u = hash_uniform(seed, np.arange(8 * (n_segments + n_pillars)))
    k = 0
    for _ in range(n_segments):
        r0 = m + int(u[k] * (rows - 2 * m)); c0 = m + int(u[k + 1] * (cols - 2 * m))
        ln = 8 + int(u[k + 2] * min(rows, cols) * 0.18)
        if u[k + 3] < 0.5:
            occ[r0, c0:min(c0 + ln, cols - m)] = 1
        else:
            occ[r0:min(r0 + ln, rows - m), c0] = 1
        k += 8
    for _ in range(n_pillars):
        r0 = m + int(u[k] * (rows - 2 * m - 6)); c0 = m + int(u[k + 1] * (cols - 2 * m - 6))
        s = 2 + int(u[k + 2] * 4)
        occ[r0:r0 + s, c0:c0 + s] = 1
        k += 8
    return occ


def centred_geometry(rows: int, cols: int, pixel: float):
    """top_left = (-W*pixel/2, -H*pixel/2) as float32."""
    return np.float32(pixel), (np.float32(-cols * pixel / 2), np.float32(-rows * pixel / 2))


def beam_angles(nbeams: int, reference_lidar: bool = False) -> np.ndarray:
    """Beam angles: 2*pi/B spacing, or the reference's 270-degree geometry accumulated in
    float like Subsystem_1/main.c:47-57."""
    if reference_lidar:
        a = np.empty(nbeams, np.float32)
        ang = np.float32(-2.351831)
        inc = np.float32(0.004363)
        for i in range(nbeams):
            a[i] = ang
            ang = np.float32(ang + inc)
        return a
    return (np.arange(nbeams, dtype=np.float64) * (2.0 * np.pi / nbeams) - np.pi).astype(np.float32)


def scan_raycast(occ: np.ndarray, pixel: float, top_left, pose, nbeams: int, max_range: float = 24.0,
                 reference_lidar: bool = False, noise_seed: int | None = None):
    """Ray-cast `occ` from `pose` (reference convention) -> sensor-frame scan points
    (float32 x, y) of the beams that hit something within max_range."""
    rows, cols = occ.shape
    ang = beam_angles(nbeams, reference_lidar).astype(np.float64)
    px, py, th = float(pose[0]), float(pose[1]), float(pose[2])
    ct, st = np.cos(th), np.sin(th)
    dx = ct * np.cos(ang) + st * np.sin(ang)       # world direction = R(-theta) * beam
    dy = -st * np.cos(ang) + ct * np.sin(ang)
    step = 0.5 * pixel
    nsteps = int(max_range / step)
    rng = np.full(nbeams, np.inf)
    alive = np.ones(nbeams, bool)
    for s0 in range(1, nsteps + 1, 64):
        ss = np.arange(s0, min(s0 + 64, nsteps + 1)) * step
        wx = px + dx[:, None] * ss[None, :]
        wy = py + dy[:, None] * ss[None, :]
        c = np.rint((wx - float(top_left[0])) / pixel).astype(np.int64)
        r = np.rint((wy - float(top_left[1])) / pixel).astype(np.int64)
        inb = (c >= 0) & (c < cols) & (r >= 0) & (r < rows)
        hit = np.zeros_like(inb)
        hit[inb] = occ[r[inb], c[inb]] != 0
        anyhit = hit.any(axis=1) & alive
        first = hit.argmax(axis=1)
        rng[anyhit] = ss[first[anyhit]]
        alive &= ~anyhit
        if not alive.any():
            break
    ok = np.isfinite(rng)
    rr = rng[ok]
    if noise_seed is not None:
        rr = rr + (hash_uniform(noise_seed, np.nonzero(ok)[0]) - 0.5) * 0.02
    x = (rr * np.cos(ang[ok])).astype(np.float32)
    y = (rr * np.sin(ang[ok])).astype(np.float32)
    return x, y


def scan_random(nbeams: int, seed: int = SEED_SCAN, rmin: float = 2.0, rspan: float = 18.0):
    """r_k = rmin + rspan * u(hash): for the Bernoulli grids (no geometry to ray-cast)."""
    ang = beam_angles(nbeams).astype(np.float64)
    r = rmin + rspan * hash_uniform(seed, np.arange(nbeams))
    return (r * np.cos(ang)).astype(np.float32), (r * np.sin(ang)).astype(np.float32)


def scan_fixed_count(occ, pixel, top_left, pose, nbeams: int, seed: int = SEED_SCAN):
    """Exactly `nbeams` sensor-frame points: ray-cast hits, topped up with hashed ranges
    for beams that saw nothing (keeps pose x beam counts at their nominal value)."""
    x, y = scan_raycast(occ, pixel, top_left, pose, nbeams)
    if len(x) < nbeams:
        fx, fy = scan_random(nbeams - len(x), seed)
        x = np.concatenate([x, fx]).astype(np.float32)
        y = np.concatenate([y, fy]).astype(np.float32)
    return x[:nbeams], y[:nbeams]


def particles_gaussian(n: int, centre, sigma_xy: float = 0.25, sigma_th: float = 0.05,
                       seed: int = SEED_PARTICLES) -> np.ndarray:
    """n poses = centre + Gaussian noise from a hashed Box-Muller in double; float32 [n][3]."""
    idx = np.arange(n, dtype=np.uint64)
    out = np.empty((n, 3), np.float32)
    for d, sig in enumerate((sigma_xy, sigma_xy, sigma_th)):
        u1 = hash_uniform(seed + 17 * d, idx * np.uint64(2))
        u2 = hash_uniform(seed + 17 * d, idx * np.uint64(2) + np.uint64(1))
        z = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
        out[:, d] = (float(centre[d]) + sig * z).astype(np.float32)
    return out


# BASELINE.json workloads --------------------------------------------------------------
WORKLOADS = {
    # configs[1]: synthetic 2048x2048 grid EDT + 64k candidate poses x 360 beams
    "config1": dict(rows=2048, cols=2048, pixel=0.1, nbeams=360, n=(64, 32, 32), grid="rooms"),
    # configs[3]: synthetic 8192x8192 grid EDT + 4M candidate poses x 1080 beams
    "config3": dict(rows=8192, cols=8192, pixel=0.1, nbeams=1080, n=(256, 128, 128), grid="rooms"),
    # small variant for smoke / CI
    "tiny": dict(rows=256, cols=320, pixel=0.1, nbeams=180, n=(8, 16, 16), grid="rooms"),
}
LATTICE_STEP = (0.05, 0.05, 0.008727)          # Subsystem_1/main.c:832
TRUE_POSE_OFFSET = (0.11, -0.07, 0.013)        # lattice centre = true pose + this


def make_workload(name: str, seed: int = SEED_GRID):
    """Returns dict(occ, pixel, top_left, scan_x, scan_y, pose0, step, n) for a named workload."""
    w = WORKLOADS[name]
    rows, cols = w["rows"], w["cols"]
    occ = grid_rooms(rows, cols, seed) if w["grid"] == "rooms" else grid_bernoulli(rows, cols, 0.01, seed)
    pixel, tl = centred_geometry(rows, cols, w["pixel"])
    true_pose = (0.37, -0.21, 0.1)
    sx, sy = scan_fixed_count(occ, float(pixel), tl, true_pose, w["nbeams"])
    pose0 = np.array([true_pose[i] + TRUE_POSE_OFFSET[i] for i in range(3)], np.float32)
    return dict(name=name, occ=occ, pixel=pixel, top_left=tl, scan_x=sx, scan_y=sy, pose0=pose0,
                step=np.array(LATTICE_STEP, np.float32), n=tuple(w["n"]), true_pose=true_pose)


# Synthetic stand-in for the reference's missing lidar_dataset.csv -------------------------
REF_BEAMS = 1079                      # Subsystem_1/main.c:7  (#define column 1079)
REF_ANGLE_MIN = -2.351831             # Subsystem_1/main.c:48
REF_ANGLE_INC = 0.004363              # Subsystem_1/main.c:50


def room_segments(width: float = 18.0, height: float = 12.0, pillars=((-3.0, 0.0), (3.5, 2.0), (-7.5, 4.0), (5.0, -3.0), (2.0, -4.5)),
                  pillar_size: float = 0.8) -> np.ndarray:
    """Axis-aligned wall segments [x0, y0, x1, y1] of a closed room with square pillars."""
    hw, hh = width / 2, height / 2
    segs = [(-hw, -hh, hw, -hh), (-hw, hh, hw, hh), (-hw, -hh, -hw, hh), (hw, -hh, hw, hh)]
    for cx, cy in pillars:
        h = pillar_size / 2
        segs += [(cx - h, cy - h, cx + h, cy - h), (cx - h, cy + h, cx + h, cy + h),
                 (cx - h, cy - h, cx - h, cy + h), (cx + h, cy - h, cx + h, cy + h)]
    return np.array(segs, np.float64)


def raycast_segments(segs: np.ndarray, px: float, py: float, ang: np.ndarray) -> np.ndarray:
    """Range of each ray (origin (px, py), world angle ang[k]) to the nearest segment."""
    dx, dy = np.cos(ang)[:, None], np.sin(ang)[:, None]
    x0, y0, x1, y1 = (segs[None, :, i] for i in range(4))
    vert = (x0 == x1)
    with np.errstate(divide="ignore", invalid="ignore"):
        tv = (x0 - px) / dx                                   # vertical segments: x = x0
        yv = py + tv * dy
        okv = vert & (tv > 1e-9) & (yv >= np.minimum(y0, y1)) & (yv <= np.maximum(y0, y1))
        th = (y0 - py) / dy                                   # horizontal segments: y = y0
        xh = px + th * dx
        okh = (~vert) & (th > 1e-9) & (xh >= np.minimum(x0, x1)) & (xh <= np.maximum(x0, x1))
    t = np.where(okv, tv, np.where(okh, th, np.inf))
    return t.min(axis=1)


def lidar_dataset(nscans: int, radius: float = 3.0, loops: float = 1.0, seed: int = SEED_SCAN,
                  noise: float = 0.004) -> np.ndarray:
    """[nscans][1079] ranges of the reference's 270-degree lidar (Subsystem_1/main.c:45-58) carried
    slowly around a circle inside room_segments(): a stand-in for the bundled lidar_dataset.csv,
    which is not in the reference tree (.MISSING_LARGE_BLOBS).  Motion per scan stays far below
    the +-1-step search window of FastMatch so the reference tracks it."""
    segs = room_segments()
    k = np.arange(REF_BEAMS, dtype=np.float64)
    beam = REF_ANGLE_MIN + k * REF_ANGLE_INC
    out = np.empty((nscans, REF_BEAMS), np.float32)
    for s in range(nscans):
        phi = 2.0 * np.pi * loops * s / max(nscans - 1, 1)
        px, py = radius * np.cos(phi) - radius, radius * np.sin(phi)     # starts at the origin
        heading = phi + np.pi / 2                                         # tangent to the circle
        r = raycast_segments(segs, px, py, heading + beam)
        if noise:
            r = r + (hash_uniform(seed + s, np.arange(REF_BEAMS)) - 0.5) * 2.0 * noise
        out[s] = r
    return out


def write_lidar_csv(path: str, ranges: np.ndarray) -> None:
    """The reference reads `%f,` x 1079 per scan (Subsystem_1/main.c:22-30)."""
    with open(path, "w") as f:
        for row in ranges:
            f.write(",".join("%.4f" % v for v in row))
            f.write(",\n")
