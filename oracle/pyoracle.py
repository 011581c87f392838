"""ctypes access to the CPU oracle and to the compiled, unmodified reference.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never imports it.

* ``Oracle``     -> oracle/liboracle.so  (slam_oracle.c, the restatement)
* ``Reference``  -> oracle/_ref/libref_{main,accel,edtfrag}.so, the reference's own
  translation units compiled unmodified (oracle/Makefile).  Struct mirrors below restate
  the reference's global types: ScanData (Subsystem_1/main.c:60-69), MyGrid (:200-213),
  MyFastMatchParameters (:374-379).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int32)


def build(quiet: bool = True) -> None:
    """make -C oracle (restatement always; reference objects when /root/reference exists)."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


class OrcMap(C.Structure):
    _fields_ = [("field", c_float_p), ("rows", C.c_int), ("cols", C.c_int), ("stride", C.c_int),
                ("pixel_size", C.c_float), ("top_left_x", C.c_float), ("top_left_y", C.c_float)]


class OrcMatch(C.Structure):
    _fields_ = [("best_index", C.c_int64), ("best_score", C.c_float), ("best_pose", C.c_float * 3),
                ("best_hits", C.c_int), ("last_hits", C.c_int)]


class Oracle:
    """The CPU restatement (slam_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.orc_edt_radius.restype = C.c_int
        L.orc_edt_radius.argtypes = [C.c_float]
        for name in ("orc_edt", "orc_edt_percell", "orc_edt_scatter"):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [c_int_p, C.c_int, c_float_p, C.c_int, C.c_int, C.c_int, C.c_float]
        L.orc_lattice_value.restype = C.c_float
        L.orc_lattice_value.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int]
        L.orc_score_lattice.restype = None
        L.orc_score_lattice.argtypes = [C.POINTER(OrcMap), c_float_p, c_float_p, C.c_int, c_float_p,
                                        c_float_p, c_int_p, c_float_p, c_float_p, C.POINTER(OrcMatch)]
        L.orc_score_poses.restype = None
        L.orc_score_poses.argtypes = [C.POINTER(OrcMap), c_float_p, c_float_p, C.c_int, c_float_p,
                                      c_float_p, c_float_p, C.c_int64, c_float_p, c_int_p,
                                      C.POINTER(OrcMatch)]
        L.orc_fastmatch.restype = None
        L.orc_fastmatch.argtypes = [C.POINTER(OrcMap), c_float_p, c_float_p, C.c_int, c_float_p,
                                    c_float_p, c_float_p, c_float_p, c_int_p]
        L.orc_lidar_angles.restype = None
        L.orc_lidar_angles.argtypes = [C.c_float, C.c_float, C.c_int, c_float_p]
        L.orc_read_scan.restype = C.c_int
        L.orc_read_scan.argtypes = [c_float_p, c_float_p, C.c_int, C.c_float, C.c_int, c_float_p, c_float_p]
        L.orc_read_csv.restype = C.c_long
        L.orc_read_csv.argtypes = [C.c_char_p, c_float_p, C.c_long]
        L.orc_transform.restype = None
        L.orc_transform.argtypes = [c_float_p, c_float_p, C.c_int, c_float_p, c_float_p, c_float_p]
        L.orc_extract_local_map.restype = C.c_int
        L.orc_extract_local_map.argtypes = [c_float_p, c_float_p, C.c_int, c_float_p, c_float_p, C.c_int, C.c_float,
                                            c_float_p, c_float_p]
        L.orc_grow_map.restype = C.c_int
        L.orc_grow_map.argtypes = [c_float_p, C.c_int, c_float_p, c_float_p, c_float_p, c_float_p, C.c_int]
        L.orc_exp_det.restype = C.c_float
        L.orc_exp_det.argtypes = [C.c_float]
        L.orc_weights_resample.restype = None
        L.orc_weights_resample.argtypes = [c_float_p, C.c_int64, C.c_float, C.c_uint32, c_float_p,
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), c_int_p]
        L.orc_pyramid_match.restype = None
        L.orc_pyramid_match.argtypes = [C.POINTER(OrcMap), C.c_int, c_float_p, c_float_p, C.c_int,
                                        c_float_p, c_float_p, c_int_p, C.POINTER(OrcMatch)]

    # -- EDT ---------------------------------------------------------------
    def edt(self, occ: np.ndarray, max_dist: float = 10.0, variant: str = "separable") -> np.ndarray:
        occ = np.ascontiguousarray(occ, dtype=np.int32)
        rows, cols = occ.shape
        out = np.empty((rows, cols), dtype=np.float32)
        fn = {"separable": self.lib.orc_edt, "percell": self.lib.orc_edt_percell,
              "scatter": self.lib.orc_edt_scatter}[variant]
        fn(_ip(occ), cols, _fp(out), cols, rows, cols, C.c_float(max_dist))
        return out

    def occupational_grid(self, x, y, pixel_size: float, cap_rows: int, cap_cols: int):
        """-> (grid[rows][cols] int32, (min_x, min_y)) per Subsystem_1/main.c:271-354 (one level)."""
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        grid = np.empty((cap_rows, cap_cols), np.int32)
        rows, cols = C.c_int(0), C.c_int(0)
        mx, my = C.c_float(0), C.c_float(0)
        self.lib.orc_occupational_grid.restype = C.c_int
        rc = self.lib.orc_occupational_grid(_fp(x), _fp(y), len(x), C.c_float(pixel_size), _ip(grid), cap_cols,
                                            cap_rows, cap_cols, C.byref(rows), C.byref(cols), C.byref(mx),
                                            C.byref(my))
        if rc != 0:
            raise ValueError(f"grid {rows.value} x {cols.value} exceeds capacity {cap_rows} x {cap_cols}")
        return grid[:rows.value, :cols.value].copy(), (np.float32(mx.value), np.float32(my.value))

    # -- scoring -----------------------------------------------------------
    @staticmethod
    def make_map(field: np.ndarray, pixel_size: float, top_left, rows=None, cols=None):
        field = np.ascontiguousarray(field, dtype=np.float32)
        m = OrcMap(_fp(field), int(rows if rows is not None else field.shape[0]),
                   int(cols if cols is not None else field.shape[1]), int(field.shape[1]),
                   C.c_float(pixel_size), C.c_float(top_left[0]), C.c_float(top_left[1]))
        m._keep = field
        return m

    # -- scan front end / map points ---------------------------------------------------
    def lidar_angles(self, angle_min=-2.351831, inc=0.004363, n=1079):
        a = np.empty(n, np.float32)
        self.lib.orc_lidar_angles(angle_min, inc, n, _fp(a))
        return a

    def read_scan(self, ranges, angles, range_min=0.023, max_range=24):
        r = np.ascontiguousarray(ranges, np.float32); a = np.ascontiguousarray(angles, np.float32)
        x, y = np.empty(len(r), np.float32), np.empty(len(r), np.float32)
        n = self.lib.orc_read_scan(_fp(r), _fp(a), len(r), range_min, int(max_range), _fp(x), _fp(y))
        return x[:n].copy(), y[:n].copy()

    def read_csv(self, path: str, max_values: int) -> np.ndarray:
        """fscanf("%f,") over the whole file (main.c:22-30)."""
        out = np.empty(max(max_values, 1), np.float32)
        n = self.lib.orc_read_csv(path.encode(), _fp(out), max_values)
        assert n >= 0, path
        return out[:n].copy()

    def transform(self, x, y, pose):
        x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
        p = np.asarray(pose, np.float32)
        tx, ty = np.empty_like(x), np.empty_like(y)
        self.lib.orc_transform(_fp(x), _fp(y), len(x), _fp(p), _fp(tx), _fp(ty))
        return tx, ty

    def extract_local_map(self, tx, ty, map_x, map_y, border=1.0):
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        mx = np.ascontiguousarray(map_x, np.float32); my = np.ascontiguousarray(map_y, np.float32)
        lx, ly = np.empty(max(len(mx), 1), np.float32), np.empty(max(len(mx), 1), np.float32)
        n = self.lib.orc_extract_local_map(_fp(tx), _fp(ty), len(tx), _fp(mx), _fp(my), len(mx), border, _fp(lx), _fp(ly))
        return lx[:n].copy(), ly[:n].copy()

    def grow_map(self, best_hits, best_hits_size, tx, ty, map_x, map_y):
        bh = np.ascontiguousarray(best_hits, np.float32)
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        n0 = len(map_x)
        mx = np.empty(n0 + best_hits_size, np.float32); my = np.empty(n0 + best_hits_size, np.float32)
        mx[:n0] = map_x; my[:n0] = map_y
        k = self.lib.orc_grow_map(_fp(bh), int(best_hits_size), _fp(tx), _fp(ty), _fp(mx), _fp(my), n0)
        return mx[:n0 + k].copy(), my[:n0 + k].copy()

    def score_lattice(self, omap, scan_x, scan_y, pose0, step, n, want_scores=True, want_last_hits=False):
        sx = np.ascontiguousarray(scan_x, np.float32)
        sy = np.ascontiguousarray(scan_y, np.float32)
        p0 = np.asarray(pose0, np.float32)
        stp = np.asarray(step, np.float32)
        nn = np.asarray(n, np.int32)
        total = int(nn[0]) * int(nn[1]) * int(nn[2])
        scores = np.empty(total, np.float32) if want_scores else None
        lh = np.zeros(len(sx), np.float32) if want_last_hits else None
        res = OrcMatch()
        self.lib.orc_score_lattice(C.byref(omap), _fp(sx), _fp(sy), len(sx), _fp(p0), _fp(stp), _ip(nn),
                                   _fp(scores) if want_scores else None,
                                   _fp(lh) if want_last_hits else None, C.byref(res))
        return res, scores, lh

    def score_poses(self, omap, scan_x, scan_y, poses, ct=None, st=None):
        sx = np.ascontiguousarray(scan_x, np.float32)
        sy = np.ascontiguousarray(scan_y, np.float32)
        poses = np.ascontiguousarray(poses, np.float32)
        P = poses.shape[0]
        scores = np.empty(P, np.float32)
        hits = np.empty(P, np.int32)
        res = OrcMatch()
        ctp = _fp(np.ascontiguousarray(ct, np.float32)) if ct is not None else None
        stp = _fp(np.ascontiguousarray(st, np.float32)) if st is not None else None
        self.lib.orc_score_poses(C.byref(omap), _fp(sx), _fp(sy), len(sx), _fp(poses), ctp, stp, P,
                                 _fp(scores), _ip(hits), C.byref(res))
        return res, scores, hits

    def fastmatch(self, omap, scan_x, scan_y, pose, res3, hits_buf=None):
        """hits_buf: a persistent float32 array standing in for the global FastMatchParameters.bestHits
        (main.c:376): only its first last_hits entries are overwritten, the rest keeps older values."""
        sx = np.ascontiguousarray(scan_x, np.float32)
        sy = np.ascontiguousarray(scan_y, np.float32)
        p = np.asarray(pose, np.float32)
        r = np.asarray(res3, np.float32)
        out = np.zeros(3, np.float32)
        hits = np.zeros(max(len(sx), 1), np.float32) if hits_buf is None else hits_buf
        n = C.c_int32(0)
        self.lib.orc_fastmatch(C.byref(omap), _fp(sx), _fp(sy), len(sx), _fp(p), _fp(r), _fp(out), _fp(hits),
                               C.byref(n))
        return out, hits, n.value

    def exp_det(self, x: float) -> float:
        return float(self.lib.orc_exp_det(C.c_float(x)))

    def weights_resample(self, scores, beta: float, u0_q32: int):
        s = np.ascontiguousarray(scores, np.float32)
        N = len(s)
        w = np.empty(N, np.float32)
        q = np.empty(N, np.uint64)
        anc = np.empty(N, np.int32)
        W = C.c_uint64(0)
        self.lib.orc_weights_resample(_fp(s), N, C.c_float(beta), C.c_uint32(u0_q32), _fp(w),
                                      q.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(W), _ip(anc))
        return w, q, int(W.value), anc

    def pyramid_match(self, omaps, scan_x, scan_y, pose0, steps, ns):
        sx = np.ascontiguousarray(scan_x, np.float32)
        sy = np.ascontiguousarray(scan_y, np.float32)
        L = len(omaps)
        arr = (OrcMap * L)(*omaps)
        p0 = np.asarray(pose0, np.float32)
        stp = np.ascontiguousarray(steps, np.float32).reshape(L, 3)
        nn = np.ascontiguousarray(ns, np.int32).reshape(L, 3)
        res = (OrcMatch * L)()
        self.lib.orc_pyramid_match(arr, L, _fp(sx), _fp(sy), len(sx), _fp(p0), _fp(stp), _ip(nn), res)
        return list(res)


# ---------------------------------------------------------------------------
# The reference itself
# ---------------------------------------------------------------------------
COLUMN = 1079  # Subsystem_1/main.c:7


class RefScanData(C.Structure):  # main.c:60-66
    _fields_ = [("x", C.c_float * COLUMN), ("y", C.c_float * COLUMN), ("tx", C.c_float * COLUMN),
                ("ty", C.c_float * COLUMN), ("size", C.c_int)]


class RefMyGrid(C.Structure):  # main.c:200-212
    _fields_ = [("grid", (C.c_int * 200) * 200), ("grid_size", C.c_int * 2),
                ("metric_grid", (C.c_float * 200) * 200), ("pixel_size", C.c_float),
                ("top_left_corner", C.c_float * 2),
                ("grid2", (C.c_int * 400) * 400), ("grid_size2", C.c_int * 2),
                ("metric_grid2", (C.c_float * 400) * 400), ("pixel_size2", C.c_float),
                ("top_left_corner2", C.c_float * 2)]


class RefFastMatchParameters(C.Structure):  # main.c:374-378
    _fields_ = [("pose", C.c_float * 3), ("bestHits", C.c_float * 2500), ("bestHits_size", C.c_int)]


class RefLocalMap(C.Structure):  # main.c:147-151
    _fields_ = [("x", C.c_float * 25000), ("y", C.c_float * 25000), ("size", C.c_int)]


class RefLidar(C.Structure):  # main.c:35-42
    _fields_ = [("angle_min", C.c_float), ("angle_max", C.c_float), ("angle_increment", C.c_float),
                ("range_min", C.c_float), ("range_max", C.c_float), ("angles", C.c_float * COLUMN)]


class RefMapPoints(C.Structure):  # main.c:121-131
    _fields_ = [("x", C.c_float * 20000), ("y", C.c_float * 20000), ("size", C.c_int),
                ("newPoints_x", C.c_float * 4000), ("newPoints_y", C.c_float * 4000), ("newPointsSize", C.c_int),
                ("pose", C.c_float * 3)]


def reference_available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f))
               for f in ("libref_main.so", "libref_accel.so", "libref_edtfrag.so"))


class Reference:
    """The reference's own functions (compiled unmodified) behind numpy-friendly calls."""

    def __init__(self, which: str = "accel"):
        path = os.path.join(REF_DIR, {"main": "libref_main.so", "accel": "libref_accel.so",
                                      "edtfrag": "libref_edtfrag.so"}[which])
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: build with `make -C oracle` where /root/reference exists")
        self.which = which
        self.lib = C.CDLL(path)
        for name in ("euclidean_distance_transform", "euclidean_distance_transform2"):
            f = getattr(self.lib, name)
            f.restype = None
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        if which != "edtfrag":
            for name in ("FastMatch", "FastMatch2"):
                f = getattr(self.lib, name)
                f.restype = None
                f.argtypes = [c_float_p, c_float_p]
            self.scan = RefScanData.in_dll(self.lib, "scan")
            self.occ_grid = RefMyGrid.in_dll(self.lib, "occ_grid")
            self.fmp = RefFastMatchParameters.in_dll(self.lib, "FastMatchParameters")

    def read_dataset_rows(self, path: str, nrows: int) -> np.ndarray:
        """The reference's own readDatasetLineByLine (main.c:22-30) on `path`, nrows times: each call fills the
        global test_input_memory[1079] from a libc FILE*."""
        libc = C.CDLL("libc.so.6")
        libc.fopen.restype = C.c_void_p
        libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        libc.fclose.argtypes = [C.c_void_p]
        fn = self.lib.readDatasetLineByLine
        fn.restype = None
        fn.argtypes = [C.c_void_p]
        mem = (C.c_float * 1079).in_dll(self.lib, "test_input_memory")
        fp = libc.fopen(path.encode(), b"r")
        assert fp, path
        rows = np.empty((nrows, 1079), np.float32)
        for r in range(nrows):
            fn(fp)
            rows[r] = np.frombuffer(mem, np.float32)
        libc.fclose(fp)
        return rows

    def edt(self, occ: np.ndarray, fine: bool = False) -> np.ndarray:
        """Reference EDT on a rows x cols grid (<=200^2 coarse / <=400^2 fine), called exactly as
        OccupationalGrid does (main.c:355-356): 3rd arg = #columns, 4th arg = #rows."""
        S = 400 if fine else 200
        rows, cols = occ.shape
        assert rows <= S and cols <= S
        inp = np.zeros((S, S), np.int32)
        inp[:rows, :cols] = occ
        out = np.full((S, S), np.float32(-1.0), np.float32)
        fn = self.lib.euclidean_distance_transform2 if fine else self.lib.euclidean_distance_transform
        fn(inp.ctypes.data, out.ctypes.data, cols, rows)
        return out[:rows, :cols].copy()

    def set_map(self, field: np.ndarray, pixel_size: float, top_left, fine: bool):
        rows, cols = field.shape
        g = self.occ_grid
        S = 400 if fine else 200
        buf = np.zeros((S, S), np.float32)
        buf[:rows, :cols] = field
        if fine:
            C.memmove(g.metric_grid2, buf.ctypes.data, buf.nbytes)
            g.grid_size2[0], g.grid_size2[1] = rows, cols
            g.pixel_size2 = pixel_size
            g.top_left_corner2[0], g.top_left_corner2[1] = top_left
        else:
            C.memmove(g.metric_grid, buf.ctypes.data, buf.nbytes)
            g.grid_size[0], g.grid_size[1] = rows, cols
            g.pixel_size = pixel_size
            g.top_left_corner[0], g.top_left_corner[1] = top_left

    def occupational_grid(self, x, y, pixel_size: float, pixel_size2: float):
        """Runs the reference OccupationalGrid (main.c:271-363) on its own globals: local_map <- (x, y).
        Returns [(grid, field, (min_x, min_y)), (grid2, field2, (min_x2, min_y2))]."""
        n = len(x)
        assert n <= 25000
        lm = RefLocalMap.in_dll(self.lib, "local_map")
        xs = np.zeros(25000, np.float32); xs[:n] = x
        ys = np.zeros(25000, np.float32); ys[:n] = y
        C.memmove(lm.x, xs.ctypes.data, xs.nbytes)
        C.memmove(lm.y, ys.ctypes.data, ys.nbytes)
        lm.size = n
        f = self.lib.OccupationalGrid
        f.restype = None
        f.argtypes = [C.c_float, C.c_float]
        f(C.c_float(pixel_size), C.c_float(pixel_size2))
        g = self.occ_grid
        out = []
        for fine in (False, True):
            S = 400 if fine else 200
            rows, cols = (g.grid_size2[0], g.grid_size2[1]) if fine else (g.grid_size[0], g.grid_size[1])
            grid = np.ctypeslib.as_array(g.grid2 if fine else g.grid).reshape(S, S)[:rows, :cols].copy()
            field = np.ctypeslib.as_array(g.metric_grid2 if fine else g.metric_grid).reshape(S, S)[:rows, :cols].copy()
            tl = g.top_left_corner2 if fine else g.top_left_corner
            out.append((grid, field, (np.float32(tl[0]), np.float32(tl[1]))))
        return out

    # -- the reference's own front-end functions on its own globals -----------------------
    def lidar_angles(self):
        """SetLidarParameters() (main.c:45-58) -> its angle table and range_min."""
        self.lib.SetLidarParameters.restype = None
        self.lib.SetLidarParameters()
        lid = RefLidar.in_dll(self.lib, "lidar")
        return np.ctypeslib.as_array(lid.angles).copy(), float(lid.range_min)

    def read_a_scan(self, ranges, usable_range=24):
        """test_input_memory <- ranges; readAScan(usable_range) (main.c:71-95) -> scan.x, scan.y."""
        r = np.ascontiguousarray(ranges, np.float32)
        assert len(r) == COLUMN
        tim = (C.c_float * COLUMN).in_dll(self.lib, "test_input_memory")
        C.memmove(tim, r.ctypes.data, r.nbytes)
        self.lib.readAScan.restype = None
        self.lib.readAScan.argtypes = [C.c_int]
        self.lib.readAScan(int(usable_range))
        n = int(self.scan.size)
        return np.ctypeslib.as_array(self.scan.x)[:n].copy(), np.ctypeslib.as_array(self.scan.y)[:n].copy()

    def transform(self, pose):
        """Transform(pose) (main.c:97-118) on the current scan -> scan.tx, scan.ty."""
        p = np.asarray(pose, np.float32)
        self.lib.Transform.restype = None
        self.lib.Transform.argtypes = [c_float_p]
        self.lib.Transform(_fp(p))
        n = int(self.scan.size)
        return np.ctypeslib.as_array(self.scan.tx)[:n].copy(), np.ctypeslib.as_array(self.scan.ty)[:n].copy()

    def extract_local_map(self, map_x, map_y, border=1.0):
        """map <- (map_x, map_y); ExtractLocalMap(border) (main.c:155-198) -> local_map."""
        mp = RefMapPoints.in_dll(self.lib, "map")
        n = len(map_x)
        assert n <= 20000
        xs = np.zeros(20000, np.float32); xs[:n] = map_x
        ys = np.zeros(20000, np.float32); ys[:n] = map_y
        C.memmove(mp.x, xs.ctypes.data, xs.nbytes)
        C.memmove(mp.y, ys.ctypes.data, ys.nbytes)
        mp.size = n
        self.lib.ExtractLocalMap.restype = None
        self.lib.ExtractLocalMap.argtypes = [C.c_float]
        self.lib.ExtractLocalMap(C.c_float(border))
        lm = RefLocalMap.in_dll(self.lib, "local_map")
        k = int(lm.size)
        return np.ctypeslib.as_array(lm.x)[:k].copy(), np.ctypeslib.as_array(lm.y)[:k].copy()

    def set_scan(self, x, y):
        n = len(x)
        assert n <= COLUMN
        xs = np.zeros(COLUMN, np.float32)
        ys = np.zeros(COLUMN, np.float32)
        xs[:n] = x
        ys[:n] = y
        C.memmove(self.scan.x, xs.ctypes.data, xs.nbytes)
        C.memmove(self.scan.y, ys.ctypes.data, ys.nbytes)
        self.scan.size = n

    def fastmatch(self, pose, res3, fine: bool):
        """Calls the reference FastMatch (coarse) / FastMatch2 (fine); returns pose, bestHits[:n], n."""
        p = np.asarray(pose, np.float32)
        r = np.asarray(res3, np.float32)
        (self.lib.FastMatch2 if fine else self.lib.FastMatch)(_fp(p), _fp(r))
        n = int(self.fmp.bestHits_size)
        hits = np.ctypeslib.as_array(self.fmp.bestHits).copy()
        return np.array(list(self.fmp.pose), np.float32), hits, n
