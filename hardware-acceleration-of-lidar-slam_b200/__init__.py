"""b200slam -- ctypes binding of libb200slam.so (include/b200slam.h).

This package is the harness-side view of the C ABI: thin wrappers that hand numpy (host)
buffers to the library, used by tests/, bench.py and __graft_entry__.py.  It contains no
numerical code of its own and NO fallback: if libb200slam.so is missing or no GPU is
usable it raises -- the CUDA library is the only implementation of the hot path.

Import it with ``importlib.import_module("hardware-acceleration-of-lidar-slam_b200")``
(the directory name is not a Python identifier).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
# B200SLAM_LIB: development aid (A/B of two builds of the same ABI on the GPU box); never a fallback
LIB_PATH = os.environ.get("B200SLAM_LIB") or os.path.join(PKG_DIR, "libb200slam.so")
DROPIN_PATH = os.path.join(PKG_DIR, "libb200slam_dropin.so")
HEADER_PATH = os.path.join(REPO_ROOT, "include", "b200slam.h")

OK, ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_NOMEM, ERR_STATE = 0, -1, -2, -3, -4, -5
UNIQUE_ID_BYTES = 128
MATCH_LATENCY, MATCH_THROUGHPUT = 0, 1
EDT_GATHER_NCCL, EDT_GATHER_P2P = 0, 1

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u64_p = C.POINTER(C.c_uint64)


class B200SlamError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200slam error {code}: {msg}")
        self.code = code


class Match(C.Structure):
    """b200slam_match (include/b200slam.h)."""
    _fields_ = [("best_index", C.c_int64), ("best_score", C.c_float), ("best_pose", C.c_float * 3),
                ("best_hits", C.c_int32), ("last_hits", C.c_int32)]

    def pose(self) -> np.ndarray:
        return np.array(list(self.best_pose), np.float32)


def build(jobs: int = 8, quiet: bool = True) -> None:
    """Compile libb200slam.so / libb200slam_dropin.so in-tree for sm_100a (nvcc cross-compiles
    without a GPU)."""
    cmd = ["make", "-C", PKG_DIR, f"-j{jobs}"] + (["-s"] if quiet else [])
    subprocess.run(cmd, check=True)


_lib = None


def load_library() -> C.CDLL:
    """dlopen libb200slam.so and declare every prototype of include/b200slam.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: run `make -C {PKG_DIR}` (or __graft_entry__.build()). "
            "There is no CPU fallback for the b200slam hot path.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    sig = {
        "b200slam_abi_version": (i, []),
        "b200slam_create": (i, [C.POINTER(vp), i]),
        "b200slam_destroy": (None, [vp]),
        "b200slam_last_error": (C.c_char_p, [vp]),
        "b200slam_sync": (i, [vp]),
        "b200slam_stream": (vp, [vp]),
        "b200slam_launch_count": (C.c_uint64, [vp]),
        "b200slam_device_info": (i, [vp, C.c_char_p, c_int_p, c_int_p, c_int_p, C.POINTER(C.c_size_t)]),
        "b200slam_host_alloc": (i, [vp, C.c_size_t, C.POINTER(vp)]),
        "b200slam_host_free": (i, [vp, vp]),
        "b200slam_edt": (i, [vp, vp, i, vp, i, i, i, f]),
        "b200slam_map_create": (i, [vp, i, i, C.POINTER(vp)]),
        "b200slam_map_destroy": (None, [vp, vp]),
        "b200slam_map_set_geometry": (i, [vp, f, f, f]),
        "b200slam_map_upload_occupancy": (i, [vp, vp, vp, i]),
        "b200slam_map_resize": (i, [vp, i, i]),
        "b200slam_map_rasterise": (i, [vp, vp, vp, vp, i, f, c_int_p, c_int_p, c_float_p]),
        "b200slam_map_rasterise_async": (i, [vp, vp, vp, vp, i, f, c_int_p, c_int_p, c_float_p]),
        "b200slam_map_download_occupancy": (i, [vp, vp, vp, i]),
        "b200slam_map_edt": (i, [vp, vp, f]),
        "b200slam_map_download_field": (i, [vp, vp, vp, i]),
        "b200slam_map_upload_field": (i, [vp, vp, vp, i]),
        "b200slam_map_device_ptrs": (i, [vp, C.POINTER(vp), c_int_p, C.POINTER(vp), c_int_p]),
        "b200slam_scan_upload": (i, [vp, vp, vp, i]),
        "b200slam_lattice_value": (f, [f, f, i, i]),
        "b200slam_score_lattice": (i, [vp, vp, c_float_p, c_float_p, c_int_p, vp, vp, C.POINTER(Match)]),
        "b200slam_score_lattice_rows": (i, [vp, vp, c_float_p, c_float_p, c_int_p, C.c_int64, C.c_int64, i,
                                            C.POINTER(Match)]),
        "b200slam_score_lattice_async": (i, [vp, vp, c_float_p, c_float_p, c_int_p, C.c_int64, C.c_int64, i]),
        "b200slam_exchange_collect_async": (i, [vp]),
        "b200slam_match_fetch": (i, [vp, C.POINTER(Match)]),
        "b200slam_match_fetch_hits": (i, [vp, vp, i]),
        "b200slam_score_poses": (i, [vp, vp, vp, vp, vp, C.c_int64, C.c_int64, vp, vp, C.POINTER(Match)]),
        "b200slam_fastmatch": (i, [vp, vp, c_float_p, c_float_p, c_float_p, vp, c_int_p]),
        "b200slam_set_match_mode": (i, [vp, i]),
        "b200slam_lidar_set": (i, [vp, c_float_p, c_float_p, i, f]),
        "b200slam_scan_read": (i, [vp, c_float_p, i, c_int_p]),
        "b200slam_scan_read_async": (i, [vp, c_float_p, i]),
        "b200slam_scan_read_resident_async": (i, [vp, C.c_int64, i]),
        "b200slam_fastmatch_pair_async": (i, [vp, vp, vp, c_float_p, c_float_p, c_float_p]),
        "b200slam_fastmatch_pair_fetch": (i, [vp, c_float_p, c_float_p, c_int_p, c_int_p]),
        "b200slam_mappoints_grow_async": (i, [vp, f]),
        "b200slam_scan_step_async": (i, [vp, c_float_p, i, vp, vp, c_float_p, c_float_p, c_float_p]),
        "b200slam_scan_step_resident_async": (i, [vp, C.c_int64, i, vp, vp, c_float_p, c_float_p, c_float_p]),
        "b200slam_scan_chain_begin": (i, [vp, i, c_float_p, c_float_p, c_float_p, f, f]),
        "b200slam_scan_chain_step_async": (i, [vp, i, i, C.c_int64, i, vp, vp, c_float_p, c_float_p]),
        "b200slam_scan_chain_fetch": (i, [vp, i, c_float_p, c_float_p, c_int_p, c_int_p, c_int_p]),
        "b200slam_csv_ingest": (i, [vp, vp, C.c_size_t, vp, C.c_int64, c_i64_p]),
        "b200slam_csv_values": (i, [vp, C.POINTER(vp), c_i64_p]),
        "b200slam_scan_transform": (i, [vp, c_float_p]),
        "b200slam_scan_download": (i, [vp, c_float_p, c_float_p, c_float_p, c_float_p, c_int_p]),
        "b200slam_mappoints_upload": (i, [vp, c_float_p, c_float_p, i, i]),
        "b200slam_mappoints_from_scan": (i, [vp]),
        "b200slam_mappoints_grow": (i, [vp, f, c_int_p]),
        "b200slam_mappoints_download": (i, [vp, c_float_p, c_float_p, c_int_p]),
        "b200slam_local_map_extract": (i, [vp, f, c_int_p]),
        "b200slam_local_map_download": (i, [vp, c_float_p, c_float_p, c_int_p]),
        "b200slam_map_rasterise_local": (i, [vp, vp, f, c_int_p, c_int_p, c_float_p]),
        "b200slam_map_edt_rows": (i, [vp, vp, f, i, i]),
        "b200slam_map_share": (i, [vp, vp]),
        "b200slam_map_edt_sharded": (i, [vp, vp, f, i]),
        "b200slam_graph_begin": (i, [vp]),
        "b200slam_graph_end": (i, [vp, C.POINTER(vp)]),
        "b200slam_graph_launch": (i, [vp, vp]),
        "b200slam_graph_destroy": (None, [vp, vp]),
        "b200slam_event_record": (i, [vp, i]),
        "b200slam_event_wait": (i, [vp, vp, i]),
        "b200slam_event_elapsed_ms": (i, [vp, i, i, c_float_p]),
        "b200slam_weights_resample": (i, [vp, f, C.c_uint32, vp, c_u64_p, vp, c_i64_p, c_i64_p]),
        "b200slam_particles_upload": (i, [vp, vp, vp, vp, C.c_int64]),
        "b200slam_particles_shard": (i, [vp, vp, vp, vp, C.c_int64, C.c_int64, C.c_int64]),
        "b200slam_particles_score_async": (i, [vp, vp]),
        "b200slam_particles_resample_async": (i, [vp, f, C.c_uint32]),
        "b200slam_particles_download": (i, [vp, vp, vp, vp, vp]),
        "b200slam_resample_owned_slots": (None, [C.c_uint64, C.c_int64, C.c_uint32, C.c_uint64, C.c_uint64,
                                                 c_i64_p, c_i64_p]),
        "b200slam_pyramid_match": (i, [vp, C.POINTER(vp), i, c_float_p, c_float_p, c_int_p, C.POINTER(Match)]),
        "b200slam_comm_unique_id": (i, [vp]),
        "b200slam_comm_init": (i, [vp, i, i, vp]),
        "b200slam_comm_destroy": (i, [vp]),
        "b200slam_comm_barrier_async": (i, [vp]),
        "b200slam_shard_range": (None, [C.c_int64, i, i, c_i64_p, c_i64_p]),
        "b200slam_pack_key": (C.c_uint64, [f, C.c_uint32]),
        "b200slam_unpack_key": (None, [C.c_uint64, c_float_p, C.POINTER(C.c_uint32)]),
        "b200slam_merge_keys": (C.c_uint64, [c_u64_p, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)      # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = None


def declared_symbols() -> list[str]:
    """Every function name declared in include/b200slam.h (parsed from the header text)."""
    import re
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200slam_[a-z0-9_]+)\s*\(", text)))


# -- pure-host helpers (no GPU needed) ---------------------------------------------------
def shard_range(total: int, nranks: int, rank: int) -> tuple[int, int]:
    L = load_library()
    b, e = C.c_int64(0), C.c_int64(0)
    L.b200slam_shard_range(total, nranks, rank, C.byref(b), C.byref(e))
    return b.value, e.value


def resample_owned_slots(w_global: int, n_global: int, u0_q32: int, rank_offset: int, w_local: int):
    kb, kc = C.c_int64(0), C.c_int64(0)
    load_library().b200slam_resample_owned_slots(w_global, n_global, u0_q32, rank_offset, w_local,
                                                 C.byref(kb), C.byref(kc))
    return kb.value, kc.value


def pack_key(score: float, index: int) -> int:
    return int(load_library().b200slam_pack_key(C.c_float(score), C.c_uint32(index)))


def unpack_key(key: int) -> tuple[float, int]:
    s, ix = C.c_float(0), C.c_uint32(0)
    load_library().b200slam_unpack_key(C.c_uint64(key), C.byref(s), C.byref(ix))
    return s.value, ix.value


def merge_keys(keys) -> int:
    a = np.ascontiguousarray(keys, np.uint64)
    return int(load_library().b200slam_merge_keys(a.ctypes.data_as(c_u64_p), len(a)))


def lattice_value(p: float, s: float, k: int, n: int) -> float:
    return float(load_library().b200slam_lattice_value(C.c_float(p), C.c_float(s), k, n))


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(c_float_p)


_LIBM = None


def libm_cosf_sinf(a: np.ndarray):
    """glibc cosf / sinf element by element (what the reference calls: main.c:90-91, 101-102)."""
    global _LIBM
    if _LIBM is None:
        _LIBM = C.CDLL("libm.so.6")
        for nm in ("cosf", "sinf"):
            getattr(_LIBM, nm).restype = C.c_float
            getattr(_LIBM, nm).argtypes = [C.c_float]
    ca = np.array([_LIBM.cosf(float(v)) for v in a], np.float32)
    sa = np.array([_LIBM.sinf(float(v)) for v in a], np.float32)
    return ca, sa


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _i3(v):
    return (C.c_int32 * 3)(*[int(x) for x in v])


class Map:
    """Device-resident MyGrid (occupancy + distance field + geometry)."""

    def __init__(self, ctx: "Context", rows: int, cols: int):
        self.ctx, self.rows, self.cols = ctx, rows, cols
        h = C.c_void_p()
        ctx._check(ctx.L.b200slam_map_create(ctx.h, rows, cols, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.ctx.L.b200slam_map_destroy(self.ctx.h, self.h)
            self.h = None

    def set_geometry(self, pixel_size: float, top_left):
        self.ctx._check(self.ctx.L.b200slam_map_set_geometry(self.h, pixel_size, top_left[0], top_left[1]))
        return self

    def upload_occupancy(self, occ: np.ndarray):
        assert occ.dtype == np.int32 and occ.ndim == 2 and occ.shape[0] == self.rows and occ.shape[1] >= self.cols
        assert occ.strides[1] == 4
        self.ctx._check(self.ctx.L.b200slam_map_upload_occupancy(self.ctx.h, self.h, occ.ctypes.data,
                                                                 occ.strides[0] // 4))
        return self

    def rasterise(self, x, y, pixel_size: float):
        """Map points -> occupancy on the device (one level of OccupationalGrid, main.c:271-354);
        the map takes the rasterised grid's size and geometry.  -> (rows, cols, (min_x, min_y))"""
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        r, c = C.c_int32(0), C.c_int32(0)
        tl = (C.c_float * 2)()
        self.ctx._check(self.ctx.L.b200slam_map_rasterise(self.ctx.h, self.h, x.ctypes.data, y.ctypes.data, len(x),
                                                          pixel_size, C.byref(r), C.byref(c), tl))
        self.rows, self.cols = r.value, c.value
        return r.value, c.value, (np.float32(tl[0]), np.float32(tl[1]))

    def rasterise_async(self, x, y, pixel_size: float, ctx=None):
        """rasterise() with x, y in page-locked memory (Context.pinned_empty): queued straight from the
        caller's arrays, which must stay unchanged until the next synchronising call.  ctx: queue on
        another context of the same GPU (double buffering)."""
        assert x.dtype == np.float32 and y.dtype == np.float32 and x.flags.c_contiguous and y.flags.c_contiguous
        c_ = ctx or self.ctx
        r, c = C.c_int32(0), C.c_int32(0)
        tl = (C.c_float * 2)()
        c_._check(c_.L.b200slam_map_rasterise_async(c_.h, self.h, x.ctypes.data, y.ctypes.data, len(x),
                                                    pixel_size, C.byref(r), C.byref(c), tl))
        self.rows, self.cols = r.value, c.value
        return r.value, c.value, (np.float32(tl[0]), np.float32(tl[1]))

    def rasterise_local(self, pixel_size: float):
        """The same from the device-resident local map (Context.local_map_extract)."""
        r, c = C.c_int32(0), C.c_int32(0)
        tl = (C.c_float * 2)()
        self.ctx._check(self.ctx.L.b200slam_map_rasterise_local(self.ctx.h, self.h, pixel_size, C.byref(r), C.byref(c), tl))
        self.rows, self.cols = r.value, c.value
        return r.value, c.value, (np.float32(tl[0]), np.float32(tl[1]))

    def download_occupancy(self) -> np.ndarray:
        out = np.empty((self.rows, self.cols), np.int32)
        self.ctx._check(self.ctx.L.b200slam_map_download_occupancy(self.ctx.h, self.h, out.ctypes.data, self.cols))
        return out

    def edt(self, max_dist: float = 10.0):
        self.ctx._check(self.ctx.L.b200slam_map_edt(self.ctx.h, self.h, max_dist))
        return self

    def edt_rows(self, row_begin: int, row_end: int, max_dist: float = 10.0):
        """Output rows [row_begin, row_end) only (this GPU)."""
        self.ctx._check(self.ctx.L.b200slam_map_edt_rows(self.ctx.h, self.h, max_dist, row_begin, row_end))
        return self

    def share(self):
        """Collective: maps every rank's copy of this map's field through CUDA IPC."""
        self.ctx._check(self.ctx.L.b200slam_map_share(self.ctx.h, self.h))
        return self

    def edt_sharded(self, mode: int = EDT_GATHER_NCCL, max_dist: float = 10.0):
        """Collective: every rank transforms its row block; blocks exchanged by NCCL or inside the kernel."""
        self.ctx._check(self.ctx.L.b200slam_map_edt_sharded(self.ctx.h, self.h, max_dist, mode))
        return self

    def download_field(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.rows, self.cols), np.float32)
        self.ctx._check(self.ctx.L.b200slam_map_download_field(self.ctx.h, self.h, out.ctypes.data,
                                                               out.strides[0] // 4))
        return out

    def upload_field(self, field: np.ndarray):
        assert field.dtype == np.float32 and field.shape[0] == self.rows and field.strides[1] == 4
        self.ctx._check(self.ctx.L.b200slam_map_upload_field(self.ctx.h, self.h, field.ctypes.data,
                                                             field.strides[0] // 4))
        return self

    def device_ptrs(self):
        occ, fld = C.c_void_p(), C.c_void_p()
        op, fp = C.c_int32(), C.c_int32()
        self.ctx._check(self.ctx.L.b200slam_map_device_ptrs(self.h, C.byref(occ), C.byref(op), C.byref(fld),
                                                            C.byref(fp)))
        return occ.value, op.value, fld.value, fp.value


class Context:
    """One b200slam_ctx (one GPU, one stream)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.b200slam_create(C.byref(h), device)
        if rc != OK:
            raise B200SlamError(rc, (self.L.b200slam_last_error(None) or b"").decode())
        self.h = h
        self._pinned = []

    def _check(self, rc: int):
        if rc != OK:
            raise B200SlamError(rc, (self.L.b200slam_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            self.L.b200slam_sync(self.h)
            for p in self._pinned:                 # arrays handed out by pinned_empty die with the context
                self.L.b200slam_host_free(self.h, p)
            self._pinned = []
            self.L.b200slam_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- misc ------------------------------------------------------------------------
    def sync(self):
        self._check(self.L.b200slam_sync(self.h))

    def stream(self) -> int:
        return int(self.L.b200slam_stream(self.h) or 0)

    def set_match_mode(self, mode: int):
        """MATCH_LATENCY (default) or MATCH_THROUGHPUT: tile-shape policy of the lattice kernel."""
        self._check(self.L.b200slam_set_match_mode(self.h, int(mode)))

    # -- scan front end / map points (device resident) ------------------------------------
    def lidar_set(self, angles, range_min: float):
        """angles: the beam angles as the reference accumulates them (main.c:53-57); cos / sin by numpy's
        float32 libm calls would differ from glibc's cosf / sinf, so they are taken from libm via ctypes."""
        a = np.ascontiguousarray(angles, np.float32)
        ca, sa = libm_cosf_sinf(a)
        self._lidar_n = len(a)
        self._check(self.L.b200slam_lidar_set(self.h, _fptr(ca), _fptr(sa), len(a), range_min))

    def scan_read(self, ranges, max_range: int = 24) -> int:
        r = np.ascontiguousarray(ranges, np.float32)
        assert len(r) == self._lidar_n
        n = C.c_int(0)
        self._check(self.L.b200slam_scan_read(self.h, _fptr(r), int(max_range), C.byref(n)))
        self._nbeams = n.value
        return n.value

    def scan_read_async(self, ranges, max_range: int = 24):
        r = np.ascontiguousarray(ranges, np.float32)
        assert len(r) == self._lidar_n
        self._check(self.L.b200slam_scan_read_async(self.h, _fptr(r), int(max_range)))
        self._nbeams = self._lidar_n

    def scan_read_resident_async(self, first_value: int, max_range: int = 24):
        self._check(self.L.b200slam_scan_read_resident_async(self.h, int(first_value), int(max_range)))
        self._nbeams = self._lidar_n

    def csv_ingest(self, text: bytes, max_values: int | None = None) -> np.ndarray:
        """The reference's CSV reader (main.c:22-30) on the GPU: raw text -> float32 values (also resident)."""
        buf = np.frombuffer(text, np.uint8)
        cap = len(buf) // 2 + 1 if max_values is None else int(max_values)
        out = np.empty(max(cap, 1), np.float32)
        n = C.c_int64(0)
        self._check(self.L.b200slam_csv_ingest(self.h, buf.ctypes.data, len(buf), out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    def fastmatch_pair_async(self, map_a: Map, map_b: Map, pose, res_a, res_b):
        self._check(self.L.b200slam_fastmatch_pair_async(self.h, map_a.h, map_b.h, _f3(pose), _f3(res_a), _f3(res_b)))

    def fastmatch_pair_fetch(self):
        pa, pb = (C.c_float * 3)(), (C.c_float * 3)()
        n, bh = C.c_int32(0), C.c_int32(0)
        self._check(self.L.b200slam_fastmatch_pair_fetch(self.h, pa, pb, C.byref(n), C.byref(bh)))
        self._nbeams = n.value
        return np.array(list(pa), np.float32), np.array(list(pb), np.float32), n.value, bh.value

    def scan_step_async(self, ranges, map_a: Map, map_b: Map, pose, res_a, res_b, max_range: int = 24):
        """readAScan + FastMatch + FastMatch2 as one kernel; fetch with fastmatch_pair_fetch."""
        r = np.ascontiguousarray(ranges, np.float32)
        assert len(r) == self._lidar_n
        self._check(self.L.b200slam_scan_step_async(self.h, _fptr(r), int(max_range), map_a.h, map_b.h, _f3(pose), _f3(res_a),
                                                    _f3(res_b)))
        self._nbeams = self._lidar_n

    def scan_step_resident_async(self, first_value: int, map_a: Map, map_b: Map, pose, res_a, res_b, max_range: int = 24):
        self._check(self.L.b200slam_scan_step_resident_async(self.h, int(first_value), int(max_range), map_a.h, map_b.h,
                                                             _f3(pose), _f3(res_a), _f3(res_b)))
        self._nbeams = self._lidar_n

    def scan_chain_begin(self, scan_index: int, pose, prev_pose, map_pose, mini_dt: float, mini_dr: float):
        """Device-resident per-scan loop: set the pose state (prev_pose None: no motion model for the first scan)."""
        self._check(self.L.b200slam_scan_chain_begin(self.h, int(scan_index), _f3(pose), None if prev_pose is None else _f3(prev_pose),
                                                     _f3(map_pose), mini_dt, mini_dr))

    def scan_chain_step_async(self, scan_index: int, first_value: int, map_a: Map, map_b: Map, res_a, res_b, max_range: int = 24,
                              nscans: int = 1):
        """Queue scans [scan_index, scan_index + nscans) as one kernel launch."""
        self._check(self.L.b200slam_scan_chain_step_async(self.h, int(scan_index), int(nscans), int(first_value), int(max_range),
                                                          map_a.h, map_b.h, _f3(res_a), _f3(res_b)))
        self._nbeams = self._lidar_n

    def scan_chain_fetch(self, scan_index: int):
        """-> (pose_a, pose_b, scan size, bestHits_size, stopped)"""
        pa, pb = (C.c_float * 3)(), (C.c_float * 3)()
        n, bh, st = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._check(self.L.b200slam_scan_chain_fetch(self.h, int(scan_index), pa, pb, C.byref(n), C.byref(bh), C.byref(st)))
        self._nbeams = n.value
        return np.array(list(pa), np.float32), np.array(list(pb), np.float32), n.value, bh.value, st.value

    def mappoints_grow_async(self, threshold: float = 1.5):
        self._check(self.L.b200slam_mappoints_grow_async(self.h, threshold))

    def scan_transform(self, pose):
        self._check(self.L.b200slam_scan_transform(self.h, _f3(pose)))

    def scan_download(self, transformed: bool = True):
        cap = 4096 if not hasattr(self, "_lidar_n") else max(4096, self._lidar_n)
        x, y = np.empty(cap, np.float32), np.empty(cap, np.float32)
        tx, ty = np.empty(cap, np.float32), np.empty(cap, np.float32)
        n = C.c_int(0)
        self._check(self.L.b200slam_scan_download(self.h, _fptr(x), _fptr(y), _fptr(tx) if transformed else None,
                                                  _fptr(ty) if transformed else None, C.byref(n)))
        k = n.value
        return (x[:k].copy(), y[:k].copy(), tx[:k].copy(), ty[:k].copy()) if transformed else (x[:k].copy(), y[:k].copy())

    def mappoints_upload(self, x, y, offset: int = 0):
        x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
        self._check(self.L.b200slam_mappoints_upload(self.h, _fptr(x), _fptr(y), len(x), offset))

    def mappoints_from_scan(self):
        self._check(self.L.b200slam_mappoints_from_scan(self.h))

    def mappoints_grow(self, threshold: float = 1.5) -> int:
        n = C.c_int(0)
        self._check(self.L.b200slam_mappoints_grow(self.h, threshold, C.byref(n)))
        return n.value

    def mappoints_download(self):
        n = C.c_int(0)
        self._check(self.L.b200slam_mappoints_download(self.h, None, None, C.byref(n)))
        x, y = np.empty(max(n.value, 1), np.float32), np.empty(max(n.value, 1), np.float32)
        self._check(self.L.b200slam_mappoints_download(self.h, _fptr(x), _fptr(y), C.byref(n)))
        return x[:n.value].copy(), y[:n.value].copy()

    def local_map_extract(self, border: float = 1.0) -> int:
        n = C.c_int(0)
        self._check(self.L.b200slam_local_map_extract(self.h, border, C.byref(n)))
        return n.value

    def local_map_download(self):
        n = C.c_int(0)
        self._check(self.L.b200slam_local_map_download(self.h, None, None, C.byref(n)))
        x, y = np.empty(max(n.value, 1), np.float32), np.empty(max(n.value, 1), np.float32)
        self._check(self.L.b200slam_local_map_download(self.h, _fptr(x), _fptr(y), C.byref(n)))
        return x[:n.value].copy(), y[:n.value].copy()

    def launch_count(self) -> int:
        return int(self.L.b200slam_launch_count(self.h))

    def device_info(self) -> dict:
        name = C.create_string_buffer(64)
        sm, ma, mi = C.c_int32(), C.c_int32(), C.c_int32()
        mem = C.c_size_t()
        self._check(self.L.b200slam_device_info(self.h, name, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"name": name.value.decode(), "sm_count": sm.value, "cc": (ma.value, mi.value),
                "total_mem": mem.value}

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array over cudaHostAlloc'ed memory (lives as long as the context)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._check(self.L.b200slam_host_alloc(self.h, max(n, 1), C.byref(p)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        self._pinned.append(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    # -- EDT -------------------------------------------------------------------------
    def edt(self, occ: np.ndarray, max_dist: float = 10.0, out: np.ndarray | None = None) -> np.ndarray:
        """One-shot host call b200slam_edt: occupancy (host) -> distance field (host)."""
        assert occ.dtype == np.int32 and occ.ndim == 2 and (occ.size == 0 or occ.strides[1] == 4)
        rows, cols = occ.shape
        if out is None:
            out = np.empty((rows, cols), np.float32)
        ostride = occ.strides[0] // 4 if occ.size else cols
        self._check(self.L.b200slam_edt(self.h, occ.ctypes.data, ostride, out.ctypes.data,
                                        out.strides[0] // 4 if out.size else cols, rows, cols, max_dist))
        return out

    def new_map(self, rows: int, cols: int) -> Map:
        return Map(self, rows, cols)

    # -- scoring ---------------------------------------------------------------------
    def scan_upload(self, x, y):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        assert len(x) == len(y)
        self._check(self.L.b200slam_scan_upload(self.h, x.ctypes.data, y.ctypes.data, len(x)))
        self._nbeams = len(x)

    def score_lattice(self, m: Map, pose0, step, n, want_scores=False, want_last_hits=False):
        total = int(n[0]) * int(n[1]) * int(n[2])
        scores = np.empty(total, np.float32) if want_scores else None
        lh = np.zeros(max(self._nbeams, 1), np.float32) if want_last_hits else None
        res = Match()
        self._check(self.L.b200slam_score_lattice(self.h, m.h, _f3(pose0), _f3(step), _i3(n),
                                                  scores.ctypes.data if want_scores else None,
                                                  lh.ctypes.data if want_last_hits else None, C.byref(res)))
        return res, scores, lh

    def score_lattice_rows(self, m: Map, pose0, step, n, row_begin, row_end, allreduce=False) -> Match:
        res = Match()
        self._check(self.L.b200slam_score_lattice_rows(self.h, m.h, _f3(pose0), _f3(step), _i3(n), row_begin,
                                                       row_end, 1 if allreduce else 0, C.byref(res)))
        return res

    def score_lattice_async(self, m: Map, pose0, step, n, row_begin=None, row_end=None, allreduce=False):
        if row_begin is None:
            row_begin, row_end = 0, int(n[0]) * int(n[1])
        self._check(self.L.b200slam_score_lattice_async(self.h, m.h, _f3(pose0), _f3(step), _i3(n), row_begin,
                                                        row_end, int(allreduce)))

    def exchange_collect_async(self):
        self._check(self.L.b200slam_exchange_collect_async(self.h))

    def match_fetch(self) -> Match:
        res = Match()
        self._check(self.L.b200slam_match_fetch(self.h, C.byref(res)))
        return res

    def match_fetch_hits(self, count: int = 2500) -> np.ndarray:
        out = np.zeros(count, np.float32)
        self._check(self.L.b200slam_match_fetch_hits(self.h, out.ctypes.data, count))
        return out

    def score_poses(self, m: Map, poses, ct=None, st=None, index_base=0, want_hits=True):
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 3)
        P = poses.shape[0]
        scores = np.empty(P, np.float32)
        hits = np.empty(P, np.int32) if want_hits else None
        ctp = np.ascontiguousarray(ct, np.float32) if ct is not None else None
        stp = np.ascontiguousarray(st, np.float32) if st is not None else None
        res = Match()
        self._check(self.L.b200slam_score_poses(self.h, m.h, poses.ctypes.data,
                                                ctp.ctypes.data if ctp is not None else None,
                                                stp.ctypes.data if stp is not None else None, P, index_base,
                                                scores.ctypes.data, hits.ctypes.data if want_hits else None,
                                                C.byref(res)))
        return res, scores, hits

    def fastmatch(self, m: Map, pose, res3, hits_buf=None):
        """hits_buf: a persistent float32 array standing in for the global FastMatchParameters.bestHits
        (at least as long as the scan): the call rewrites its leading entries exactly as the reference does."""
        out = (C.c_float * 3)()
        hits = np.zeros(max(self._nbeams, 1), np.float32) if hits_buf is None else hits_buf
        n = C.c_int32(0)
        self._check(self.L.b200slam_fastmatch(self.h, m.h, _f3(pose), _f3(res3), out, hits.ctypes.data,
                                              C.byref(n)))
        return np.array(list(out), np.float32), hits, n.value

    def weights_resample(self, N_alloc: int, beta: float, u0_q32: int, want_weights=True):
        w = np.empty(N_alloc, np.float32) if want_weights else None
        anc = np.empty(N_alloc, np.int32)
        W = C.c_uint64(0)
        kb, kc = C.c_int64(0), C.c_int64(0)
        self._check(self.L.b200slam_weights_resample(self.h, beta, u0_q32, w.ctypes.data if want_weights else None,
                                                     C.byref(W), anc.ctypes.data, C.byref(kb), C.byref(kc)))
        return w, int(W.value), anc[:kc.value], kb.value, kc.value

    # -- device-resident particle set (single GPU) ----------------------------------------
    def particles_upload(self, poses, ct=None, st=None):
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 3)
        ctp = np.ascontiguousarray(ct, np.float32) if ct is not None else None
        stp = np.ascontiguousarray(st, np.float32) if st is not None else None
        self._check(self.L.b200slam_particles_upload(self.h, poses.ctypes.data,
                                                     ctp.ctypes.data if ctp is not None else None,
                                                     stp.ctypes.data if stp is not None else None, poses.shape[0]))
        self._P = poses.shape[0]

    def particles_shard(self, poses, index_base: int, n_global: int, ct=None, st=None):
        """Collective: this rank's slice [index_base, index_base + len(poses)) of a set of n_global particles."""
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 3)
        ctp = np.ascontiguousarray(ct, np.float32) if ct is not None else None
        stp = np.ascontiguousarray(st, np.float32) if st is not None else None
        self._check(self.L.b200slam_particles_shard(self.h, poses.ctypes.data,
                                                    ctp.ctypes.data if ctp is not None else None,
                                                    stp.ctypes.data if stp is not None else None, poses.shape[0],
                                                    index_base, n_global))
        self._P = poses.shape[0]

    def particles_score_async(self, m: Map):
        self._check(self.L.b200slam_particles_score_async(self.h, m.h))

    def particles_resample_async(self, beta: float, u0_q32: int):
        self._check(self.L.b200slam_particles_resample_async(self.h, beta, u0_q32))

    def particles_download(self, want_scores=True, want_weights=True, want_ancestors=True):
        P = self._P
        poses = np.empty((P, 3), np.float32)
        sc = np.empty(P, np.float32) if want_scores else None
        w = np.empty(P, np.float32) if want_weights else None
        a = np.empty(P, np.int32) if want_ancestors else None
        self._check(self.L.b200slam_particles_download(self.h, poses.ctypes.data,
                                                       sc.ctypes.data if want_scores else None,
                                                       w.ctypes.data if want_weights else None,
                                                       a.ctypes.data if want_ancestors else None))
        return poses, sc, w, a

    def pyramid_match(self, maps, pose0, steps, ns):
        Ln = len(maps)
        arr = (C.c_void_p * Ln)(*[m.h for m in maps])
        stp = np.ascontiguousarray(steps, np.float32).reshape(Ln, 3)
        nn = np.ascontiguousarray(ns, np.int32).reshape(Ln, 3)
        res = (Match * Ln)()
        self._check(self.L.b200slam_pyramid_match(self.h, arr, Ln, _f3(pose0), stp.ctypes.data_as(c_float_p),
                                                  nn.ctypes.data_as(c_int_p), res))
        return list(res)

    # -- graphs ----------------------------------------------------------------------
    def graph_begin(self):
        self._check(self.L.b200slam_graph_begin(self.h))

    def graph_end(self):
        g = C.c_void_p()
        self._check(self.L.b200slam_graph_end(self.h, C.byref(g)))
        return g

    def graph_launch(self, g):
        self._check(self.L.b200slam_graph_launch(self.h, g))

    def graph_destroy(self, g):
        self.L.b200slam_graph_destroy(self.h, g)

    # -- timing ----------------------------------------------------------------------
    def event_record(self, slot: int):
        self._check(self.L.b200slam_event_record(self.h, slot))

    def event_wait(self, other: "Context", slot: int):
        self._check(self.L.b200slam_event_wait(self.h, other.h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float(0)
        self._check(self.L.b200slam_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return float(ms.value)

    # -- multi-GPU -------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(UNIQUE_ID_BYTES)
        rc = self.L.b200slam_comm_unique_id(buf)
        if rc != OK:
            raise B200SlamError(rc, (self.L.b200slam_last_error(None) or b"").decode())
        return buf.raw

    def comm_barrier_async(self):
        self._check(self.L.b200slam_comm_barrier_async(self.h))

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        assert len(uid) == UNIQUE_ID_BYTES
        self._check(self.L.b200slam_comm_init(self.h, nranks, rank, uid))
