"""CPU: the C-ABI shared library loads, exports every symbol include/b200slam.h declares,
its pure-host helpers work, and it refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(b200slam):
    L = b200slam.load_library()
    declared = b200slam.declared_symbols()
    assert len(declared) >= 35
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, f"declared in include/b200slam.h but not exported: {missing}"
    out = subprocess.run(["nm", "-D", "--defined-only", b200slam.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(declared) <= exported
    assert L.b200slam_abi_version() == 1


def test_library_is_sm100a_only(b200slam):
    out = subprocess.run(["cuobjdump", "-lelf", b200slam.LIB_PATH], capture_output=True, text=True).stdout
    archs = {ln.split(".")[-2] for ln in out.splitlines() if "sm_" in ln}
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(b200slam):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b200slam.B200SlamError) as e:
        b200slam.Context(0)
    assert e.value.code == b200slam.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_link_oracle(b200slam):
    out = subprocess.run(["ldd", b200slam.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", b200slam.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in syms


def test_shard_range_partitions(b200slam):
    for total in (0, 1, 7, 64, 32768, 4194304 + 3):
        for nranks in (1, 2, 3, 8):
            edges = [b200slam.shard_range(total, nranks, r) for r in range(nranks)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(nranks - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1


def test_packed_key_orders_like_score_then_index(b200slam):
    rng = np.random.default_rng(0)
    scores = np.concatenate([rng.random(200).astype(np.float32) * 1000, np.zeros(3, np.float32),
                             np.full(4, 17.25, np.float32)])
    idx = rng.permutation(len(scores)).astype(np.uint32)
    keys = np.array([b200slam.pack_key(float(s), int(i)) for s, i in zip(scores, idx)], np.uint64)
    best = b200slam.merge_keys(keys)
    s, i = b200slam.unpack_key(best)
    order = np.lexsort((idx, scores))
    assert s == scores[order[0]] and i == idx[order[0]]
    for k, (sc, ix) in zip(keys, zip(scores, idx)):
        assert b200slam.unpack_key(int(k)) == (float(sc), int(ix))


def test_lattice_value_matches_reference_triplet(b200slam):
    # n == 3 must reproduce {p - s, p, p + s} (Subsystem_1/main.c:424-426) bit for bit
    rng = np.random.default_rng(1)
    for _ in range(200):
        p, s = np.float32(rng.normal() * 3), np.float32(rng.random() * 0.1)
        want = [np.float32(p - s), p, np.float32(p + s)]
        got = [np.float32(b200slam.lattice_value(float(p), float(s), k, 3)) for k in range(3)]
        assert [w.tobytes() for w in want] == [g.tobytes() for g in got]


def test_scoring_kernels_contain_no_fused_multiply_add(b200slam):
    """Bit-exactness rests on every float product and sum being rounded separately (the reference is built
    without contraction).  nvcc -fmad=false and __fmul_rn / __fadd_rn guarantee that for
    scalar code, but ptxas was caught contracting mul.rn.f32x2 + add.rn.f32x2 into FFMA2 -- so the SASS of the
    kernels that compute cell indices and scores is checked for any FFMA / FFMA2."""
    import re
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([cuobjdump, "-sass", b200slam.LIB_PATH], capture_output=True, text=True, check=True).stdout
    # (raster_scatter_kernel is not on the list: its IEEE division, __fdiv_rn, is itself built from FFMAs)
    watched = ("lattice_kernel", "poses_kernel", "fastmatch_kernel", "scan_read_kernel", "scan_transform_kernel", "local_map_kernel")
    current, seen, bad = None, set(), []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            current = next((w for w in watched if w in m.group(1)), None)
            if current:
                seen.add(current)
            continue
        if current and re.search(r"\bFFMA2?\b", line):
            bad.append((current, line.strip()[:80]))
    assert seen == set(watched), f"kernels not found in the SASS dump: {set(watched) - seen}"
    assert not bad, f"fused multiply-add in a bit-exact kernel: {bad[:3]}"
