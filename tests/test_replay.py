"""End-to-end drop-in check.

The UNMODIFIED reference program (Subsystem_1/main_accelerated.c, compiled by oracle/Makefile)
replays a synthetic lidar CSV twice:
  * as shipped (CPU):                    oracle/_ref/ref_replay_accel
  * with libb200slam_dropin.so interposed: oracle/_ref/ref_replay_accel_dropin -- the reference's
    own calls to euclidean_distance_transform{,2} and FastMatch{,2} (main.c:355-356, 902-918)
    bind to the B200 path.
Every `pose = ...` line (main.c:965) and the final map_output.csv (main.c:982-988) must be
byte-identical.  The bundled lidar_dataset.csv is not in the reference tree
(.MISSING_LARGE_BLOBS); synth.lidar_dataset() generates a stand-in in the same format.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
NSCANS = 3480                     # main_accelerated.c:6 (#define row 3480)


def _run(exe, csv, mapout, timeout=900):
    env = dict(os.environ, B200SLAM_REF_DATASET=csv, B200SLAM_REF_MAPOUT=mapout)
    p = subprocess.run([os.path.join(REF, exe)], env=env, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, f"{exe} failed ({p.returncode}): {p.stderr[-2000:]}"
    lines = [ln for ln in p.stdout.splitlines() if not ln.startswith("time taken")]
    return lines, open(mapout, "rb").read()


@pytest.fixture(scope="module")
def dataset(tmp_path_factory, synth):
    d = tmp_path_factory.mktemp("replay")
    csv = str(d / "lidar_dataset.csv")
    synth.write_lidar_csv(csv, synth.lidar_dataset(NSCANS))
    return d, csv


@pytest.fixture(scope="module")
def reference_run(dataset):
    d, csv = dataset
    if not os.path.exists(os.path.join(REF, "ref_replay_accel")):
        pytest.skip("oracle/_ref/ref_replay_accel not built (reference tree absent at build time)")
    return _run("ref_replay_accel", csv, str(d / "map_ref.csv"))


def test_dataset_generator_is_deterministic(synth):
    a = synth.lidar_dataset(5)
    b = synth.lidar_dataset(5)
    assert a.shape == (5, synth.REF_BEAMS) and np.array_equal(a, b)
    assert a.min() > 0.5 and a.max() < 24.0          # all beams survive readAScan's range filter (main.c:78)


def test_reference_tracks_the_synthetic_loop(reference_run):
    """The unmodified reference follows the 3 m-radius loop and closes it (theta = -2 pi)."""
    lines, mapout = reference_run
    poses = [ln for ln in lines if ln.startswith("pose =")]
    assert len(poses) == NSCANS - 1
    x, y, th = (float(v) for v in poses[-1].split("=")[1].split())
    assert abs(x) < 0.1 and abs(y) < 0.1 and abs(th + 2 * np.pi) < 0.02
    assert len(mapout) > 1000


@pytest.mark.gpu
def test_dropin_replay_is_byte_identical(reference_run, dataset, b200slam):
    d, csv = dataset
    exe = os.path.join(REF, "ref_replay_accel_dropin")
    assert os.path.exists(exe), "oracle/_ref/ref_replay_accel_dropin missing: run __graft_entry__.build()"
    ref_lines, ref_map = reference_run
    gpu_lines, gpu_map = _run("ref_replay_accel_dropin", csv, str(d / "map_gpu.csv"))
    assert len(gpu_lines) == len(ref_lines)
    diff = [i for i, (a, b) in enumerate(zip(ref_lines, gpu_lines)) if a != b]
    assert not diff, f"first differing line {diff[0]}: ref={ref_lines[diff[0]]!r} gpu={gpu_lines[diff[0]]!r}"
    assert gpu_map == ref_map


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["device_loop", "host_loop", "host_parse"])
def test_device_resident_replay_is_byte_identical(reference_run, dataset, b200slam, mode):
    """b200slam_replay (host/slam_replay.c): the control flow of the reference's main() in C over the ABI,
    with readAScan / Transform / ExtractLocalMap / OccupationalGrid / both EDTs / FastMatch / FastMatch2 / map
    growth all on the device, the dataset parsed on the GPU (rank 4; host_parse: by the
    host's fscanf instead), one synchronisation per scan (FastMatch and FastMatch2 queued as a pair, sizes kept
    on the device).  Same pose trace, same map dump, byte for byte -- which also pins the inline map-growth code
    of main() (main.c:942-948) end to end.
    device_loop (the default): between map updates the loop runs on the device -- motion model, lattice tables with
    the device's restatement of glibc's cosf / sinf, mini-update test -- with scans queued 16 ahead of their results
    (b200slam_scan_chain_*); host_loop: B200SLAM_REPLAY_NO_CHAIN=1, one kernel and one synchronisation per scan."""
    host_parse = mode == "host_parse"
    d, csv = dataset
    exe = os.path.join(os.path.dirname(b200slam.LIB_PATH), "b200slam_replay")
    assert os.path.exists(exe), "b200slam_replay missing: run __graft_entry__.build()"
    mapout = str(d / "map_dev.csv")
    env = dict(os.environ)
    if host_parse:
        env["B200SLAM_REPLAY_HOST_PARSE"] = "1"
    if mode == "host_loop":
        env["B200SLAM_REPLAY_NO_CHAIN"] = "1"
    p = subprocess.run([exe, csv, mapout, str(NSCANS)], capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    gpu_lines = [ln for ln in p.stdout.splitlines() if not ln.startswith("time taken")]
    ref_lines, ref_map = reference_run
    assert len(gpu_lines) == len(ref_lines)
    diff = [i for i, (a, b) in enumerate(zip(ref_lines, gpu_lines)) if a != b]
    assert not diff, f"first differing line {diff[0]}: ref={ref_lines[diff[0]]!r} gpu={gpu_lines[diff[0]]!r}"
    assert open(mapout, "rb").read() == ref_map
    if mode == "device_loop":
        assert "device-side loop unavailable" not in p.stderr, p.stderr[-1000:]
        ran = [ln for ln in p.stderr.splitlines() if "device-side chains" in ln]
        assert ran and int(ran[0].split(":")[1].split()[0]) > NSCANS * 0.9, p.stderr[-1000:]      # ~98 % of the scans


def test_device_trig_restatement_equals_libm_on_every_float(tmp_path):
    """csrc/trig.cuh restates glibc's sinf / cosf (double-precision polynomial after a fixed-point range reduction)
    so that the device-resident scan loop can build its lattice tables without the host (main.c:434-435 take them
    from libm).  Host build of the same source against the running libm, EVERY float of |y| <= 16 (the range the
    device serves), both functions, both signs: bit-identical."""
    import shutil
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = str(tmp_path / "trig_check")
    subprocess.run([gxx, "-O2", "-ffp-contract=off", "-mfma", "-pthread", "-o", exe, os.path.join(ROOT, "tools", "trig_check.cpp"), "-lm"],
                   check=True)
    p = subprocess.run([exe, "16", str(min(32, os.cpu_count() or 4))], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "mismatches 0 0" in p.stdout, p.stdout[-500:]


@pytest.mark.gpu
def test_scan_chain_matches_the_host_driven_steps(dataset, synth, b200slam):
    """b200slam_scan_chain_*: scans queued ahead on the device (pose state, motion model, tables, mini-update test
    in the kernels) against the same scans driven one by one from the host (b200slam_scan_step_resident_async with
    the host's cosf / sinf): every pose bit-identical, the chain stops exactly where the host's test fires and the
    scan left on the device is that scan's (the kernels queued behind it did not run)."""
    nscan = 120
    d, csv = dataset
    with open(csv, "rb") as f:
        txt = b"".join(f.readline() for _ in range(nscan))
    n = synth.REF_BEAMS
    a = np.empty(n, np.float32)
    v = np.float32(-2.351831)
    for i in range(n):                                   # accumulated in float like main.c:47-57
        a[i] = v
        v = np.float32(v + np.float32(0.004363))
    res_a = np.array([0.05, 0.05, 0.008727], np.float32)
    res_b = np.array([0.025, 0.025, 0.004363], np.float32)
    dt, dr = np.float32(0.3), np.float32(0.0872665)

    def run(chain, dt=dt, dr=dr, upto=60, window=None):
        c = b200slam.Context(0)
        try:
            c.lidar_set(a, 0.023)
            assert len(c.csv_ingest(txt, nscan * n)) == nscan * n
            fine = c.new_map(400, 400)
            pose = np.zeros(3, np.float32)
            c.scan_read_resident_async(0)
            c.scan_transform(pose)
            c.mappoints_from_scan()
            c.local_map_extract(1.0)
            fine.rasterise_local(0.1)
            fine.edt(10.0)
            poses, stops = [pose.copy()], []
            map_pose = pose.copy()
            k = 1
            last = min(nscan, upto)
            if chain:
                c.scan_chain_begin(1, pose, None, map_pose, float(dt), float(dr))
                q = 1
                while k < last:
                    # launches of 1, 2, 3, ... scans, queued up to `window` scans ahead of the result being read
                    # (the ring of results holds 64: with a window the slots are reused)
                    while q < last and (window is None or q < k + window):
                        nb = min(1 + (q % 7), last - q)
                        c.scan_chain_step_async(q, q * n, fine, fine, res_a, res_b, nscans=nb)
                        q += nb
                    pa, pb, sz, bh, st = c.scan_chain_fetch(k)
                    poses.append(pb.copy())
                    k += 1
                    if st:
                        stops.append(k - 1)
                        break
            else:
                while k < last:
                    guess = pose.copy() if k == 1 else (pose + (pose - poses[-2]).astype(np.float32)).astype(np.float32)
                    c.scan_step_resident_async(k * n, fine, fine, guess, res_a, res_b)
                    pa, pb, sz, bh = c.fastmatch_pair_fetch()
                    pose = pb.copy()
                    poses.append(pose.copy())
                    k += 1
                    dd = np.abs((pose - map_pose).astype(np.float32))
                    if dd[0] > dt or dd[1] > dt or dd[2] > dr:
                        stops.append(k - 1)
                        break
            c.sync()
            sx = c.scan_download(transformed=False)
            return poses, stops, sx
        finally:
            c.close()

    # thresholds nothing reaches: one chain over all 119 scans, results through a ring slot twice, 24 scans ahead at most
    big = np.float32(1e6)
    ph2, sh2, _ = run(False, big, big, upto=nscan)
    pc2, sc2, _ = run(True, big, big, upto=nscan, window=24)
    assert len(ph2) == nscan and len(pc2) == nscan and sh2 == sc2 == []
    assert all(np.array_equal(p.view(np.uint32), q.view(np.uint32)) for p, q in zip(ph2, pc2))
    ph, sh, xh = run(False)
    pc, sc, xc = run(True)
    assert len(ph) == len(pc) and sh == sc, (len(ph), len(pc), sh, sc)
    assert len(ph) > 10
    assert all(np.array_equal(p.view(np.uint32), q.view(np.uint32)) for p, q in zip(ph, pc))
    assert np.array_equal(np.asarray(xh[0]).view(np.uint32), np.asarray(xc[0]).view(np.uint32))      # the scan left on the device


@pytest.mark.gpu
def test_scan_chain_out_of_order_scan_does_not_run(dataset, synth, b200slam):
    """A chained scan only runs when it is the one the device state expects: queued with a gap it returns at once, the
    state stays put, and fetching its result is an error (after the bounded wait), not a hang or a stale slot."""
    d, csv = dataset
    n = synth.REF_BEAMS
    with open(csv, "rb") as f:
        txt = b"".join(f.readline() for _ in range(12))
    a = np.empty(n, np.float32)
    v = np.float32(-2.351831)
    for i in range(n):
        a[i] = v
        v = np.float32(v + np.float32(0.004363))
    res_a = np.array([0.05, 0.05, 0.008727], np.float32)
    res_b = np.array([0.025, 0.025, 0.004363], np.float32)
    c = b200slam.Context(0)
    try:
        c.lidar_set(a, 0.023)
        c.csv_ingest(txt, 12 * n)
        fine = c.new_map(400, 400)
        pose = np.zeros(3, np.float32)
        c.scan_read_resident_async(0)
        c.scan_transform(pose)
        c.mappoints_from_scan()
        c.local_map_extract(1.0)
        fine.rasterise_local(0.1)
        fine.edt(10.0)
        c.scan_chain_begin(3, pose, None, pose, 0.3, 0.0872665)
        c.scan_chain_step_async(5, 5 * n, fine, fine, res_a, res_b)          # the chain expects scan 3
        with pytest.raises(b200slam.B200SlamError):
            c.scan_chain_fetch(5)
        c.scan_chain_step_async(3, 3 * n, fine, fine, res_a, res_b, nscans=2)
        _, p3, _, _, st3 = c.scan_chain_fetch(3)
        _, p4, _, _, st4 = c.scan_chain_fetch(4)
        assert st3 == 0 and st4 == 0 and np.all(np.isfinite(p3)) and np.all(np.isfinite(p4))
        with pytest.raises(b200slam.B200SlamError):                          # more scans than the CSV holds
            c.scan_chain_step_async(5, 5 * n, fine, fine, res_a, res_b, nscans=8)
    finally:
        c.close()
