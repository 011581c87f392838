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
@pytest.mark.parametrize("host_parse", [False, True])
def test_device_resident_replay_is_byte_identical(reference_run, dataset, b200slam, host_parse):
    """b200slam_replay (host/slam_replay.c): the control flow of the reference's main() in C over the ABI,
    with readAScan / Transform / ExtractLocalMap / OccupationalGrid / both EDTs / FastMatch / FastMatch2 / map
    growth all on the device, the dataset parsed on the GPU (rank 4; host_parse: by the
    host's fscanf instead), one synchronisation per scan (FastMatch and FastMatch2 queued as a pair, sizes kept
    on the device).  Same pose trace, same map dump, byte for byte -- which also pins the inline map-growth code
    of main() (main.c:942-948) end to end."""
    d, csv = dataset
    exe = os.path.join(os.path.dirname(b200slam.LIB_PATH), "b200slam_replay")
    assert os.path.exists(exe), "b200slam_replay missing: run __graft_entry__.build()"
    mapout = str(d / "map_dev.csv")
    env = dict(os.environ)
    if host_parse:
        env["B200SLAM_REPLAY_HOST_PARSE"] = "1"
    p = subprocess.run([exe, csv, mapout, str(NSCANS)], capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    gpu_lines = [ln for ln in p.stdout.splitlines() if not ln.startswith("time taken")]
    ref_lines, ref_map = reference_run
    assert len(gpu_lines) == len(ref_lines)
    diff = [i for i, (a, b) in enumerate(zip(ref_lines, gpu_lines)) if a != b]
    assert not diff, f"first differing line {diff[0]}: ref={ref_lines[diff[0]]!r} gpu={gpu_lines[diff[0]]!r}"
    assert open(mapout, "rb").read() == ref_map
