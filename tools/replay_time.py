#!/usr/bin/env python
"""tools/replay_time.py -- wall time of the unmodified reference program replaying the synthetic
3480-scan dataset: as shipped (CPU) vs with libb200slam_dropin.so interposed (GPU hot path)."""
import importlib, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
synth = importlib.import_module("hardware-acceleration-of-lidar-slam_b200.synth")
d = tempfile.mkdtemp()
csv = os.path.join(d, "lidar.csv")
synth.write_lidar_csv(csv, synth.lidar_dataset(3480))
for exe in ("ref_replay_accel", "ref_replay_accel_dropin", "ref_replay_main"):
    if exe == "ref_replay_main" and "--main" not in sys.argv:
        continue
    env = dict(os.environ, B200SLAM_REF_DATASET=csv, B200SLAM_REF_MAPOUT=os.path.join(d, exe + ".map"))
    t0 = time.perf_counter()
    p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", exe)], env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    tt = [ln for ln in p.stdout.splitlines() if ln.startswith("time taken")]
    print(f"{exe:28s} rc={p.returncode} wall={dt:7.3f} s  {tt[-1] if tt else ''}  last: {p.stdout.splitlines()[-2] if p.stdout else ''}")
exe = os.path.join(ROOT, "hardware-acceleration-of-lidar-slam_b200", "b200slam_replay")
t0 = time.perf_counter()
p = subprocess.run([exe, csv, os.path.join(d, "dev.map"), "3480"], capture_output=True, text=True)
dt = time.perf_counter() - t0
tt = [ln for ln in p.stdout.splitlines() if ln.startswith("time taken")]
print(f"{'b200slam_replay (device resident)':28s} rc={p.returncode} wall={dt:7.3f} s  {tt[-1] if tt else ''}  last: {p.stdout.splitlines()[-2] if p.stdout else ''}  {p.stderr.strip()[-700:]}")
same = open(os.path.join(d, "dev.map"), "rb").read() == open(os.path.join(d, "ref_replay_accel.map"), "rb").read()
print("device-resident map dump identical to the reference's:", same)
