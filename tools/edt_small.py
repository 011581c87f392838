import importlib, numpy as np, sys
sys.path.insert(0, "/root/repo")
mod = importlib.import_module("hardware-acceleration-of-lidar-slam_b200")
synth = importlib.import_module("hardware-acceleration-of-lidar-slam_b200.synth")
from oracle.pyoracle import Oracle
orc = Oracle()
with mod.Context(0) as ctx:
    for (r, c) in [(64, 64), (200, 150), (400, 400), (1000, 777)]:
        occ = synth.grid_bernoulli(r, c, 0.01, 5)
        out = ctx.edt(occ)
        ref = orc.edt(occ)
        print(r, c, "equal" if np.array_equal(out.view(np.uint32), ref.view(np.uint32)) else "DIFF %d" % (out != ref).sum(), flush=True)
