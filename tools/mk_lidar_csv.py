#!/usr/bin/env python
"""tools/mk_lidar_csv.py OUT.csv NSCANS -- the first NSCANS rows of the synthetic 3480-scan stand-in dataset."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
synth = importlib.import_module("hardware-acceleration-of-lidar-slam_b200.synth")
synth.write_lidar_csv(sys.argv[1], synth.lidar_dataset(3480)[:int(sys.argv[2])])
