#!/bin/bash
# ncu evidence of round 2 (run on the GPU box through gpurun; every command first runs once WITHOUT ncu).
# Reports land in gpurun_out/; summaries are extracted with tools/ncu_summary.py and copied to profiles/.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-extra"
NCU="ncu --clock-control none"
run() { echo "== $*"; "$@"; echo "   exit $?"; }
{
run $B --workload config3 > gpurun_out/r02_plain_c3.log 2>&1 &&
run $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches_config3.csv $B --workload config3 > gpurun_out/r02_ncu_l3.log 2>&1
run $NCU --set full --import-source on -k regex:lattice_kernel -s 8 -c 1 -o gpurun_out/r02_full_c3_lattice -f $B --workload config3 > gpurun_out/r02_ncu_f3l.log 2>&1
run $NCU --set full --import-source on -k regex:edt_tma -s 8 -c 1 -o gpurun_out/r02_full_c3_edt -f $B --workload config3 > gpurun_out/r02_ncu_f3e.log 2>&1
run $B --workload config1 > gpurun_out/r02_plain_c1.log 2>&1 &&
run $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches_config1.csv $B --workload config1 > gpurun_out/r02_ncu_l1.log 2>&1
run $NCU --set full --import-source on -k regex:"edt_tma|lattice_kernel" -s 40 -c 2 -o gpurun_out/r02_full_c1 -f $B --workload config1 > gpurun_out/r02_ncu_f1.log 2>&1
run $NCU --set full --import-source on -k regex:poses_kernel -s 4 -c 1 -o gpurun_out/r02_full_c2_poses -f python bench.py --steps 4 --warmup 3 --workload config2 > gpurun_out/r02_ncu_f2.log 2>&1
python - <<'PY'
import importlib, os, sys
sys.path.insert(0, ".")
synth = importlib.import_module("hardware-acceleration-of-lidar-slam_b200.synth")
synth.write_lidar_csv("gpurun_out/_lidar60.csv", synth.lidar_dataset(60))
PY
R=hardware-acceleration-of-lidar-slam_b200/b200slam_replay
run $R gpurun_out/_lidar60.csv gpurun_out/_map60.csv 60 > /dev/null 2> gpurun_out/r02_plain_replay.log &&
run $NCU --set full --import-source on -k regex:fastmatch_kernel -s 20 -c 2 -o gpurun_out/r02_full_fastmatch -f $R gpurun_out/_lidar60.csv gpurun_out/_map60.csv 60 > gpurun_out/r02_ncu_fm.log 2>&1
run $NCU --metrics gpu__time_duration.sum -s 100 -c 60 --csv --log-file gpurun_out/r02_launches_replay.csv $R gpurun_out/_lidar60.csv gpurun_out/_map60.csv 60 > /dev/null 2>&1
rm -f gpurun_out/_lidar60.csv gpurun_out/_map60.csv
} 2>&1 | tee gpurun_out/r02_profile.log | grep -v "^==PROF==" | tail -40
ls -la gpurun_out/r02_*
