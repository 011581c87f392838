#!/bin/bash
# A/B of two builds of libb200slam.so on the GPU box (development aid).  usage: tools/ab_r2.sh
mkdir -p gpurun_out
{
echo "== parity of the changed kernels"
timeout 900 python -m pytest tests/test_edt_gpu.py tests/test_score_gpu.py tests/test_parity_holes_gpu.py -x -q 2>&1 | tail -5
echo "== EDT chunk heights"
timeout 300 python tools/kbench.py edtcb 2>&1
for v in 5 6 7; do echo "-- RR variant $v"; B200SLAM_RR_VARIANT=$v timeout 300 python tools/kbench.py lattice 2>&1 | grep config3; done
} > gpurun_out/r2_ab3.log 2>&1
tail -70 gpurun_out/r2_ab3.log
