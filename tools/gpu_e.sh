#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name 'regex:k_plane_tables' --launch-skip 1 -c 1 -o gpurun_out/r2e_plane -f python tools/prof_target.py 2922 > gpurun_out/r2e_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2e_ncu.log
