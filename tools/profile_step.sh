#!/bin/bash
# Run on the GPU box (gpurun): launch list of one full-size step and one `--set full` capture of the cube-sized kernels.
# Outputs land in gpurun_out/; summaries are copied into profiles/ by tools/summarise_ncu.py on the build box.
TAG=${1:-r1c}
set -x
mkdir -p gpurun_out
# every launch of full-size steps (bench.py: 1 warm-up + 1 timed step; shares are computed from the LAST step's launches)
KERNELS='regex:k_threshold|k_zero_fill|k_paint|k_plane|k_global|k_scan|k_compact|k_extract|k_iota|k_ccl|k_comp_|k_seam|k_cls|k_pairs|k_class|k_seg_|k_apply|k_run_values|k_row_stats'
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KERNELS" --csv --log-file gpurun_out/${TAG}_launches_T10957.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/${TAG}_launches_T10957.bench.log 2>&1
# the same for a shard-sized cube (1370 planes = one rank of an 8-GPU run): the plane kernel builds the tables there
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KERNELS" --csv --log-file gpurun_out/${TAG}_launches_T1370.csv \
    python bench.py --T 1370 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/${TAG}_launches_T1370.bench.log 2>&1
# full capture of the cube-sized kernels (T=1461: 6 GB in / 6 GB out, far beyond the 126 MB L2; 4 years of days so that
# calc_clim has 4 members per group); chunks=1 so that the threshold is one launch
ncu --set full --clock-control none --import-source on \
    -k regex:'k_threshold|k_zero_fill|k_paint|k_anom|k_group_mean|k_clim_smooth|k_compact_runs|k_flag_count|k_plane_tables|k_global_phase' -c 44 -f \
    -o gpurun_out/${TAG}_full python tools/prof_target.py 1461 chunks=1 > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out
