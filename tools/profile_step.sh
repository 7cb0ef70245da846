#!/bin/bash
# Run on the GPU box (gpurun): launch list of one full-size step and one `--set full` capture of the cube-sized kernels.
# Outputs land in gpurun_out/; summaries are copied into profiles/ by tools/summarise_ncu.py on the build box.
set -x
mkdir -p gpurun_out
free -g | head -2; nproc
# every launch of one step at the full size (1 warm-up step skipped by position: counted below from the csv)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_T10957.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/launches_T10957.bench.log 2>&1
# full capture of the three cube-sized kernels (T=512: 2.1 GB in / 2.1 GB out, far beyond the 126 MB L2)
ncu --set full --clock-control none --import-source on \
    -k regex:'k_threshold|k_zero_fill|k_paint' -s 3 -c 3 -f -o gpurun_out/prof_T512 \
    python bench.py --T 512 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/prof_T512.bench.log 2>&1
ls -la gpurun_out
