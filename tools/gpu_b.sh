#!/bin/bash
# round-2 call B: staged, every step under a short timeout (a hung kernel must not hold the box)
mkdir -p gpurun_out
timeout 120 python tools/dbg_fast.py plane > gpurun_out/r2b_plane.log 2>&1; rc=$?; echo "plane rc=$rc"; tail -8 gpurun_out/r2b_plane.log
[ $rc -ne 0 ] && exit 1
timeout 120 python tools/dbg_fast.py global > gpurun_out/r2b_global.log 2>&1; rc=$?; echo "global rc=$rc"; tail -8 gpurun_out/r2b_global.log
[ $rc -ne 0 ] && exit 1
timeout 400 python -m pytest tests/test_gpu_fastpath.py -x -q > gpurun_out/r2b_fast.log 2>&1; rc=$?; echo "fastpath rc=$rc"
tail -25 gpurun_out/r2b_fast.log
[ $rc -eq 124 ] && exit 1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -15 gpurun_out/r2b_pytest.log
[ $rc -eq 124 ] && exit 1
timeout 400 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2b_bench_n1.json
