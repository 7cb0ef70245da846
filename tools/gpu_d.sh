#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fastpath.py -x -q > gpurun_out/r2d_fast.log 2>&1; rc=$?; echo "fastpath rc=$rc"
tail -5 gpurun_out/r2d_fast.log
[ $rc -eq 124 ] && exit 1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -8 gpurun_out/r2d_pytest.log
[ $rc -eq 124 ] && exit 1
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r2d_bench_c1.json 2> gpurun_out/r2d_bench_c1.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity --opt fast_chunks=4 > gpurun_out/r2d_bench_c4.json 2> gpurun_out/r2d_bench_c4.err; echo "bench4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:k_(plane|global|threshold|paint|zero|comp_values|apply)' -c 60 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2d_ncu_bench.log 2>&1; echo "ncu rc=$?"
