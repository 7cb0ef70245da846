#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fastpath.py -x -q > gpurun_out/r2c_fast.log 2>&1; rc=$?; echo "fastpath rc=$rc"
tail -5 gpurun_out/r2c_fast.log
[ $rc -eq 124 ] && exit 1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -8 gpurun_out/r2c_pytest.log
[ $rc -eq 124 ] && exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2c_ncu_bench.log 2>&1; echo "ncu rc=$?"
