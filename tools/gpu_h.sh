#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2h_sharded.log 2>&1; rc=$?; echo "sharded rc=$rc"
tail -15 gpurun_out/r2h_sharded.log
[ $rc -eq 124 ] && exit 1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -8 gpurun_out/r2h_pytest.log
