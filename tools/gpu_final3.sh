#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -3 gpurun_out/r2k_pytest.log
[ $rc -ne 0 ] && exit 1
timeout 120 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r2k_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -2 gpurun_out/r2k_sanitizer_racecheck.log
