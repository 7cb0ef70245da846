#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -8 gpurun_out/r2p_pytest.log
