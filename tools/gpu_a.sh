#!/bin/bash
# round-2 call A: GPU tests, sanitizer, baseline bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" 
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/r2a_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r2a_sanitizer_$tool.log
done
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"
python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2a_bench_config3.json 2> gpurun_out/r2a_bench_config3.err; echo "bench3 rc=$?"
tail -c 600 gpurun_out/r2a_pytest.log
