#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_extras.py -x -q > gpurun_out/r2k_new.log 2>&1; rc=$?; echo "new tests rc=$rc"
tail -12 gpurun_out/r2k_new.log
[ $rc -eq 124 ] && exit 1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -6 gpurun_out/r2k_pytest.log
[ $rc -eq 124 ] && exit 1
timeout 300 python bench.py --T 1370 --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity --opt plane_kernel=1 > gpurun_out/r2k_bench_T1370_plane.json 2> gpurun_out/r2k_bench_T1370_plane.err; echo "bench rc=$?"
timeout 300 python bench.py --T 2739 --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity --opt plane_kernel=1 > gpurun_out/r2k_bench_T2739_plane.json 2> gpurun_out/r2k_bench_T2739_plane.err; echo "bench rc=$?"
timeout 300 python bench.py --T 1370 --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity --opt plane_kernel=0 > gpurun_out/r2k_bench_T1370_classic.json 2> gpurun_out/r2k_bench_T1370_classic.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],3), d['gpu_launches'], d['breakdown_ms'])
    except Exception as e: print(f,'ERR',e)
PY
