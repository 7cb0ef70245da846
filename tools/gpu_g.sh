#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2g_sharded.log 2>&1; rc=$?; echo "sharded rc=$rc"
tail -15 gpurun_out/r2g_sharded.log
[ $rc -eq 124 ] && exit 1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -8 gpurun_out/r2g_pytest.log
[ $rc -eq 124 ] && exit 1
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity --opt coop_global=0 > gpurun_out/r2g_bench_nocoop.json 2> gpurun_out/r2g_bench_nocoop.err; echo "bench nocoop rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/r2g_bench.json','gpurun_out/r2g_bench_nocoop.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],3), d['gpu_launches'], {k:round(v,2) for k,v in d['breakdown_ms'].items()}, round(d['roofline_path']['frac'],4))
    except Exception as e: print(f,'ERR',e)
PY
