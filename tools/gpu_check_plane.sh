#!/bin/bash
# after the race-free union-find flatten: GPU tests + racecheck + shard-sized bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -3 gpurun_out/r2_pytest.log
[ $rc -eq 124 ] && exit 1
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -2 gpurun_out/r2_sanitizer_racecheck.log
timeout 300 python bench.py --T 1370 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_T1370.json 2> gpurun_out/r2_bench_T1370.err; echo "T1370 rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/r2_bench_T1370.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],3), d.get('gpu_launches'), (d.get('parity') or {}).get('bit_exact_vs_oracle'), d['breakdown_ms'])
    except Exception as e: print(f,'ERR',e)
PY
