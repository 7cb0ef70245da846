#!/bin/bash
# single-GPU pass: GPU tests, compute-sanitizer, bench lines (default, config 3, reference arm)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -5 gpurun_out/r2_pytest.log
[ $rc -eq 124 ] && exit 1
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/r2_sanitizer_$tool.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r2_bench_config3.json 2> gpurun_out/r2_bench_config3.err; echo "bench3 rc=$?"
timeout 300 python bench.py --T 1370 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_T1370.json 2> gpurun_out/r2_bench_T1370.err; echo "T1370 rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/r2_bench_n1.json','gpurun_out/r2_bench_config3.json','gpurun_out/r2_bench_T1370.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],3), d.get('gpu_launches'), (d.get('parity') or {}).get('checksum'), (d.get('parity') or {}).get('bit_exact_vs_oracle'), 'e2e', d.get('e2e',{}).get('value'), d.get('e2e',{}).get('checksum'), d.get('roofline',{}).get('frac'), d.get('roofline_path',{}).get('frac'), d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
