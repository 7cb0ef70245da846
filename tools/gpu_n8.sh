#!/bin/bash
# 8 GPUs: default bench (T=10957, parity + e2e) and config 5 (T=43828, exact sharded quantile)
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_n$N.err
if [ "$N" = "8" ] && [ -z "$2" ]; then
  timeout 600 $TR bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_bench_n${N}_config5.json 2> gpurun_out/r2_bench_n${N}_config5.err; echo "bench5 rc=$?"
  tail -c 400 gpurun_out/r2_bench_n${N}_config5.err
fi
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n$N*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],3), d['gpu_launches'], d.get('parity'), 'e2e', d.get('e2e',{}).get('value'), d.get('e2e',{}).get('checksum'), d['config'].get('threshold_value'), d['clocks'])
        for r in d.get('shard_ms_all_ranks', [])[:3]: print(r)
    except Exception as e: print(f,'ERR',e)
PY
