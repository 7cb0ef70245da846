#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --T 2740 --steps 10 --warmup 3 --no-cpu --no-parity --no-e2e > gpurun_out/r2_n2_T2740.json 2> gpurun_out/r2_n2_T2740.err; echo "rc=$?"
timeout 500 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "rc=$?"
tail -c 300 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
for f in ['gpurun_out/r2_n2_T2740.json','gpurun_out/r2_bench_n2.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],3), d['gpu_launches'], d.get('parity'), 'e2e', d.get('e2e',{}).get('value'), d.get('e2e',{}).get('checksum'))
        for r in d.get('shard_ms_all_ranks', []): print(r)
    except Exception as e: print(f,'ERR',e)
PY
