#!/bin/bash
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_n$N.err
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -k torchrun 2>&1 | tail -3
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), d['gpu_launches'], d['parity']['checksum'], d['parity']['bit_exact_vs_oracle'], 'e2e', d['e2e']['value'], d['e2e']['checksum'])
for r in d.get('shard_ms_all_ranks', []): print(r)
PY
