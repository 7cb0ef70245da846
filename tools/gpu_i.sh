#!/bin/bash
# 2 GPUs: NCCL test of the sharded call + a short bench (parity block included)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py -q -k "torchrun" > gpurun_out/r2i_nccl_test.log 2>&1; rc=$?; echo "nccl test rc=$rc"
tail -12 gpurun_out/r2i_nccl_test.log
[ $rc -eq 124 ] && exit 1
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/r2i_nccl.%p.log timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2i_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2i_bench_n2.json').read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step'],3), d['gpu_launches'], d.get('parity'))
    for r in d.get('shard_ms_all_ranks', []): print(r)
except Exception as e: print('ERR', e)
PY
grep -h "Init COMPLETE\|nranks" gpurun_out/r2i_nccl.*.log | head -6
