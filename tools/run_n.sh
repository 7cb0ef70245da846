#!/bin/bash
# usage: tools/run_n.sh N tag [extra bench args]  -- bench.py on N GPUs of this box (torchrun), output under gpurun_out/
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
fi
echo "rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "features", "breakdown_ms", "shard_ms")})
    print("e2e", d.get("e2e", {}).get("value"))
    for r in d.get("shard_ms_all_ranks", []): print(r)
except Exception as e:
    print("no json:", e)
PY
tail -15 gpurun_out/${TAG}.err
