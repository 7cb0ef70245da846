#!/bin/bash
# N-GPU bench variants of the fill scheduling (no e2e / cpu legs)
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; shift; timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/tail_n${N}_$tag.json 2> gpurun_out/tail_n${N}_$tag.err; echo "$tag rc=$?"; }
run tail40 --opt fill_tail=40
run tail60 --opt fill_tail=60
run ctas1 --opt fill_ctas=1
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/tail_n$N*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],3), d['gpu_launches'], d['parity']['checksum'], d['parity']['bit_exact_vs_oracle'])
        for r in d.get('shard_ms_all_ranks', [])[:3]: print({k: r.get(k) for k in ('ms_threshold','ms_zero_fill','ms_plane_kernel','ms_exchange','ms_global_kernel','ms_host_tables','ms_paint','ms_total')})
    except Exception as e: print(f,'ERR',e)
PY
