#!/bin/bash
# peer-window exchange: sharded tests (in-process group + 2-GPU NCCL/IPC), then bench at N = $1 with both transports
mkdir -p gpurun_out
N=${1:-2}
if [ -z "$2" ]; then timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/p2p_pytest.log 2>&1; echo "pytest rc=$?"; fi
tail -5 gpurun_out/p2p_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for P in 1 0; do
  timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu --opt p2p=$P > gpurun_out/p2p_bench_n${N}_p$P.json 2> gpurun_out/p2p_bench_n${N}_p$P.err; echo "bench p2p=$P rc=$?"
  tail -c 300 gpurun_out/p2p_bench_n${N}_p$P.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/p2p_bench_n$N*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],3), d['gpu_launches'], d['parity']['checksum'], d['parity']['bit_exact_vs_oracle'])
        for r in d.get('shard_ms_all_ranks', [])[:8]: print({k: r.get(k) for k in ('ms_threshold','ms_plane_kernel','ms_exchange','ms_global_kernel','ms_host_tables','ms_paint','ms_total','p2p','kernel_launches')})
    except Exception as e: print(f,'ERR',e)
PY
