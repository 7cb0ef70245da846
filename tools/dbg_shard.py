"""Per-rank cost of the sharded path, emulated on ONE GPU: N contexts play the ranks one after the other (each rank's
threshold+tables pipeline, then one global phase, then each rank's paint).  usage: dbg_shard.py [T] [N ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from contrack_b200 import Engine, sharded
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10957
Ns = [int(a) for a in sys.argv[2:] if '=' not in a] or [2, 8]
OPTS = [a.split('=') for a in sys.argv[2:] if '=' in a]
lat, lon = bench.grid(); w = bench.reference_weights(lat, lon)
anom = torch.empty((T, bench.H, bench.W), dtype=torch.float32, device='cuda'); bench.synth_fill(anom, 0, T); torch.cuda.synchronize()
for N in Ns:
    engines = [Engine(0) for _ in range(N)]
    for e in engines:
        for k, v in OPTS:
            e.set_option(k, int(v))
    bounds = sharded.shard_bounds(T, N)
    parts = [anom[a:b] for a, b in bounds]
    outs_buf = [torch.empty((b - a, bench.H, bench.W), dtype=torch.int32, device='cuda') for a, b in bounds]
    for it in range(3):
        outs = None
        outs, n, info = sharded.run_contrack_sharded_local_dev(engines, parts, T, w, 160, True, 0, 0.5, 5, True, outs=outs_buf)
    r = round
    print('N=%d features=%d' % (N, n))
    print('  tables_dev ms per rank', [r(x, 2) for x in info['ms_tables']])
    print('  global ms', r(info['ms_global'], 2), ' paint ms per rank', [r(x, 2) for x in info['ms_paint']])
    for e in engines[:2]:
        print('  shard stats', {k: r(v, 3) for k, v in e.stats().items() if k.startswith("ms_") or k in ("chunks", "runs", "comps2d")})
    print('  global stats', {k: r(v, 3) for k, v in info['global_stats'].items() if k.startswith('ms_') or k in ('sweeps', 'kernel_launches', 'comps2d', 'seam_segments', 'labels3d', 'ht_walked', 'seam_events')})
    del outs, outs_buf
    for e in engines:
        for h in (e, getattr(e, '_global', None)):
            if h is not None and h.handle:
                h.lib.ct_destroy(h.handle); h.handle = None
    torch.cuda.empty_cache()
