"""The replicated global phase of a sharded run at any cube size on ONE GPU: the rank tables are built one time shard after
the other (each shard's cube slice is generated, reduced to tables and freed), then merged and the global phase is timed.
usage: dbg_global.py T N [overlap persistence threshold]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from contrack_b200 import Engine, sharded
T, N = int(sys.argv[1]), int(sys.argv[2])
ov = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
pers = int(sys.argv[4]) if len(sys.argv) > 4 else 5
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 160.0
H, W = bench.H, bench.W
lat, lon = bench.grid(); w = bench.reference_weights(lat, lon)
bounds = sharded.shard_bounds(T, N)
exports, counts, edge = [], [], None
for r, (a, b) in enumerate(bounds):
    x = torch.empty((b - a, H, W), dtype=torch.float32, device='cuda'); bench.synth_fill(x, a, T)
    e = Engine(0)
    e.set_option('overlap_zero', 0)
    sh = sharded.Shard(e, x, a, r > 0, out=torch.empty((1,), dtype=torch.int32, device='cuda'))
    new_edge = sh.begin(w, thr, True, 0)
    if r > 0:
        sh.import_halo(edge)
    edge = new_edge
    # no flag cube here: tables only
    c8 = (sharded.C.c_long * 8)(); nb = sharded.C.c_long(0)
    sharded._lib.check(e.lib.ct_shard_tables_dev(e.handle, None, sh._stream(), c8, sharded.C.byref(nb)))
    k = np.array(list(c8), np.int64); k[0] = a - (1 if r > 0 else 0)
    buf = torch.empty(int(nb.value), dtype=torch.uint8, device='cuda')
    sharded._lib.check(e.lib.ct_shard_export_tables(e.handle, sharded.C.c_void_p(buf.data_ptr()), int(buf.numel()), sh._stream()))
    torch.cuda.synchronize()
    exports.append(buf); counts.append(k)
    e.lib.ct_destroy(e.handle); e.handle = None
    del x, sh
    torch.cuda.empty_cache()
    print('rank', r, 'comps', int(k[1]), 'pairs', int(k[3]), 'bytes', int(nb.value), flush=True)
stride = (max(b.numel() for b in exports) + 255) // 256 * 256
gathered = torch.zeros(stride * N, dtype=torch.uint8, device='cuda')
for r, b in enumerate(exports):
    gathered[r * stride:r * stride + b.numel()] = b
torch.cuda.synchronize()
g = Engine(0)
for it in range(3):
    t0 = time.perf_counter()
    n = sharded.global_phase(g, np.stack(counts), gathered, stride, T, H, W, w, ov, pers, True, fetch=None,
                             stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
print('T=%d N=%d ov=%.2f pers=%d: features %d, global phase %.2f ms' % (T, N, ov, pers, n, ms))
print({k: round(v, 3) for k, v in g.stats().items()})
