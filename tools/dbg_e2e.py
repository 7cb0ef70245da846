"""Host-buffer entry point: where the time goes (H2D alone vs the call, host thread counts).  usage: dbg_e2e.py [T]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from contrack_b200 import Engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2707
H, W = bench.H, bench.W
eng = Engine.get(0)
lat, lon = bench.grid(); w = bench.reference_weights(lat, lon)
d = torch.empty((T, H, W), dtype=torch.float32, device='cuda'); bench.synth_fill(d, 0, T)
xin = torch.empty((T, H, W), dtype=torch.float32, pin_memory=True); xin.copy_(d)
fout = torch.empty((T, H, W), dtype=torch.int32, pin_memory=True)
torch.cuda.synchronize()
for chunk in ((64, 256, 1024) if len(sys.argv) <= 2 else ()):
    n = chunk * 1024 * 1024 // 4
    flat_h, flat_d = xin.view(-1), d.view(-1)
    for streams in (1, 2):
        ss = [torch.cuda.Stream() for _ in range(streams)]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i, off in enumerate(range(0, flat_h.numel(), n)):
            with torch.cuda.stream(ss[i % streams]):
                flat_d[off:off + n].copy_(flat_h[off:off + n], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print('H2D alone: chunk %4d MB, %d stream(s): %.1f GB/s' % (chunk, streams, flat_h.numel() * 4 / dt / 1e9), flush=True)
xn, fn = xin.numpy(), fout.numpy()
for threads in [int(a) for a in sys.argv[2:]] or [0, 2, 4, 8, 16]:
    eng.set_option('host_zero_threads', threads)
    for it in range(2):
        t0 = time.perf_counter()
        eng.run_contrack(xn, w, 160, True, 0, 0.5, 5, True, out=fn)
        dt = time.perf_counter() - t0
    s = eng.stats()
    print('host_zero_threads %2d: %.1f ms  (h2d+threshold %.1f, tables %.1f, paint %.1f)  %.0f timesteps/s' % (
        threads, dt * 1e3, s['ms_h2d_threshold'], s['ms_tables'], s['ms_paint_d2h'], T / dt), flush=True)

