"""Target of the ncu captures: one warm-up pass + one pass of calc_clim, calc_anom, run_contrack (one threshold launch) on a
T-step synthetic cube.  usage: prof_target.py T [key=value ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from contrack_b200 import Engine
from contrack_b200.contrack import time_group_keys
T = int(sys.argv[1])
e = Engine.get(0)
for k, v in [kv.split('=') for kv in sys.argv[2:]]:
    e.set_option(k, int(v))
H, W = bench.H, bench.W
z = torch.empty((T, H, W), dtype=torch.float32, device='cuda'); bench.synth_fill(z, 0, T, season=True)
a = torch.empty_like(z); f = torch.empty((T, H, W), dtype=torch.int32, device='cuda')
lat, lon = bench.grid(); w = bench.reference_weights(lat, lon)
times = (np.datetime64('1981-01-01') + np.arange(T).astype('timedelta64[D]')).astype('datetime64[ns]')
u, g = np.unique(time_group_keys(times, 'dayofyear'), return_inverse=True)
for i in range(2):
    clim = e.calc_clim(z, g.astype(np.int32), len(u), 31)
    e.calc_anom(z, g.astype(np.int32), len(u), clim, 2, out=a)
    e.run_contrack(a, w, 160, True, 0, 0.5, 5, True, out=f)
    e.flag_count(f, 1)
torch.cuda.synchronize()
print({k: round(v, 3) for k, v in e.stats().items() if k.startswith('ms_')})
