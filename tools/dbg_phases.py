"""Per-phase stats of full-size single-GPU steps for a few chunk counts (debug aid)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from contrack_b200 import Engine
eng = Engine.get(0)
T=10957
lat, lon = bench.grid(); w = bench.reference_weights(lat, lon)
anom = torch.empty((T, bench.H, bench.W), dtype=torch.float32, device='cuda'); bench.synth_fill(anom, 0, T); torch.cuda.synchronize()
flag = torch.empty((T, bench.H, bench.W), dtype=torch.int32, device='cuda')
for ch in (4, 3, 5):
    eng.set_option('chunks', ch)
    for i in range(3):
        eng.run_contrack(anom, w, 160, True, 0, 0.5, 5, True, out=flag)
    s = eng.stats()
    print(ch, {k: round(v,3) for k,v in s.items() if k.startswith('ms_') or k in ('chunks','sweeps','kernel_launches')})
