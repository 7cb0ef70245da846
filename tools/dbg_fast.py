"""Staged smoke of the fast table pipeline (short timeouts around it in the calling script): stage 1 exercises the plane
kernel alone (debug stage -> ordered phase on the host), then the cooperative global kernel.  usage: dbg_fast.py plane|global"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from oracle import contrack_oracle as oracle
from contrack_b200 import Engine
from _common import row_weights, same_partition
mode = sys.argv[1]
eng = Engine.get(0)
d = np.load(os.path.join(ROOT, 'tests', 'golden', 'anom_test.npz'))
a, lat, lon = d['anom'], d['latitude'], d['longitude']
w = row_weights(lat, lon)
xd = torch.from_numpy(a).cuda()
st = {}
ref = oracle.run_contrack(a, lat, lon, 150, '>=', 0.5, 5, True, stages=st)
t0 = time.time()
if mode == 'plane':
    for stage, key in ((1, 'label2d'), (2, 'label2d_seam'), (4, 'label3d')):
        f, n = eng.run_contrack(xd, w, 150, True, 0, 0.5, 5, True, stage=stage)
        torch.cuda.synchronize()
        f = f.cpu().numpy()
        ok = np.array_equal(f, st[key]) if stage == 4 else same_partition(f, st[key])
        print('stage', stage, 'ok' if ok else 'MISMATCH', {k: v for k, v in eng.stats().items() if k in ('fast_path', 'plane_attempts', 'runs', 'comps2d', 'pairs', 'seam_segments')}, flush=True)
        assert ok
else:
    for i in range(3):
        f, n = eng.run_contrack(xd, w, 150, True, 0, 0.5, 5, True)
        torch.cuda.synchronize()
        ok = np.array_equal(f.cpu().numpy(), ref)
        print('final', 'ok' if ok else 'MISMATCH', n, {k: v for k, v in eng.stats().items() if not k.startswith('ms_')}, flush=True)
        assert ok
print('done in %.1fs' % (time.time() - t0))
