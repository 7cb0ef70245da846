"""Target of the compute-sanitizer runs (memcheck / racecheck / synccheck): the reference's fixture, one stale-box quirk cube,
a cube with pole rows, the time-sharded path (three contexts on one GPU) and calc_anom / quantile / lifecycle, each checked
against the oracle.  Small on purpose: the sanitizer slows kernels by 10-100x.  usage: sanitize_target.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from oracle import contrack_oracle as oracle
from contrack_b200 import Engine
from _common import row_weights
from _synth import synth_cube, regular_grid

eng = Engine.get(0)
d = np.load(os.path.join(ROOT, 'tests', 'golden', 'anom_test.npz'))
cases = [(d['anom'], d['latitude'], d['longitude'], 150, 0.5, 5, True),
         (synth_cube(1396, 12, 24, 16, (1.5, 2, 2)),) + regular_grid(24, 16) + (60, 0.0, 1, False),
         (synth_cube(1003, 12, 24, 16, (1.5, 2, 2)),) + regular_grid(24, 16) + (60, 0.5, 2, True),
         (synth_cube(2, 6, 721, 1440, (2.5, 24, 40)), np.linspace(90, -90, 721).astype(np.float32),
          (np.arange(1440) * 0.25).astype(np.float32), 160, 0.5, 2, True)]
for x, lat, lon, thr, ov, pers, two in cases:
    force = len(lat) == 721
    ref = oracle.run_contrack(x, lat, lon, thr, '>=', ov, pers, two, force=force)
    w = oracle.weight_grid(lat, oracle.resolution(lat, force), oracle.resolution(lon, force), len(lon))[:, 0].copy()
    f, n = eng.run_contrack(torch.from_numpy(x).cuda(), w, thr, True, 0, ov, pers, two)
    torch.cuda.synchronize()
    assert np.array_equal(f.cpu().numpy(), ref), 'device path'
    fh, _ = eng.run_contrack(x, w, thr, True, 0, ov, pers, two)
    assert np.array_equal(fh, ref), 'host-buffer path'
    print('run_contrack ok', x.shape, n, flush=True)
# sharded, three contexts on this GPU
from test_gpu_sharded import run_local
a, lat, lon = cases[0][:3]
f, n, _ = run_local(a, row_weights(lat, lon), (4, 3, 4), 150, '>=', 0.5, 5, True)              # tables through peer windows
f0, n0, _ = run_local(a, row_weights(lat, lon), (4, 3, 4), 150, '>=', 0.5, 5, True, opts={'p2p': 0})   # all-gather
assert np.array_equal(f, f0) and n == n0
for opts in ({'plane_kernel': 1}, {'plane_kernel': 0}, {'gpu_tables': 0}):                   # both table builders, host ordered phase
    for k, v in opts.items():
        eng.set_option(k, v)
    f1, _ = eng.run_contrack(torch.from_numpy(a).cuda(), row_weights(lat, lon), 150, True, 0, 0.5, 5, True)
    assert np.array_equal(f1.cpu().numpy(), f)
eng.set_option('plane_kernel', 2); eng.set_option('gpu_tables', 1)
assert np.array_equal(f, oracle.run_contrack(a, lat, lon, 150, '>=', 0.5, 5, True))
print('sharded ok', flush=True)
# calc_anom, quantile, lifecycle
T = 400
z = (5500 + 100 * np.random.default_rng(0).standard_normal((T, 12, 20))).astype(np.float32)
from contrack_b200.contrack import time_group_keys
keys = time_group_keys(np.datetime64('2001-01-01') + np.arange(T).astype('timedelta64[D]'), 'dayofyear')
u, g = np.unique(keys, return_inverse=True)
clim = eng.calc_clim(z, g.astype(np.int32), len(u), 5)
an = eng.calc_anom(z, g.astype(np.int32), len(u), clim, 2)
np.testing.assert_allclose(an, oracle.calc_anom(z, keys, 5, 2), rtol=1e-5, atol=4e-3, equal_nan=True)
q = eng.quantile_time(z, [0.1, 0.9])
assert np.array_equal(q, oracle.quantile_time(z, [0.1, 0.9]))
fl, _ = eng.run_contrack(a, row_weights(lat, lon), 150, True, 0, 0.5, 5, True)
eng.run_lifecycle(fl, a, row_weights(lat, lon))
print('anom / quantile / lifecycle ok', flush=True)
