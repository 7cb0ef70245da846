"""Config 3 of BASELINE.json on one GPU: calc_anom(smooth=2, window=31, groupby=dayofyear) on a synthetic 10957x721x1440
z cube (anomaly field + smooth seasonal cycle), then run_contrack on the anomaly; plus the README's quantile threshold and
blocking frequency.  CUDA-event times and algorithmic GB/s per stage (calc_clim 4 B/cell, calc_anom 8 B/cell,
run_contrack 8 B/cell).  usage: bench_anom.py [T] [reps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from contrack_b200 import Engine
from contrack_b200.contrack import time_group_keys

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10957
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W = bench.H, bench.W
cells = T * H * W
eng = Engine.get(0)
lat, lon = bench.grid(); w = bench.reference_weights(lat, lon)
times = (np.datetime64('1981-01-01') + np.arange(T).astype('timedelta64[D]')).astype('datetime64[ns]')
doy = time_group_keys(times, 'dayofyear')
uniq, gidx = np.unique(doy, return_inverse=True); gidx = gidx.astype(np.int32); G = len(uniq)
z = torch.empty((T, H, W), dtype=torch.float32, device='cuda')
bench.synth_fill(z, 0, T, season=True); torch.cuda.synchronize()
anom = torch.empty_like(z)
flag = torch.empty((T, H, W), dtype=torch.int32, device='cuda')


def timed(fn, n=reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r

out = {'T': T, 'H': H, 'W': W, 'peak_gbs': bench.measured_peak()[0]}
ms, clim = timed(lambda: eng.calc_clim(z, gidx, G, 31))
out['calc_clim'] = {'ms': ms, 'gbs': cells * 4 / ms / 1e6, 'bytes_per_cell': 4}
ms, _ = timed(lambda: eng.calc_anom(z, gidx, G, clim, 2, out=anom))
out['calc_anom'] = {'ms': ms, 'gbs': cells * 8 / ms / 1e6, 'bytes_per_cell': 8}
thr_q = None
y0, y1 = 40, 161                                                   # 80N .. 50N at 0.25 deg
ms, qf = timed(lambda: eng.quantile_time(anom, [0.9], y0, y1), 1)
out['quantile_90_band'] = {'ms': ms, 'points': (y1 - y0) * W, 'gbs': (y1 - y0) * W * T * 4 * 10 / ms / 1e6,
                           'note': '10 passes over the band (count + 8 radix passes + neighbour pass)',
                           'threshold': float(np.nanmean(qf.cpu().numpy()))}
ms, (f, n) = timed(lambda: eng.run_contrack(anom, w, 160, True, 0, 0.5, 5, True, out=flag))
out['run_contrack'] = {'ms': ms, 'gbs': cells * 8 / ms / 1e6, 'features': int(n), 'timesteps_per_s': T / ms * 1e3}
ms, cnt = timed(lambda: eng.flag_count(flag, 1))
out['blocking_frequency'] = {'ms': ms, 'gbs': cells * 4 / ms / 1e6}
t0 = time.perf_counter()
res = eng.run_lifecycle(flag, anom, w)
torch.cuda.synchronize()
t1 = time.perf_counter()
res = eng.run_lifecycle(flag, anom, w)
torch.cuda.synchronize()
ms = (time.perf_counter() - t1) * 1e3
out['run_lifecycle'] = {'ms': ms, 'first_call_ms': (t1 - t0) * 1e3, 'rows': int(len(res['t'])), 'gbs': cells * 8 / ms / 1e6,
                        'stats': {k: v for k, v in eng.stats().items() if k.startswith('ms_lc') or k.startswith('lc_')}}
tot = out['calc_clim']['ms'] + out['calc_anom']['ms'] + out['run_contrack']['ms']
out['config3_total'] = {'ms': tot, 'timesteps_per_s': T / tot * 1e3, 'gbs': cells * 20 / tot / 1e6,
                        'frac_of_peak': cells * 20 / tot / 1e6 / out['peak_gbs'], 'bytes_per_cell': 20}
print(json.dumps(out))
