#!/usr/bin/env python3
"""Golden vectors produced by the UNMODIFIED reference source, run in the BUILD container.

`import contrack` of /root/reference fails here only because xarray has no wheel (contrack.py:19).  This script puts
tests/golden/xr_shim (a stand-in for the labelled-array plumbing, see its docstring) in front of /root/reference on
sys.path, imports the reference package as it is, and calls its own public API -- `contrack()`, `read_xarray`, `set_up`,
`run_contrack`, `run_lifecycle`, `calc_clim`, `calc_anom` -- on the reference's fixture and on seeded synthetic cubes.
Every arithmetic statement that runs is the reference's (contrack.py:646-772, 845-907); the shim supplies indexing,
transpose, isel, roll, where, and (for calc_clim / calc_anom only) pandas-backed groupby / rolling.

Writes tests/golden/reference_run.json: sha256 of each result ('<i4' C-order flag cube), ids, counts, the lifecycle table,
and for calc_clim / calc_anom sha256 + a few probe values of the float32 arrays.  tests/test_oracle.py holds the oracle to
these.  /root/reference is absent on the GPU box: nothing at test time reads it, only this generator does.
"""
import hashlib
import json
import logging
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'xr_shim'))
warnings.simplefilter('ignore')

import xarray as xr                                    # noqa: E402  (the shim)
assert xr.__version__.endswith('shim')
from contrack import contrack                          # noqa: E402  (the reference, unmodified)
import contrack as _pkg                                # noqa: E402
assert os.path.realpath(_pkg.__file__).startswith('/root/reference/'), _pkg.__file__
logging.disable(logging.CRITICAL)

from _synth import synth_cube, regular_grid            # noqa: E402


def sha_i4(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype('<i4').tobytes()).hexdigest()


def sha_raw(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def dataset(x, lat, lon, time, name='anom', dims=('time', 'latitude', 'longitude')):
    order = [('time', 'latitude', 'longitude').index(d) for d in dims]
    coords = {'time': time, 'latitude': lat, 'longitude': lon}
    return xr.Dataset({name: (dims, np.transpose(x, order), {'units': 'm', 'long_name': 'Geopotential Height'})},
                      coords=coords)


def reference_run(x, lat, lon, time, thr, gorl, ov, pers, two, lifecycle=False, force=False):
    c = contrack()
    c.read_xarray(dataset(x, lat, lon, time))
    c.set_up(force=force)
    c.run_contrack(variable='anom', threshold=thr, gorl=gorl, overlap=ov, persistence=pers, twosided=two)
    f = np.asarray(c.ds['flag'].data)
    rows = None
    if lifecycle:
        df = c.run_lifecycle(flag='flag', variable='anom')
        rows = [[int(r.Flag), str(r.Date), int(r.Longitude), int(r.Latitude), float(r.Intensity).hex(),
                 float(r.Size).hex()] for r in df.itertuples()]
    return f, rows, c


def days(n, start='2016-10-02'):
    return (np.datetime64(start) + np.arange(n).astype('timedelta64[D]')).astype('datetime64[ns]')


def main():
    fx = np.load(os.path.join(HERE, 'anom_test.npz'))
    a, lat, lon, time = fx['anom'], fx['latitude'], fx['longitude'], fx['time']
    out = {'how': 'unmodified /root/reference/contrack/contrack.py executed under tests/golden/xr_shim '
                  '(make_reference_golden.py); numpy %s, scipy %s' % (np.__version__, __import__('scipy').__version__),
           'fixture': [], 'synthetic': [], 'quirk': [], 'anom': []}

    for thr, ov, pers, two in [(150, .5, 5, False), (150, .5, 5, True), (160, .5, 5, True), (100, .7, 3, True)]:
        f, rows, c = reference_run(a, lat, lon, time, thr, '>=', ov, pers, two, lifecycle=True)
        key = 'thr%d_ov%02d_p%d_%s' % (thr, int(ov * 10), pers, 'two' if two else 'one')
        out['fixture'].append(dict(key=key, threshold=thr, gorl='>=', overlap=ov, persistence=pers, twosided=two,
                                   dtype=str(f.dtype), ids=[int(i) for i in np.unique(f)[1:]],
                                   nonzero=int((f > 0).sum()), sha256=sha_i4(f), lifecycle=rows,
                                   variables=list(c.variables), attrs=dict(c.ds['flag'].attrs)))
        print(key, f.dtype, out['fixture'][-1]['ids'], len(rows))

    for seed, T, H, W, sig, thr, gorl, ov, pers, two in [
            (1, 30, 91, 180, (2.5, 3, 5), 160, '>=', .5, 5, True),
            (1, 30, 91, 180, (2.5, 3, 5), 160, '>=', .5, 5, False),
            (1, 30, 91, 180, (2.5, 3, 5), -160, '<', .5, 5, True),
            (1, 30, 91, 180, (2.5, 3, 5), -150.7, 'le', .5, 5, True),
            (3, 40, 91, 180, (1.5, 2, 3), 120, '>', .3, 3, True),
            (3, 40, 91, 180, (1.5, 2, 3), np.float64(120.3), 'gt', .3, 3, True),
            (5, 24, 181, 360, (2.0, 4, 6), 150, 'ge', .7, 4, True),
            (7, 20, 46, 90, (1.0, 1.5, 2), 100, '>=', .5, 2, True),
            (8, 20, 46, 90, (1.0, 1.5, 2), 100, '>=', .9, 2, False)]:
        x = synth_cube(seed, T, H, W, sig)
        la, lo = regular_grid(H, W)
        f, rows, _ = reference_run(x, la, lo, days(T, '2000-01-01'), thr, gorl, ov, pers, two, lifecycle=True, force=True)
        out['synthetic'].append(dict(seed=seed, shape=[T, H, W], sigma=list(sig), threshold=float(thr),
                                     threshold_is_np_float64=isinstance(thr, np.float64), gorl=gorl, overlap=ov,
                                     persistence=pers, twosided=two, input_sha256=sha_raw(x),
                                     ids=[int(i) for i in np.unique(f)[1:]], nonzero=int((f > 0).sum()),
                                     sha256=sha_i4(f), lifecycle=rows))
        print('synthetic', seed, (T, H, W), len(out['synthetic'][-1]['ids']), 'features', len(rows), 'rows')

    # stale-bounding-box date-line quirk (contrack.py:753-763): whole run_contrack with overlap 0 (filter keeps all)
    for seed in [1011, 1137, 1141, 1207, 1219, 1233, 1260, 1280, 1288, 1317, 1339, 1341, 1367, 1396]:
        x = synth_cube(seed, 12, 24, 16, (1.5, 2, 2))
        la, lo = regular_grid(24, 16)
        f, _, _ = reference_run(x, la, lo, days(12, '2000-01-01'), 60, '>=', 0.0, 1, False, force=True)
        out['quirk'].append(dict(seed=seed, shape=[12, 24, 16], sigma=[1.5, 2, 2], threshold=60, gorl='>=', overlap=0.0,
                                 persistence=1, twosided=False, ids=[int(i) for i in np.unique(f)[1:]], sha256=sha_i4(f)))
    print('quirk', len(out['quirk']))

    # a non-(time, lat, lon) dimension order the reference supports (involutive permutation)
    x = synth_cube(1, 30, 91, 180, (2.5, 3, 5))
    la, lo = regular_grid(91, 180)
    c = contrack()
    c.read_xarray(dataset(x, la, lo, days(30, '2000-01-01'), dims=('latitude', 'time', 'longitude')))
    c.set_up(force=True)
    c.run_contrack(variable='anom', threshold=160, gorl='>=', overlap=.5, persistence=5, twosided=True)
    f = np.asarray(c.ds['flag'].data)
    out['dim_order'] = dict(dims=list(c.ds['flag'].dims), shape=list(f.shape), sha256=sha_i4(f))

    # more of the parameter space, held against the oracle only (tests/test_oracle.py): extreme overlaps, persistence 1 and
    # longer than the cube, every gorl spelling, a 2-degree grid with pole rows, float64 input
    out['oracle_only'] = []
    for seed, T, H, W, sig, thr, gorl, ov, pers, two, f64 in [
            (21, 16, 91, 180, (1.5, 3, 5), 120, '>=', 0.0, 1, True, False),
            (21, 16, 91, 180, (1.5, 3, 5), 120, '>=', 1.0, 1, True, False),
            (21, 16, 91, 180, (1.5, 3, 5), 120, 'ge', 0.99, 2, False, False),
            (22, 12, 91, 180, (1.5, 3, 5), -120, 'lt', 0.5, 3, True, False),
            (22, 12, 91, 180, (1.5, 3, 5), -120, '<=', 0.5, 30, True, False),
            (23, 25, 46, 90, (2.0, 1.5, 2), 90, 'gt', 0.3, 6, True, True),
            (24, 10, 181, 360, (1.0, 6, 10), 140, '>=', 0.6, 2, False, False),
            (25, 3, 46, 90, (0.5, 1.5, 2), 80, '>', 0.5, 1, True, False),
            (26, 2, 46, 90, (0.5, 1.5, 2), 80, '>', 0.5, 1, True, False)]:      # (one time step: the reference's set_up raises)
        x = synth_cube(seed, T, H, W, sig)
        if f64:
            x = x.astype(np.float64)
        la, lo = regular_grid(H, W)
        f, rows, _ = reference_run(x, la, lo, days(T, '2000-01-01'), thr, gorl, ov, pers, two, lifecycle=T > 1, force=True)
        out['oracle_only'].append(dict(seed=seed, shape=[T, H, W], sigma=list(sig), threshold=float(thr), gorl=gorl, overlap=ov,
                                       persistence=pers, twosided=two, float64=f64, dtype=str(f.dtype),
                                       ids=[int(i) for i in np.unique(f)[1:]], sha256=sha_i4(f), lifecycle=rows))
    print('oracle_only', len(out['oracle_only']), [len(r['ids']) for r in out['oracle_only']])

    # calc_clim / calc_anom (pandas-backed groupby / rolling in the shim: an interpretation of xarray)
    for seed, nyear, H, W, window, smooth, groupby in [(11, 3, 5, 8, 31, 2, 'dayofyear'), (12, 2, 5, 8, 5, 3, 'dayofyear'),
                                                       (13, 2, 5, 8, 1, 1, 'month')]:
        T = 365 * nyear + 1
        t = days(T, '1999-01-01')
        rng = np.random.default_rng(seed)
        doy = np.asarray((t.astype('datetime64[D]') - t.astype('datetime64[Y]').astype('datetime64[D]')).astype(int))
        z = (5500 + 80 * np.cos(2 * np.pi * doy / 365.25)[:, None, None] +
             30 * rng.standard_normal((T, H, W))).astype(np.float32)
        la, lo = regular_grid(H, W)
        c = contrack()
        c.read_xarray(dataset(z, la, lo, t, name='z'))
        c.set_up(force=True)
        clim = np.asarray(c.calc_clim('z', window=window, groupby=groupby).data)
        c.calc_anom('z', window=window, smooth=smooth, groupby=groupby)
        an = np.asarray(c.ds['anom'].data)
        np.savez_compressed(os.path.join(HERE, 'anom_ref_%d.npz' % seed), z=z, time=t, clim=clim, anom=an)
        out['anom'].append(dict(seed=seed, shape=[T, H, W], window=window, smooth=smooth, groupby=groupby,
                                file='anom_ref_%d.npz' % seed, clim_shape=list(clim.shape),
                                nan_anom=int(np.isnan(an).sum()), clim_dtype=str(clim.dtype), anom_dtype=str(an.dtype),
                                anom_attrs=dict(c.ds['anom'].attrs)))
        print('anom', seed, clim.shape, an.dtype, int(np.isnan(an).sum()))

    # calculate_gph_from_gp (contrack.py:386-425) and calc_anom with a supplied climatology on a coarser grid (551-565)
    rng = np.random.default_rng(21)
    T, H, W = 40, 9, 16
    t = days(T, '2001-02-20')
    la, lo = regular_grid(H, W)
    gp = (53000 + 900 * rng.standard_normal((T, H, W))).astype(np.float32)
    c = contrack()
    ds = xr.Dataset({'z': (('time', 'latitude', 'longitude'), gp, {'units': 'm**2 s**-2', 'long_name': 'Geopotential'})},
                    coords={'time': t, 'latitude': la, 'longitude': lo})
    c.read_xarray(ds)
    c.set_up(force=True)
    c.calculate_gph_from_gp(gp_name='z', gp_unit='m**2 s**-2', gph_name='z_height')
    gph = np.asarray(c.ds['z_height'].data)
    Hc, Wc = 5, 7
    clat = np.linspace(90, -90, Hc).astype(np.float32)
    clon = (np.arange(Wc) * (360.0 / Wc)).astype(np.float32)
    cdoy = np.arange(1, 367)
    cl = (5400 + 60 * np.cos(2 * np.pi * cdoy / 366.0)[:, None, None] + 20 * rng.standard_normal((366, Hc, Wc))).astype(np.float32)
    clim_da = xr.DataArray(cl, ('dayofyear', 'latitude', 'longitude'),
                           coords={'dayofyear': cdoy, 'latitude': clat, 'longitude': clon})
    c.calc_anom('z_height', smooth=3, clim=clim_da)
    an = np.asarray(c.ds['anom'].data)
    np.savez_compressed(os.path.join(HERE, 'gph_extclim_ref.npz'), gp=gp, time=t, gph=gph, clim=cl, clim_lat=clat,
                        clim_lon=clon, clim_doy=cdoy, anom=an)
    out['gph_extclim'] = dict(file='gph_extclim_ref.npz', shape=[T, H, W], smooth=3, gph_dtype=str(gph.dtype),
                              gph_attrs=dict(c.ds['z_height'].attrs), nan_anom=int(np.isnan(an).sum()))
    print('gph / external clim', gph.dtype, an.dtype, int(np.isnan(an).sum()))

    json.dump(out, open(os.path.join(HERE, 'reference_run.json'), 'w'), indent=0)
    print('wrote reference_run.json')


if __name__ == '__main__':
    main()
