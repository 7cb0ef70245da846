"""Callers either side of run_contrack: quantile threshold, blocking frequency, geopotential height,
external climatology -- through the class (ctypes -> C ABI -> CUDA), against numpy / the oracle / the reference's outputs."""
import os

import numpy as np
import pytest

from oracle import contrack_oracle as oracle
from _synth import synth_cube, regular_grid

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def eng():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from contrack_b200 import Engine
    return Engine.get(0)


def _contrack(x, lat, lon, time, name='anom', attrs=None):
    from contrack import contrack
    from contrack_b200 import Dataset
    ds = Dataset({name: (('time', 'latitude', 'longitude'), x, attrs or {'units': 'm', 'long_name': 'Geopotential Height'})},
                 coords={'time': time, 'latitude': lat, 'longitude': lon})
    c = contrack()
    c.read_xarray(ds)
    c.set_up(force=True)
    return c


def _days(n, start='2000-01-01'):
    return (np.datetime64(start) + np.arange(n).astype('timedelta64[D]')).astype('datetime64[ns]')


@pytest.mark.parametrize('T', [1, 2, 7, 64, 365, 1001])
def test_quantile_time_is_numpy_nanquantile_bit_for_bit(eng, T):
    rng = np.random.default_rng(T)
    H, W = 6, 37
    x = (rng.standard_normal((T, H, W)) * 100).astype(np.float32)
    x[:, 0, :5] = np.round(x[:, 0, :5] / 50) * 50                      # heavy ties
    x[:, 1, 3] = 7.0                                                   # constant column
    if T > 2:
        x[rng.random((T, H, W)) < 0.1] = np.nan                        # NaN are skipped ...
        x[:, 2, 4] = np.nan                                            # ... an all-NaN column gives NaN
        x[0, 3, 0] = -0.0; x[1, 3, 0] = 0.0
    q = [0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 1.0, 1 / 3]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = oracle.quantile_time(x, q)
    got = eng.quantile_time(x, q)
    assert got.dtype == np.float64 and got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    m = ~np.isnan(ref)
    assert np.array_equal(got[m], ref[m])                              # float64 values identical
    sub = eng.quantile_time(x, [0.9], 2, 5)
    assert np.array_equal(sub[0][~np.isnan(sub[0])], ref[5, 2:5][~np.isnan(ref[5, 2:5])])
    with pytest.raises(ValueError):
        eng.quantile_time(x, [1.5])


def test_quantile_time_float64_input(eng):
    """A float64 cube is ranked and interpolated in float64 (numpy keeps the input precision)."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((57, 5, 21)) * 100 + 1e-9 * rng.standard_normal((57, 5, 21))
    x[rng.random(x.shape) < 0.1] = np.nan
    x[:, 1, 2] = np.nan
    x[:, 0, :4] = np.round(x[:, 0, :4] / 50) * 50
    q = [0.0, 0.1, 0.5, 0.9, 1.0, 1 / 3]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = oracle.quantile_time(x, q)
    got = eng.quantile_time(x, q)
    assert got.dtype == np.float64 and np.array_equal(np.isnan(got), np.isnan(ref))
    m = ~np.isnan(ref)
    assert np.array_equal(got[m], ref[m])
    assert not np.array_equal(got[m], oracle.quantile_time(x.astype(np.float32), q)[m])       # (float32 would differ)


@pytest.mark.parametrize('parts,dtype', [((30, 21, 40), np.float32), ((1, 50, 40), np.float32), ((45, 46), np.float64)])
def test_quantile_time_sharded_over_time_is_exact(parts, dtype):
    """The cube split along time over several ranks (in-process group: one host thread and context per rank): the per-point
    tallies are summed over the ranks between the passes, every rank gets np.nanquantile of the WHOLE cube bit for bit."""
    import threading
    import torch
    from contrack_b200 import Engine
    from contrack_b200.sharded import Comm
    rng = np.random.default_rng(11)
    T, H, W = sum(parts), 7, 33
    x = (rng.standard_normal((T, H, W)) * 100).astype(dtype)
    x[rng.random(x.shape) < 0.15] = np.nan
    x[:parts[0], 3, 5] = np.nan                                            # a point with no valid value on the first rank
    x[:, 4, 6] = np.nan
    x[:, 0, :6] = np.round(x[:, 0, :6] / 50) * 50
    q = [0.0, 0.1, 0.5, 0.9, 1.0, 1 / 3]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = oracle.quantile_time(x, q)[:, 1:6]
    n = len(parts)
    engines = [Engine(0) for _ in range(n)]
    comms = Comm.local_group(n)
    bounds = np.cumsum([0] + list(parts))
    res, errs = [None] * n, [None] * n

    def work(r):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                xd = torch.from_numpy(np.ascontiguousarray(x[bounds[r]:bounds[r + 1]])).cuda()
                res[r] = engines[r].quantile_time(xd, q, 1, 6, comm=comms[r]).cpu().numpy()
        except BaseException as e:                                         # noqa: BLE001
            errs[r] = e
    th = [threading.Thread(target=work, args=(r,)) for r in range(n)]
    [t.start() for t in th]
    [t.join() for t in th]
    for c in comms:
        c.close()
    for e in engines:
        e.lib.ct_destroy(e.handle); e.handle = None
    assert not any(errs), errs
    m = ~np.isnan(ref)
    for r in range(n):
        assert np.array_equal(np.isnan(res[r]), np.isnan(ref)) and np.array_equal(res[r][m], ref[m]), r


def test_quantile_threshold_recipe(eng):
    import torch
    T, H, W = 90, 91, 180
    x = synth_cube(4, T, H, W, (2.0, 3, 5))
    x[0] = np.nan                                                       # calc_anom(smooth=2) leaves the first step NaN
    lat, lon = regular_grid(H, W)
    c = _contrack(x, lat, lon, _days(T))
    thr = c.quantile_threshold('anom', 0.9, latitude=slice(80, 50))     # README.rst:151
    assert isinstance(thr, float) and thr == oracle.quantile_threshold(x, lat, 0.9, 80, 50)
    qf = c.quantile('anom', [0.5, 0.9], latitude=slice(80, 50))
    assert qf.dims == ('quantile', 'latitude', 'longitude') and qf.shape == (2, 16, 180)
    assert float(qf['latitude'].data[0]) == 80.0 and float(qf['latitude'].data[-1]) == 50.0
    # device-resident cube: same value, the result stays on the device
    from contrack_b200 import Dataset
    c2 = _contrack(torch.from_numpy(x).cuda(), lat, lon, _days(T))
    assert c2.quantile_threshold('anom', 0.9, latitude=slice(80, 50)) == thr
    # the threshold feeds run_contrack like any Python float
    c.run_contrack('anom', thr, '>=', 0.5, 3)
    ref = oracle.run_contrack(x, lat, lon, thr, '>=', 0.5, 3, True, force=True)
    assert np.array_equal(np.asarray(c['flag']), ref)


def test_blocking_frequency(eng):
    T, H, W = 60, 91, 180
    x = synth_cube(6, T, H, W, (2.0, 3, 5))
    lat, lon = regular_grid(H, W)
    c = _contrack(x, lat, lon, _days(T))
    c.run_contrack('anom', 120, '>=', 0.5, 3)
    f = np.asarray(c['flag'])
    freq = c.blocking_frequency('flag')
    assert freq.dims == ('latitude', 'longitude') and np.asarray(freq).dtype == np.float64
    assert np.array_equal(np.asarray(freq), oracle.blocking_frequency(f))              # README.rst:161
    assert np.array_equal(np.asarray(c.blocking_frequency('flag', greater_than=0)), oracle.blocking_frequency(f, 0))
    # odd plane size (no 16-byte vector path)
    g = (np.random.default_rng(1).random((13, 7, 9)) * 5).astype(np.int32)
    assert np.array_equal(eng.flag_count(g, 1), (g > 1).sum(0))


def test_gph_and_external_climatology_match_the_reference(reference_run):
    from contrack_b200 import DataArray
    r = reference_run['gph_extclim']
    d = np.load(os.path.join(HERE, 'golden', r['file']))
    T, H, W = r['shape']
    lat, lon = regular_grid(H, W)
    c = _contrack(d['gp'], lat, lon, d['time'], name='z', attrs={'units': 'm**2 s**-2', 'long_name': 'Geopotential'})
    with pytest.raises(ValueError, match='Geopotential unit should be'):
        c.calculate_gph_from_gp(gp_name='z', gp_unit='m')
    c.calculate_gph_from_gp(gp_name='z', gp_unit='m**2 s**-2', gph_name='z_height')
    gph = np.asarray(c['z_height'])
    assert str(gph.dtype) == r['gph_dtype'] and np.array_equal(gph, d['gph'])          # bit exact
    assert dict(c['z_height'].attrs) == r['gph_attrs'] and c.variables == ['z', 'z_height']
    clim = DataArray(d['clim'], ('dayofyear', 'latitude', 'longitude'),
                     coords={'dayofyear': DataArray(d['clim_doy'], ('dayofyear',)),
                             'latitude': DataArray(d['clim_lat'], ('latitude',)),
                             'longitude': DataArray(d['clim_lon'], ('longitude',))})
    c.calc_anom('z_height', smooth=r['smooth'], clim=clim)
    an = np.asarray(c['anom'])
    assert np.array_equal(np.isnan(an), np.isnan(d['anom']))
    np.testing.assert_allclose(an, d['anom'], rtol=1e-5, atol=4e-3)     # float32, 1 ulp at 5500 = 4.9e-4
    # a climatology without the group dimension cannot work in the reference either (contrack.py:562-565)
    with pytest.raises(ValueError):
        c.calc_anom('z_height', clim=DataArray(d['clim'][0], ('latitude', 'longitude'),
                                               coords={'latitude': DataArray(d['clim_lat'], ('latitude',)),
                                                       'longitude': DataArray(d['clim_lon'], ('longitude',))}))


def test_calc_mean(eng):
    rng = np.random.default_rng(3)
    T, H, W = 33, 9, 12
    z = (5500 + 50 * rng.standard_normal((T, H, W))).astype(np.float32)
    z[5, 2, 3] = np.nan
    lat, lon = regular_grid(H, W)
    c = _contrack(z, lat, lon, _days(T), name='z')
    m = c.calc_mean('z')
    assert m.dims == ('latitude', 'longitude')
    np.testing.assert_allclose(np.asarray(m), np.nanmean(z.astype(np.float64), axis=0), rtol=1e-6)
    assert c.calc_mean('nope') is None
