"""GPU parity against results of the UNMODIFIED reference source (tests/golden/reference_run.json, produced in the build
container by tests/golden/make_reference_golden.py).  Everything goes through the reference-facing class
(contrack.read_xarray / set_up / run_contrack / run_lifecycle / calc_clim / calc_anom) -> ctypes -> C ABI -> CUDA.
No oracle involved: the expected values are the reference's own outputs."""
import os

import numpy as np
import pytest

from _common import sha_i4
from _synth import synth_cube, regular_grid

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _times(n, start='2000-01-01'):
    return (np.datetime64(start) + np.arange(n).astype('timedelta64[D]')).astype('datetime64[ns]')


def _thr(r):
    return np.float64(r['threshold']) if r.get('threshold_is_np_float64') else (
        int(r['threshold']) if float(r['threshold']).is_integer() else r['threshold'])


def _contrack(x, lat, lon, time, name='anom', dims=('time', 'latitude', 'longitude'), force=True):
    from contrack import contrack
    from contrack_b200 import Dataset
    order = [('time', 'latitude', 'longitude').index(d) for d in dims]
    ds = Dataset({name: (dims, np.ascontiguousarray(np.transpose(x, order)),
                         {'units': 'm', 'long_name': 'Geopotential Height'})},
                 coords={'time': time, 'latitude': lat, 'longitude': lon})
    c = contrack()
    c.read_xarray(ds)
    c.set_up(force=force)
    return c


def _life(df):
    return [[int(r.Flag), str(r.Date), int(r.Longitude), int(r.Latitude), float(r.Intensity).hex(), float(r.Size).hex()]
            for r in df.itertuples()]


def test_fixture(fixture_cube, reference_run):
    a, lat, lon = fixture_cube
    for r in reference_run['fixture']:
        c = _contrack(a, lat, lon, _times(11, '2016-10-02'), force=False)
        c.run_contrack(variable='anom', threshold=r['threshold'], gorl=r['gorl'], overlap=r['overlap'],
                       persistence=r['persistence'], twosided=r['twosided'])
        f = np.asarray(c['flag'])
        assert sha_i4(f) == r['sha256'], r['key']
        assert str(f.dtype) == r['dtype'] and c.variables == r['variables']
        assert dict(c['flag'].attrs) == r['attrs']
        assert _life(c.run_lifecycle(flag='flag', variable='anom')) == r['lifecycle'], r['key']


def test_synthetic(reference_run):
    for r in reference_run['synthetic']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        lat, lon = regular_grid(H, W)
        c = _contrack(x, lat, lon, _times(T))
        c.run_contrack('anom', _thr(r), r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(np.asarray(c['flag'])) == r['sha256'], r
        assert _life(c.run_lifecycle('flag', 'anom')) == r['lifecycle'], r


def test_stale_box_quirk(reference_run):
    for r in reference_run['quirk']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        lat, lon = regular_grid(H, W)
        c = _contrack(x, lat, lon, _times(T))
        c.run_contrack('anom', r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        f = np.asarray(c['flag'])
        assert sha_i4(f) == r['sha256'], r['seed']
        assert [int(i) for i in np.unique(f)[1:]] == r['ids']


def test_dim_order(reference_run):
    r = reference_run['dim_order']
    x = synth_cube(1, 30, 91, 180, (2.5, 3, 5))
    lat, lon = regular_grid(91, 180)
    c = _contrack(x, lat, lon, _times(30), dims=tuple(r['dims']))
    c.run_contrack('anom', 160, '>=', .5, 5, True)
    f = np.asarray(c['flag'])
    assert list(c['flag'].dims) == r['dims'] and list(f.shape) == r['shape']
    assert sha_i4(f) == r['sha256']


def test_calc_clim_anom(reference_run):
    """float32; tolerance 1e-5 relative / 4e-3 absolute on values of magnitude 5500 (1 ulp = 4.9e-4); the expected arrays
    come from the reference's calc_clim / calc_anom lines running on pandas-backed groupby / rolling."""
    for r in reference_run['anom']:
        d = np.load(os.path.join(HERE, 'golden', r['file']))
        T, H, W = r['shape']
        lat, lon = regular_grid(H, W)
        c = _contrack(d['z'], lat, lon, d['time'], name='z')
        clim = c.calc_clim('z', window=r['window'], groupby=r['groupby'])
        assert list(clim.shape) == r['clim_shape'] and clim.dims[0] == r['groupby']
        np.testing.assert_allclose(np.asarray(clim), d['clim'], rtol=1e-5, atol=4e-3)
        c.calc_anom('z', window=r['window'], smooth=r['smooth'], groupby=r['groupby'])
        an = np.asarray(c['anom'])
        assert np.array_equal(np.isnan(an), np.isnan(d['anom']))
        np.testing.assert_allclose(an, d['anom'], rtol=1e-5, atol=4e-3)
        assert dict(c['anom'].attrs) == r['anom_attrs']
