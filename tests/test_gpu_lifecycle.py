"""run_lifecycle (reference contrack.py:799-907) on the GPU against the oracle -- needs a B200 (`-m gpu`).

Every field must be IDENTICAL to the oracle's (which calls the reference's numpy / scipy functions): ids, date strings,
the int()-truncated centre of mass, and the float64 Intensity / Size -- the kernels accumulate in numpy's pairwise order
and in scipy's (np.bincount) sequential order, so no tolerance is needed."""
import json
import os

import numpy as np
import pytest

from oracle import contrack_oracle as oracle
from _common import row_weights
from _synth import synth_cube, regular_grid

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def eng():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from contrack_b200 import Engine
    return Engine.get(0)


def make_contrack(x, lat, lon, t0='2001-01-01', step_h=24):
    from contrack import contrack
    from contrack_b200 import Dataset, DataArray
    time = (np.datetime64(t0) + (np.arange(x.shape[0]) * step_h).astype('timedelta64[h]')).astype('datetime64[ns]')
    ds = Dataset({'anom': DataArray(x, ('time', 'latitude', 'longitude'), attrs={'units': 'gpm', 'long_name': 'Z'})},
                 coords={'time': time, 'latitude': lat, 'longitude': lon})
    c = contrack()
    c.read_xarray(ds)
    return c, time


def as_tuples(df):
    return [tuple(r) for r in df.itertuples(index=False, name=None)]


def same_rows(got, ref):
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert int(g[0]) == int(r[0]) and g[1] == r[1] and int(g[2]) == int(r[2]) and int(g[3]) == int(r[3]), (g, r)
        assert float(g[4]).hex() == float(r[4]).hex() and float(g[5]).hex() == float(r[5]).hex(), (g, r)


def test_reference_fixture_lifecycle():
    """reference tests/test_contrack.py:93-103: DataFrame, 3 flags, 28 rows -- and the whole table against the oracle and
    against the committed golden table."""
    import pandas as pd
    from contrack import contrack
    c = contrack(os.path.join(HERE, 'golden', 'anom_test.npz'))
    c.run_contrack(variable='anom', threshold=150, gorl='>=', overlap=0.5, persistence=5, twosided=False)
    test = c.run_lifecycle(flag='flag', variable='anom')
    assert type(test) == pd.DataFrame
    assert list(test.columns) == ['Flag', 'Date', 'Longitude', 'Latitude', 'Intensity', 'Size']
    assert len(test.Flag.unique()) == 3
    assert len(test) == 28
    d = np.load(os.path.join(HERE, 'golden', 'anom_test.npz'))
    f = oracle.run_contrack(d['anom'], d['latitude'], d['longitude'], 150, '>=', 0.5, 5, False)
    same_rows(as_tuples(test), oracle.run_lifecycle(f, d['anom'], d['latitude'], d['longitude'], d['time']))
    gold = json.load(open(os.path.join(HERE, 'golden', 'lifecycle_fixture.json')))
    same_rows(as_tuples(test), [(r[0], r[1], r[2], r[3], float.fromhex(r[4]), float.fromhex(r[5])) for r in gold['rows']])


@pytest.mark.parametrize('seed,shape,sigma,thr', [(3, (20, 91, 180), (2.0, 3, 5), 110), (7, (14, 181, 360), (2.0, 5, 8), 120),
                                                  (11, (9, 60, 47), (1.0, 3, 4), 80)])
def test_synthetic_lifecycle(eng, seed, shape, sigma, thr):
    """Features that cross the date line (rolled centre of mass), float32 and float64 variables, device-resident flag."""
    T, H, W = shape
    x = synth_cube(seed, T, H, W, sigma)
    lat = np.linspace(80, -80, H).astype(np.float32) if H == 60 else regular_grid(H, W)[0]
    lon = (np.arange(W) * (360.0 / W)).astype(np.float32) if H == 60 else regular_grid(H, W)[1]
    c, time = make_contrack(x, lat, lon, step_h=6)
    c.set_up(force=True)
    c.run_contrack('anom', thr, '>=', 0.4, 3, True)
    f = np.asarray(c.ds['flag'].data)
    ref = oracle.run_lifecycle(f, x, lat, lon, time, force=True)
    assert len(ref) > 0
    same_rows(as_tuples(c.run_lifecycle('flag', 'anom')), ref)
    st = eng.stats()
    assert st['lc_rows'] == len(ref)
    if seed != 11:
        assert st['lc_rolled'] > 0                                   # the cases must exercise the rolled branch
    # float64 variable: products and sums are formed from the float64 values
    c.ds['anom64'] = c._variable(('time', 'latitude', 'longitude'), x.astype(np.float64) * 1.000000123, {})
    ref64 = oracle.run_lifecycle(f, x.astype(np.float64) * 1.000000123, lat, lon, time, force=True)
    same_rows(as_tuples(c.run_lifecycle('flag', 'anom64')), ref64)


def test_arbitrary_flag_arrays(eng):
    """The flag variable need not come from run_contrack: ids that touch each other, an id that occupies every column
    (lon_roll = 1), an id present in both date-line columns with several gaps, negative variable values, empty planes."""
    rng = np.random.default_rng(5)
    T, H, W = 6, 12, 40
    lat = np.linspace(55, -55, H).astype(np.float32)
    lon = (np.arange(W) * 9.0).astype(np.float32)
    flag = np.zeros((T, H, W), np.int32)
    flag[0, 2:5, :] = 7                       # every column
    flag[0, 6:9, 0:4] = 9; flag[0, 6:9, 30:40] = 9; flag[0, 7, 12:15] = 9      # two gaps, wraps
    flag[1, 3, 5:9] = 4; flag[1, 3, 9:14] = 5; flag[1, 4, 5:14] = 4          # touching ids
    flag[2] = rng.integers(0, 4, (H, W))      # noise: many short runs, ids 1..3 in every column
    flag[4, 0, 0] = 2; flag[4, H - 1, W - 1] = 2                              # one pixel in each date-line column
    flag[5, 5, 10:30] = 123456
    var = (rng.standard_normal((T, H, W)) * 50 + 200).astype(np.float32)
    time = np.datetime64('1999-12-30T18') + (np.arange(T) * 6).astype('timedelta64[h]')
    ref = oracle.run_lifecycle(flag, var, lat, lon, time)
    res = eng.run_lifecycle(flag, var, row_weights(lat, lon))
    assert len(res['t']) == len(ref)
    from contrack import contrack
    from contrack_b200 import Dataset, DataArray
    dims = ('time', 'latitude', 'longitude')
    ds = Dataset({'v': DataArray(var, dims), 'flag': DataArray(flag, dims)},
                 coords={'time': time.astype('datetime64[ns]'), 'latitude': lat, 'longitude': lon})
    c = contrack(ds=ds)
    same_rows(as_tuples(c.run_lifecycle('flag', 'v')), ref)
    roll = {(int(t), int(l)): int(r) for t, l, r in zip(res['t'], res['label'], res['roll'])}
    assert roll[(0, 7)] == 1 and roll[(0, 9)] == 30 and roll[(1, 4)] == -1 and roll[(4, 2)] == W - 1
    # an all-zero cube gives an empty table
    empty = eng.run_lifecycle(np.zeros((3, 4, 8), np.int32), np.ones((3, 4, 8), np.float32), np.ones(4))
    assert len(empty['t']) == 0


def test_stale_box_pieces_lifecycle(eng):
    """Flags of cubes where the date-line merge splits components: pieces with different ids are 8-neighbours."""
    la, lo = regular_grid(24, 16)
    time = np.datetime64('2020-02-27') + np.arange(12).astype('timedelta64[D]')
    for seed in [1396, 1933]:
        xs = synth_cube(seed, 12, 24, 16, (1.5, 2, 2))
        f = oracle.track_persistence((xs >= 60).astype(int), 1).astype(np.int32)
        ref = oracle.run_lifecycle(f, xs, la, lo, time)
        res = eng.run_lifecycle(f, xs, row_weights(la, lo))
        assert len(res['t']) == len(ref)
        # compare the raw sums with the reference's calls for every row
        wg = oracle.weight_grid(la, oracle.resolution(la), oracle.resolution(lo), 16)
        for i in range(len(res['t'])):
            m = f[res['t'][i]] == res['label'][i]
            assert np.sum(wg[m]).hex() == float(res['area'][i]).hex()
            assert np.sum(wg[m] * xs[res['t'][i]][m]).hex() == float(res['wsum'][i]).hex()
