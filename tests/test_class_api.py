"""The `contrack` class: the reference's own tests (tests/test_contrack.py) restated for this package.
CPU tests cover the container / set_up logic; the tracking calls need a B200."""
import os

import numpy as np
import pytest

from contrack import contrack                     # drop-in import name (contrack/__init__.py:26 in the reference)
from contrack_b200 import Dataset, DataArray

HERE = os.path.dirname(os.path.abspath(__file__))
dataset = os.path.join(HERE, 'golden', 'anom_test.npz')


@pytest.fixture
def contracks():
    c = contrack()
    c.read(dataset)
    return c


def test_init_empty():
    assert contrack().ds is None                                  # tests/test_contrack.py:28-30


def test_init_file():
    c = contrack(dataset)                                         # :32-37
    assert type(c) == contrack and c.ds is not None


def test_read_wrong():
    with pytest.raises(IOError) as err:                           # :39-44
        contrack(os.path.join(HERE, 'golden', 'golden.json'))
    assert err.value.args[0] == "Unkown fileformat. Known formats are netcdf."


def test_read_xarray_like(contracks):
    c = contrack()
    c.read_xarray(contracks.ds)                                   # :45-49
    assert c.ds is contracks.ds
    with pytest.raises(ValueError, match='already set'):
        c.read_xarray(contracks.ds)
    with pytest.raises(ValueError, match='has to be a xarray data set'):
        contrack().read_xarray(np.zeros(3))


def test_len_ntime_dims_variables(contracks):
    assert len(contracks) == 1                                    # :51-52
    assert contracks.ntime == 11                                  # :54-55
    assert contracks.dimensions == ['latitude', 'longitude', 'time']   # :57-58
    assert contracks.variables == ['anom']                        # :60-61


def test_set_up(contracks):
    contracks.set_up(time_name='time', longitude_name='longitude', latitude_name='latitude')   # :63-69
    assert (contracks._time_name, contracks._longitude_name, contracks._latitude_name) == ('time', 'longitude', 'latitude')
    c2 = contrack()
    c2.read(dataset)
    c2.set_up()                                                   # :71-75
    assert (c2._time_name, c2._longitude_name, c2._latitude_name) == ('time', 'longitude', 'latitude')
    assert c2._dlat.shape == (1,) and c2._dlat[0] == 1 and c2._dlon[0] == 1      # length-1 arrays (contrack.py:352-355)


def test_irregular_grid_raises():
    lat = np.array([10, 9, 7.5, 7], np.float32)
    ds = Dataset({'anom': (('time', 'latitude', 'longitude'), np.zeros((2, 4, 3), np.float32))},
                 coords={'time': np.array(['2000-01-01', '2000-01-02'], 'datetime64[ns]'), 'latitude': lat,
                         'longitude': np.array([0, 1, 2], np.float32)})
    c = contrack(ds=ds)
    with pytest.raises(ValueError, match='No regular grid'):
        c.set_up()
    c.set_up(force=True)


def test_area_weights_match_reference_expression(contracks):
    from oracle import contrack_oracle as oracle
    contracks.set_up()
    lat, lon = contracks.ds['latitude'].data, contracks.ds['longitude'].data
    ref = oracle.weight_grid(lat, oracle.resolution(lat), oracle.resolution(lon), len(lon))[:, 0]
    assert np.array_equal(contracks.area_weights(), ref)


def test_bad_gorl_raises_before_touching_the_gpu(contracks):
    with pytest.raises(ValueError, match='Please select from'):
        contracks.run_contrack('anom', 150, '=>', 0.5, 5)


@pytest.mark.gpu
def test_run_contrack(contracks, golden):
    contracks.run_contrack(variable='anom', threshold=150, gorl='>=', overlap=0.5, persistence=5, twosided=False)   # :83-91
    assert contracks.variables == ['anom', 'flag']
    assert len(np.unique(contracks.flag)) - 1 == 3
    flag = contracks['flag']
    assert flag.dims == ('time', 'latitude', 'longitude') and flag.attrs['units'] == 'flag'
    assert 'threshold = >= 150' in flag.attrs['history']
    gold = np.load(os.path.join(HERE, 'golden', 'flags_fixture.npz'))['thr150_ov05_p5_one']
    assert np.array_equal(np.asarray(flag), gold)
    # 28 lifecycle rows = (time step, id) pairs (tests/test_contrack.py:93-103)
    f = np.asarray(flag)
    assert sum(len(np.unique(f[t])) - 1 for t in range(f.shape[0])) == 28


@pytest.mark.gpu
def test_run_contrack_other_dim_order_and_device_data(golden):
    import torch
    d = np.load(dataset)
    # (latitude, longitude, time) order is not involutive: the reference's transpose(sort) only works for
    # (time, lat, lon) and for swaps; use (time, longitude, latitude) -> sort = [0, 2, 1]
    a = np.ascontiguousarray(np.transpose(d['anom'], (0, 2, 1)))
    ds = Dataset({'anom': (('time', 'longitude', 'latitude'), torch.from_numpy(a).cuda())},
                 coords={'time': d['time'], 'latitude': d['latitude'], 'longitude': d['longitude']})
    c = contrack(ds=ds)
    c.run_contrack('anom', 150, '>=', 0.5, 5, twosided=True)
    flag = c['flag']
    assert flag.dims == ('time', 'longitude', 'latitude') and flag.data.is_cuda
    gold = np.load(os.path.join(HERE, 'golden', 'flags_fixture.npz'))['thr150_ov05_p5_two']
    assert np.array_equal(flag.values, np.transpose(gold, (0, 2, 1)))


@pytest.mark.gpu
def test_dayofyear_threshold_dataarray():
    from oracle import contrack_oracle as oracle
    d = np.load(dataset)
    c = contrack(dataset)
    from contrack_b200.contrack import time_group_keys
    doy = time_group_keys(d['time'], 'dayofyear')
    thr_vals = np.linspace(120, 180, len(doy))
    thr = DataArray(thr_vals, ('dayofyear',), coords={'dayofyear': DataArray(doy, ('dayofyear',))})
    c.run_contrack('anom', thr, '>=', 0.5, 3)
    ref = oracle.run_contrack(d['anom'], d['latitude'], d['longitude'], thr_vals, '>=', 0.5, 3, True)
    assert np.array_equal(np.asarray(c['flag']), ref)


@pytest.mark.gpu
def test_dayofyear_threshold_with_unsorted_coordinate_and_reference_dtype():
    """xarray aligns a day-of-year threshold by LABEL: a rolled coordinate gives the same result; reference_dtype=True must
    work for host and device data (the cube is far below 2**31 - 2 cells, so the result stays int32)."""
    import torch
    d = np.load(dataset)
    from contrack_b200.contrack import time_group_keys
    doy = time_group_keys(d['time'], 'dayofyear')
    thr_vals = np.linspace(120, 180, len(doy))
    c = contrack(dataset)
    c.run_contrack('anom', DataArray(thr_vals, ('dayofyear',), coords={'dayofyear': DataArray(doy, ('dayofyear',))}), '>=', 0.5, 3)
    ref = np.asarray(c['flag']).copy()
    c2 = contrack(dataset)
    k = 4
    rolled = DataArray(np.roll(thr_vals, k), ('dayofyear',), coords={'dayofyear': DataArray(np.roll(doy, k), ('dayofyear',))})
    c2.run_contrack('anom', rolled, '>=', 0.5, 3, reference_dtype=True)
    assert np.array_equal(np.asarray(c2['flag']), ref) and np.asarray(c2['flag']).dtype == np.int32
    ds = Dataset({'anom': (('time', 'latitude', 'longitude'), torch.from_numpy(d['anom']).cuda())},
                 coords={'time': d['time'], 'latitude': d['latitude'], 'longitude': d['longitude']})
    c3 = contrack(ds=ds)
    c3.run_contrack('anom', rolled, '>=', 0.5, 3, reference_dtype=True)
    assert c3['flag'].data.is_cuda and np.array_equal(c3['flag'].values, ref)


@pytest.mark.gpu
def test_calc_clim_and_anom_via_class():
    from oracle import contrack_oracle as oracle
    T, H, W = 800, 9, 12
    rng = np.random.default_rng(0)
    time = (np.datetime64('2001-01-01') + np.arange(T).astype('timedelta64[D]')).astype('datetime64[ns]')
    z = (5500 + 50 * rng.standard_normal((T, H, W))).astype(np.float32)
    ds = Dataset({'z': (('time', 'latitude', 'longitude'), z, {'units': 'm', 'long_name': 'Geopotential Height'})},
                 coords={'time': time, 'latitude': np.linspace(80, 40, H).astype(np.float32),
                         'longitude': np.arange(W).astype(np.float32) * 2})
    c = contrack(ds=ds)
    c.set_up()
    clim = c.calc_clim('z')                                        # tests/test_contrack.py:77-81
    assert isinstance(clim, DataArray)
    assert c.calc_clim('z', groupby='month').dims == ('month', 'latitude', 'longitude')
    c.calc_anom('z', window=31, smooth=2)
    assert c.variables == ['z', 'anom']
    assert c['anom'].attrs['long_name'] == 'Geopotential Height Anomaly'
    from contrack_b200.contrack import time_group_keys
    ref = oracle.calc_anom(z, time_group_keys(time, 'dayofyear'), 31, 2)
    np.testing.assert_allclose(np.asarray(c['anom']), ref, rtol=1e-5, atol=4e-3, equal_nan=True)


def test_single_time_step_raises_like_the_reference():
    # contrack.py:371 evaluates delta[0] of an empty difference vector: IndexError in set_up (observed by running the
    # reference on a one-step cube, tests/golden/make_reference_golden.py)
    ds = Dataset({'anom': (('time', 'latitude', 'longitude'), np.zeros((1, 4, 6), np.float32))},
                 coords={'time': np.array(['2000-01-01'], 'datetime64[ns]'), 'latitude': np.linspace(60, 30, 4).astype(np.float32),
                         'longitude': np.arange(6, dtype=np.float32) * 60})
    c = contrack(ds=ds)
    with pytest.raises(IndexError):
        c.set_up()


def test_time_shard_plumbing(monkeypatch):
    """run_contrack(time_shard=(t_begin, T_total)) hands this rank's planes, its own slice of a dayofyear threshold, the shard
    position and the communicator to the collective call, and stores what comes back as ds['flag'] (no GPU: the collective
    itself is replaced here; tests/test_zz_examples.py runs it on the GPU)."""
    from contrack_b200 import sharded
    from contrack_b200.contrack import time_group_keys
    d = np.load(dataset)
    T, H, W = d['anom'].shape
    doy = time_group_keys(d['time'], 'dayofyear')
    thr_vals = np.linspace(120, 180, T)
    thr = DataArray(thr_vals, ('dayofyear',), coords={'dayofyear': DataArray(doy, ('dayofyear',))})
    seen = {}

    def fake(engine, anom_local, t_begin, T_total, w, thresholds, thr_is_f32, op, overlap, persistence, twosided, out=None,
             comm=None, group=None):
        seen.update(engine=engine, shape=tuple(anom_local.shape), t_begin=t_begin, T_total=T_total, thr=np.array(thresholds),
                    thr_is_f32=thr_is_f32, op=op, overlap=overlap, persistence=persistence, twosided=twosided, comm=comm, w=w)
        return np.full(anom_local.shape, 7, np.int32), 1, {}
    monkeypatch.setattr(sharded, 'run_contrack_sharded', fake)
    lo, hi = 4, 9
    ds = Dataset({'anom': (('time', 'latitude', 'longitude'), d['anom'][lo:hi])},
                 coords={'time': d['time'][lo:hi], 'latitude': d['latitude'], 'longitude': d['longitude']})
    c = contrack(ds=ds)
    monkeypatch.setattr(c, '_engine', lambda: 'the engine')
    c.run_contrack('anom', thr, '>=', 0.5, 3, time_shard=(lo, T), comm='the communicator')
    assert seen['engine'] == 'the engine' and seen['comm'] == 'the communicator'
    assert seen['shape'] == (hi - lo, H, W) and (seen['t_begin'], seen['T_total']) == (lo, T)
    assert np.array_equal(seen['thr'], thr_vals[lo:hi]) and not seen['thr_is_f32']          # float64 thresholds
    assert (seen['op'], seen['overlap'], seen['persistence'], seen['twosided']) == (0, 0.5, 3, True)
    assert seen['w'].shape == (H,) and np.asarray(c['flag']).shape == (hi - lo, H, W) and (np.asarray(c['flag']) == 7).all()
    with pytest.raises(ValueError):
        c.run_contrack('anom', 150, '>=', 0.5, 3, time_shard=(T - 2, T))
