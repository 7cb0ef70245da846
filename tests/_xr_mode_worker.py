"""Worker of tests/test_xarray_mode.py: the class boundary with an `xarray` module importable.

The image has no xarray wheel, so `tests/golden/xr_shim` (the stand-in the golden generator already uses: labelled-array
plumbing only) is put on sys.path AS `xarray` before `contrack_b200.contrack` is imported: its `xr is not None` branches
run.  The assertions are the reference's own tests (tests/test_contrack.py:28-103) restated against the fixture cube
(`anom_test.npz`, the raw array of the reference's anom_test.nc: no netCDF reader exists here, so the Dataset is built
with read_xarray instead of read).  usage: _xr_mode_worker.py cpu|gpu
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden', 'xr_shim'))

import xarray as xr                                     # noqa: E402  (the stand-in)
import pandas as pd                                     # noqa: E402
from contrack import contrack                           # noqa: E402  (repo-root package: the drop-in import name)
mod = sys.modules['contrack_b200.contrack']              # (the package re-exports the class under the same name)

assert mod.xr is xr, 'contrack_b200.contrack did not pick up the xarray module'


def fixture_ds():
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'anom_test.npz'))
    time = (np.datetime64('2016-10-02') + np.arange(d['anom'].shape[0]).astype('timedelta64[D]')).astype('datetime64[ns]')
    return xr.Dataset({'anom': (('time', 'latitude', 'longitude'), d['anom'], {'units': 'm', 'long_name': 'anomaly'})},
                      coords={'time': time, 'latitude': d['latitude'], 'longitude': d['longitude']})


def contracks():
    c = contrack()
    c.read_xarray(fixture_ds())
    return c


def cpu_tests():
    assert contrack().ds is None                                             # test_init_empty
    c = contracks()
    assert type(c.ds) is xr.Dataset                                          # test_init_netcdf / test_read_xarray
    try:                                                                     # test_read_wrong (the stand-in has no reader)
        contrack(os.path.join(ROOT, 'tests', 'golden', 'golden.json'))
        raise AssertionError('no IOError')
    except IOError as err:
        assert err.args[0] == "Unkown fileformat. Known formats are netcdf."
    try:
        c.read_xarray(fixture_ds())
        raise AssertionError('double read accepted')
    except ValueError as err:
        assert str(err) == 'contrack() is already set!'
    try:
        contrack().read_xarray(np.zeros(3))
        raise AssertionError('non-dataset accepted')
    except ValueError as err:
        assert str(err) == 'ds has to be a xarray data set!'
    assert len(c) == 1                                                       # test_len
    assert c.ntime == 11                                                     # test_ntime
    assert c.dimensions == ['latitude', 'longitude', 'time']                 # test_dimensions
    assert c.variables == ['anom']                                           # test_variables
    c.set_up(time_name='time', longitude_name='longitude', latitude_name='latitude')       # test_set_up_manually
    assert (c._time_name, c._longitude_name, c._latitude_name) == ('time', 'longitude', 'latitude')
    c = contracks()
    c.set_up()                                                               # test_set_up_automatic
    assert (c._time_name, c._longitude_name, c._latitude_name) == ('time', 'longitude', 'latitude')


def gpu_tests():
    c = contracks()
    c.set_up()
    clim = c.calc_clim('anom')                                               # test_calc_clim
    assert type(clim) is xr.DataArray
    m = c.calc_clim('anom', groupby='month')
    assert type(m) is xr.DataArray and m.dims == ('month', 'latitude', 'longitude')
    assert type(c.calc_mean('anom')) is xr.DataArray
    c = contracks()
    c.run_contrack(variable='anom', threshold=150, gorl='>=', overlap=0.5, persistence=5, twosided=False)   # test_run_caltrack
    assert c.variables == ['anom', 'flag']
    assert len(np.unique(c.flag)) - 1 == 3                                   # obj.flag through __getattr__ -> ds.flag
    assert type(c.ds['flag']) is xr.DataArray and isinstance(c.ds['flag'].data, np.ndarray)
    assert c['flag'].dims == ('time', 'latitude', 'longitude') and c['flag'].attrs['units'] == 'flag'
    test = c.run_lifecycle(flag='flag', variable='anom')                     # test_run_lifecycle
    assert type(test) == pd.DataFrame
    assert len(test.Flag.unique()) == 3
    assert len(test) == 28
    c.calc_anom('anom', window=3, smooth=2)
    assert type(c.ds['anom']) is xr.DataArray and c.ds['anom'].data.dtype == np.float32
    q = c.quantile('anom', [0.9])
    assert type(q) is xr.DataArray and q.dims == ('quantile', 'latitude', 'longitude')
    assert type(c.blocking_frequency('flag')) is xr.DataArray


if __name__ == '__main__':
    cpu_tests()
    if sys.argv[1:] == ['gpu']:
        gpu_tests()
    print('xarray-mode ok')
