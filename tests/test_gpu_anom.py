"""calc_clim / calc_anom kernels against the numpy restatement in oracle/ (float32, tolerance) -- needs a B200."""
import numpy as np
import pytest

from oracle import contrack_oracle as oracle

pytestmark = pytest.mark.gpu

# float32 data of magnitude ~5500 (geopotential height): 1 ulp = 4.9e-4
RTOL, ATOL = 1e-6, 2e-3


@pytest.fixture(scope='module')
def eng():
    import torch
    assert torch.cuda.is_available()
    from contrack_b200 import Engine
    return Engine.get(0)


def make_z(T, H, W, seed=0, nan=False):
    rng = np.random.default_rng(seed)
    t = np.arange(T)[:, None, None]
    z = (5500 + 80 * np.cos(2 * np.pi * t / 365.25) + 100 * rng.standard_normal((T, H, W))).astype(np.float32)
    if nan:
        z[5:9, 2, 3] = np.nan
        z[:, 4, 4] = np.nan
    return z


def doy_groups(T, start='2001-01-01'):
    time = np.datetime64(start) + np.arange(T).astype('timedelta64[D]')
    from contrack_b200.contrack import time_group_keys
    keys = time_group_keys(time, 'dayofyear')
    uniq, idx = np.unique(keys, return_inverse=True)
    return keys, uniq, idx.astype(np.int32)


@pytest.mark.parametrize('window', [1, 2, 5, 31])
def test_calc_clim(eng, window):
    T, H, W = 3 * 365 + 200, 12, 20
    z = make_z(T, H, W, 1)
    keys, uniq, idx = doy_groups(T)
    k2, ref = oracle.calc_clim(z, keys, window)
    got = eng.calc_clim(z, idx, len(uniq), window)
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL)


def test_calc_clim_nan_and_leap_day(eng):
    T, H, W = 4 * 365 + 1 + 50, 6, 8                      # 2000 is a leap year: day 366 has a single sample
    z = make_z(T, H, W, 2, nan=True)
    keys, uniq, idx = doy_groups(T, '2000-01-01')
    assert uniq.max() == 366
    _, ref = oracle.calc_clim(z, keys, 31)
    got = eng.calc_clim(z, idx, len(uniq), 31)
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL, equal_nan=True)


@pytest.mark.parametrize('window,smooth', [(1, 1), (31, 2), (5, 3), (31, 8)])
def test_calc_anom(eng, window, smooth):
    T, H, W = 2 * 365 + 100, 10, 16
    z = make_z(T, H, W, 3)
    keys, uniq, idx = doy_groups(T)
    ref = oracle.calc_anom(z, keys, window, smooth)
    clim = eng.calc_clim(z, idx, len(uniq), window)
    got = eng.calc_anom(z, idx, len(uniq), clim, smooth)
    assert got.shape == z.shape and got.dtype == np.float32
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2 * ATOL, equal_nan=True)
    if smooth == 2:
        assert np.isnan(got[0]).all() and not np.isnan(got[1:]).any()


def test_device_resident_and_odd_plane_size(eng):
    import torch
    T, H, W = 400, 7, 9                                    # H*W = 63: the scalar (non float4) path
    z = make_z(T, H, W, 4)
    keys, uniq, idx = doy_groups(T)
    zd = torch.from_numpy(z).cuda()
    clim = eng.calc_clim(zd, idx, len(uniq), 5)
    an = eng.calc_anom(zd, idx, len(uniq), clim, 2)
    assert clim.is_cuda and an.is_cuda
    ref = oracle.calc_anom(z, keys, 5, 2)
    np.testing.assert_allclose(an.cpu().numpy(), ref, rtol=1e-5, atol=2 * ATOL, equal_nan=True)
