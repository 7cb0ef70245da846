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


def test_float64_input_stays_float64(eng):
    """xarray keeps a float64 cube in float64 (packed ERA5 decoded with scale/offset): climatology and anomaly are float64
    and equal the float64 restatement to rounding, not merely to float32 precision."""
    T, H, W = 2 * 365 + 40, 6, 10
    z = make_z(T, H, W, 5).astype(np.float64) + 1e-7 * np.arange(H * W).reshape(H, W)     # not representable in float32
    keys, uniq, idx = doy_groups(T)
    _, cref = oracle.calc_clim(z, keys, 31)
    ref = oracle.calc_anom(z, keys, 31, 2)
    clim = eng.calc_clim(z, idx, len(uniq), 31)
    got = eng.calc_anom(z, idx, len(uniq), clim, 2)
    assert clim.dtype == np.float64 and got.dtype == np.float64
    np.testing.assert_allclose(clim, cref, rtol=1e-13, atol=1e-9)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-9, equal_nan=True)
    # float32 cube against a float64 climatology: numpy promotes the difference to float64
    got2 = eng.calc_anom(z.astype(np.float32), idx, len(uniq), clim, 1)
    assert got2.dtype == np.float64
    np.testing.assert_allclose(got2, z.astype(np.float32).astype(np.float64) - cref[np.searchsorted(uniq, keys)], rtol=1e-13,
                               atol=1e-9)
    # odd plane size: the scalar float64 path
    z3 = z[:, :3, :7].copy()
    c3 = eng.calc_clim(z3, idx, len(uniq), 5)
    np.testing.assert_allclose(eng.calc_anom(z3, idx, len(uniq), c3, 3), oracle.calc_anom(z3, keys, 5, 3), rtol=1e-12,
                               atol=1e-9, equal_nan=True)


def test_nan_gap_in_time_only_poisons_its_windows(eng):
    """A NaN at one time step makes exactly the rolling windows that contain it NaN (xarray's rolling mean), in the
    climatology smoothing and in the anomaly smoothing alike."""
    T, H, W = 365 + 60, 4, 8
    z = make_z(T, H, W, 6)
    z[100, 1, 2] = np.nan                                  # a single-sample group: the group mean itself is NaN
    keys, uniq, idx = doy_groups(T)
    ref = oracle.calc_anom(z, keys, 5, 3)
    clim = eng.calc_clim(z, idx, len(uniq), 5)
    got = eng.calc_anom(z, idx, len(uniq), clim, 3)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.isnan(ref[95:110, 1, 2]).sum() == 3 and not np.isnan(ref[120:-1, 1, 2]).any()       # (the last window is incomplete)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2 * ATOL, equal_nan=True)
