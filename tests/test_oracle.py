"""The oracle against every pin the reference offers for this path -- CPU only."""
import numpy as np
import pytest

from oracle import contrack_oracle as oracle
from _common import sha_i4
from _synth import synth_cube, regular_grid


def test_reference_test_pins(fixture_cube):
    # reference tests/test_contrack.py:83-91 (3 features) and 93-103 (28 lifecycle rows = (t, id) pairs)
    a, lat, lon = fixture_cube
    f = oracle.run_contrack(a, lat, lon, 150, '>=', 0.5, 5, twosided=False)
    assert oracle.num_features(f) == 3
    assert sum(len(np.unique(f[t])) - 1 for t in range(f.shape[0])) == 28
    assert f.dtype == np.int32 and f.shape == a.shape


def test_fixture_golden_hashes(fixture_cube, golden):
    a, lat, lon = fixture_cube
    for r in golden['fixture']:
        f = oracle.run_contrack(a, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(f) == r['sha256'], r['key']
        assert [int(i) for i in np.unique(f)[1:]] == r['ids']
        assert int((f > 0).sum()) == r['nonzero']


def test_survey_recorded_hashes(golden):
    # sha256 prefixes recorded independently in SURVEY.md section 8(c)
    want = {'thr150_ov05_p5_one': 'c392d20e9e2a626d', 'thr150_ov05_p5_two': '5c0d724cfe86d530',
            'thr160_ov05_p5_two': '1cb7c70a5c95dcc9', 'thr100_ov07_p3_two': 'fde4c3db10975813'}
    got = {r['key']: r['sha256'][:16] for r in golden['fixture']}
    assert got == want


def test_synthetic_golden(golden):
    for r in golden['synthetic']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        lat, lon = regular_grid(H, W)
        f = oracle.run_contrack(x, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(f) == r['sha256']


def test_stale_box_quirk_vectors(golden):
    for r in golden['quirk']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        f = oracle.track_persistence((x >= r['threshold']).astype(int), r['persistence'])
        assert sha_i4(f) == r['sha256']
        assert [int(i) for i in np.unique(f)[1:]] == r['ids']


def test_bad_gorl_raises(fixture_cube):
    a, lat, lon = fixture_cube
    with pytest.raises(ValueError, match='Please select from'):
        oracle.run_contrack(a, lat, lon, 150, '=>', 0.5, 5)


def test_rolling_mean_convention():
    x = np.arange(6, dtype=np.float32)[:, None]
    r2 = oracle._rolling_mean_centered(x, 2)[:, 0]
    assert np.isnan(r2[0]) and np.allclose(r2[1:], [0.5, 1.5, 2.5, 3.5, 4.5])
    r3 = oracle._rolling_mean_centered(x, 3)[:, 0]
    assert np.isnan(r3[0]) and np.isnan(r3[-1]) and np.allclose(r3[1:-1], [1, 2, 3, 4])


def test_lifecycle_reference_pins_and_golden(fixture_cube):
    """reference tests/test_contrack.py:93-103: 3 unique flags, 28 rows; plus the committed table of this restatement."""
    import json
    import os
    a, lat, lon = fixture_cube
    time = np.datetime64('2016-10-02') + np.arange(11).astype('timedelta64[D]')
    f = oracle.run_contrack(a, lat, lon, 150, '>=', 0.5, 5, twosided=False)
    rows = oracle.run_lifecycle(f, a, lat, lon, time)
    assert len(rows) == 28 and len({r[0] for r in rows}) == 3
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'lifecycle_fixture.json')))['rows']
    got = [[int(r[0]), r[1], int(r[2]), int(r[3]), float(r[4]).hex(), float(r[5]).hex()] for r in rows]
    assert got == gold
    assert rows == sorted(rows, key=lambda x: (x[0], x[1]))


# ---- pins made by the UNMODIFIED reference source executing in the build container (tests/golden/reference_run.json,
# ---- written by tests/golden/make_reference_golden.py under the xarray stand-in tests/golden/xr_shim) ----------------

def _times(n, start='2000-01-01'):
    return (np.datetime64(start) + np.arange(n).astype('timedelta64[D]')).astype('datetime64[ns]')


def _thr(r):
    return np.float64(r['threshold']) if r.get('threshold_is_np_float64') else (
        int(r['threshold']) if float(r['threshold']).is_integer() else r['threshold'])


def _life(rows):
    return [[int(r[0]), r[1], int(r[2]), int(r[3]), float(r[4]).hex(), float(r[5]).hex()] for r in rows]


def test_reference_run_fixture(fixture_cube, reference_run):
    a, lat, lon = fixture_cube
    time = _times(11, '2016-10-02')
    for r in reference_run['fixture']:
        f = oracle.run_contrack(a, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(f) == r['sha256'], r['key']
        assert str(f.dtype) == r['dtype']
        assert _life(oracle.run_lifecycle(f, a, lat, lon, time)) == r['lifecycle'], r['key']


def test_reference_run_synthetic(reference_run):
    for r in reference_run['synthetic']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        lat, lon = regular_grid(H, W)
        f = oracle.run_contrack(x, lat, lon, _thr(r), r['gorl'], r['overlap'], r['persistence'], r['twosided'], force=True)
        assert sha_i4(f) == r['sha256'], r
        assert _life(oracle.run_lifecycle(f, x, lat, lon, _times(T), force=True)) == r['lifecycle']


def test_reference_run_quirk(reference_run):
    for r in reference_run['quirk']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        lat, lon = regular_grid(H, W)
        f = oracle.run_contrack(x, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'],
                                force=True)
        assert sha_i4(f) == r['sha256'], r['seed']


def test_reference_run_calc_clim_anom(reference_run):
    """pandas-backed pin (the shim's groupby / rolling are an interpretation of xarray): float32, 1e-5 relative."""
    import os
    from contrack_b200.contrack import time_group_keys
    for r in reference_run['anom']:
        d = np.load(os.path.join(os.path.dirname(__file__), 'golden', r['file']))
        g = time_group_keys(d['time'], r['groupby'])
        keys, clim = oracle.calc_clim(d['z'], g, r['window'])
        assert list(clim.shape) == r['clim_shape']
        np.testing.assert_allclose(clim, d['clim'], rtol=1e-5, atol=4e-3)
        an = oracle.calc_anom(d['z'], g, r['window'], r['smooth'])
        assert np.array_equal(np.isnan(an), np.isnan(d['anom']))
        np.testing.assert_allclose(an, d['anom'], rtol=1e-5, atol=4e-3)


def test_reference_run_gph_and_external_climatology(reference_run):
    import os
    from contrack_b200.contrack import time_group_keys
    r = reference_run['gph_extclim']
    d = np.load(os.path.join(os.path.dirname(__file__), 'golden', r['file']))
    gph = oracle.gph_from_gp(d['gp'])
    assert gph.dtype == np.float32 and np.array_equal(gph, d['gph'])                   # float32 division: bit exact
    T, H, W = r['shape']
    lat, lon = regular_grid(H, W)
    an = oracle.calc_anom_external(d['gph'], time_group_keys(d['time'], 'dayofyear'), d['clim'], d['clim_doy'],
                                   d['clim_lat'], d['clim_lon'], lat, lon, r['smooth'])
    assert np.array_equal(np.isnan(an), np.isnan(d['anom'])) and int(np.isnan(an).sum()) == r['nan_anom']
    np.testing.assert_allclose(an, d['anom'], rtol=1e-5, atol=4e-3)


def test_readme_recipes_restated():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((50, 7, 9)).astype(np.float32)
    lat = np.linspace(90, -90, 7).astype(np.float32)
    thr = oracle.quantile_threshold(x, lat, 0.9, 80, 20)
    sl = oracle.label_slice(lat, 80, 20)
    assert (sl.start, sl.stop) == (1, 3)
    assert thr == float(np.mean(np.quantile(x[:, 1:3].astype(np.float32), np.array([0.9]), axis=0)))
    f = (rng.random((20, 4, 5)) * 4).astype(np.int32)
    assert np.array_equal(oracle.blocking_frequency(f), (f > 1).sum(0) / 20 * 100)


def test_reference_run_wider_parameter_space(reference_run):
    """Extreme overlaps, persistence 1 and longer than the cube, every gorl spelling, pole rows, float64 input, 2- and
    3-step cubes: outputs of the unmodified reference (tests/golden/make_reference_golden.py, key 'oracle_only')."""
    for r in reference_run['oracle_only']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        if r['float64']:
            x = x.astype(np.float64)
        lat, lon = regular_grid(H, W)
        f = oracle.run_contrack(x, lat, lon, _thr(r), r['gorl'], r['overlap'], r['persistence'], r['twosided'], force=True)
        assert sha_i4(f) == r['sha256'] and str(f.dtype) == r['dtype'], r
        assert [int(i) for i in np.unique(f)[1:]] == r['ids']
        assert _life(oracle.run_lifecycle(f, x, lat, lon, _times(T), force=True)) == r['lifecycle'], r
