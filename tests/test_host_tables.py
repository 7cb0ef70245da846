"""Ordered table phase of the library (ct_host_tables: host C++, no GPU) against the oracle -- CPU only.

The tables are built with numpy/scipy (tests/_tables_np.py) exactly as the CUDA kernels define them, so this checks the
time-sequential overlap filter, scipy-order 3-D numbering, the stale-box date-line merge and the persistence filter."""
import numpy as np
import pytest

from oracle import contrack_oracle as oracle
from _common import row_weights, sha_i4
from _synth import synth_cube, regular_grid
from _tables_np import build_tables, host_tables

# 12x24x16 cubes (sigma 1.5,2,2; mask >= 60) in which the stale-box date-line merge cuts a component in two
SPLIT_SEEDS = [1396, 1933, 2136, 2257]


@pytest.mark.parametrize('thr,ov,pers,two', [(150, .5, 5, False), (150, .5, 5, True), (160, .5, 5, True),
                                             (100, .7, 3, True)])
def test_fixture(fixture_cube, golden, thr, ov, pers, two):
    a, lat, lon = fixture_cube
    tb = build_tables(a >= thr, row_weights(lat, lon))
    f, st = host_tables(tb, ov, pers, two)
    key = 'thr%d_ov%02d_p%d_%s' % (thr, int(ov * 10), pers, 'two' if two else 'one')
    want = [r for r in golden['fixture'] if r['key'] == key][0]
    assert sha_i4(f) == want['sha256']
    assert st[0] == len(want['ids'])


def test_stages(fixture_cube):
    a, lat, lon = fixture_cube
    stages = {}
    oracle.run_contrack(a, lat, lon, 150, '>=', .5, 5, True, stages=stages)
    tb = build_tables(a >= 150, row_weights(lat, lon))
    f3, _ = host_tables(tb, .5, 5, True, stage=3)
    assert np.array_equal(f3 > 0, stages['filtered'] > 0)
    f4, _ = host_tables(tb, .5, 5, True, stage=4)
    assert np.array_equal(f4, stages['label3d'])


def test_synthetic_golden(golden):
    for r in golden['synthetic']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        lat, lon = regular_grid(H, W)
        op = {'>=': np.greater_equal, 'ge': np.greater_equal, '<': np.less, '>': np.greater}[r['gorl']]
        tb = build_tables(op(x, np.float32(r['threshold'])), row_weights(lat, lon))
        f, _ = host_tables(tb, r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(f) == r['sha256']


def test_stale_box_quirk_vectors(golden):
    lat, lon = regular_grid(24, 16)
    w = row_weights(lat, lon)
    for r in golden['quirk']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        tb = build_tables(x >= r['threshold'], w)
        f, _ = host_tables(tb, 0.0, r['persistence'], False)      # overlap 0, one-sided: step 3 keeps everything
        assert sha_i4(f) == r['sha256']


def test_random_seam_heavy_cubes():
    lat, lon = regular_grid(24, 16)
    w = row_weights(lat, lon)
    splits = 0
    for seed in list(range(1000, 1100)) + SPLIT_SEEDS:
        x = synth_cube(seed, 12, 24, 16, (1.5, 2, 2))
        two, pers, ov = seed % 2 == 0, 1 + seed % 3, [0.3, 0.5, 0.7][seed % 3]
        tb = build_tables(x >= 60, w)
        f, st = host_tables(tb, ov, pers, two)
        assert np.array_equal(f, oracle.run_contrack(x, lat, lon, 60, '>=', ov, pers, two)), seed
        f, st = host_tables(tb, 0.0, 1, False)
        assert np.array_equal(f, oracle.track_persistence((x >= 60).astype(int), 1)), seed
        splits += st[4]
    assert splits >= 14          # SPLIT_SEEDS split a component at a stale box


def test_split_without_runs_is_an_error():
    lat, lon = regular_grid(24, 16)
    w = row_weights(lat, lon)
    from contrack_b200._lib import ContrackLibError
    x = synth_cube(SPLIT_SEEDS[0], 12, 24, 16, (1.5, 2, 2))
    tb = build_tables(x >= 60, w)
    _, st = host_tables(tb, 0.0, 1, False)
    assert st[4] > 0
    with pytest.raises(ContrackLibError):
        host_tables(tb, 0.0, 1, False, with_runs=False)


def test_near_tie_on_pole_rows_uses_numpy_summation_order(fixture_cube):
    """overlap placed exactly on a fraction of a class that mixes a pole row with ordinary rows (contrack.py:721-742)."""
    from _common import pole_tie_overlaps
    a, lat, lon = fixture_cube
    w = row_weights(lat, lon)
    ties = 0
    for thr in (100, 150):
        tb = build_tables(a >= thr, w)
        vals = pole_tie_overlaps(a, lat, lon, thr)
        assert vals
        for ov in vals:
            for two in (True, False):
                f, st = host_tables(tb, ov, 1, two)
                assert np.array_equal(f, oracle.run_contrack(a, lat, lon, thr, '>=', ov, 1, two)), (thr, ov, two)
                ties += st[5]
    assert ties > 0                          # the exact resolver really ran
    with pytest.raises(Exception):           # and without row-runs the library refuses to guess
        host_tables(build_tables(a >= 100, w), pole_tie_overlaps(a, lat, lon, 100)[0], 1, True, with_runs=False)


def test_label_granular_event_replay(monkeypatch, fixture_cube, golden, reference_run):
    """track_events_fast (the product path's date-line merge + persistence at label granularity, from the event list the
    cooperative kernel emits) through the all-host entry point: same results as the reference, and the per-component replay
    takes over when a label straddles a box."""
    monkeypatch.setenv('CT_TRACK_EVENTS', '1')
    lat, lon = regular_grid(24, 16)
    w = row_weights(lat, lon)
    fast, slow = 0, 0
    for r in reference_run['quirk']:
        x = synth_cube(r['seed'], 12, 24, 16, (1.5, 2, 2))
        f, st = host_tables(build_tables(x >= 60, w), 0.0, r['persistence'], False)
        assert sha_i4(f) == r['sha256'], r['seed']
        fast += st[5] >= 1000000
        slow += st[5] < 1000000
    for seed in SPLIT_SEEDS + list(range(1000, 1060)):
        x = synth_cube(seed, 12, 24, 16, (1.5, 2, 2))
        tb = build_tables(x >= 60, w)
        f, st = host_tables(tb, 0.0, 1, False)
        assert np.array_equal(f, oracle.track_persistence((x >= 60).astype(int), 1)), seed
        fast += st[5] >= 1000000
        slow += st[5] < 1000000
        f, st = host_tables(tb, 0.5, 2, True)
        assert np.array_equal(f, oracle.run_contrack(x, lat, lon, 60, '>=', 0.5, 2, True)), seed
    assert fast > 10 and slow > 0, (fast, slow)
    a, la, lo = fixture_cube
    for r in golden['fixture']:
        f, st = host_tables(build_tables(a >= r['threshold'], row_weights(la, lo)), r['overlap'], r['persistence'],
                            r['twosided'])
        assert sha_i4(f) == r['sha256'] and st[0] == len(r['ids'])
