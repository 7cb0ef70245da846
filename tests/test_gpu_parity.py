"""Parity of the CUDA path (through the C-ABI library) with the oracle -- needs a B200 (`-m gpu`)."""
import os

import numpy as np
import pytest

from oracle import contrack_oracle as oracle
from _common import row_weights, sha_i4, same_partition
from _synth import synth_cube, regular_grid

pytestmark = pytest.mark.gpu

OPS = {'>=': 0, 'ge': 0, '<=': 1, 'le': 1, '>': 2, 'gt': 2, '<': 3, 'lt': 3}


@pytest.fixture(scope='module')
def eng():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from contrack_b200 import Engine
    return Engine.get(0)


def gpu_run(eng, x, lat, lon, thr, gorl, ov, pers, two, stage=0, thr_is_f32=True):
    import torch
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    flag, n = eng.run_contrack(xd, row_weights(lat, lon), thr, thr_is_f32, OPS[gorl], ov, pers, two, stage=stage)
    torch.cuda.synchronize()
    return flag.cpu().numpy(), n


@pytest.mark.parametrize('thr,ov,pers,two', [(150, .5, 5, False), (150, .5, 5, True), (160, .5, 5, True),
                                             (100, .7, 3, True)])
def test_fixture_bit_exact(eng, fixture_cube, golden, thr, ov, pers, two):
    a, lat, lon = fixture_cube
    f, n = gpu_run(eng, a, lat, lon, thr, '>=', ov, pers, two)
    key = 'thr%d_ov%02d_p%d_%s' % (thr, int(ov * 10), pers, 'two' if two else 'one')
    want = [r for r in golden['fixture'] if r['key'] == key][0]
    assert f.dtype == np.int32 and f.shape == a.shape
    assert sha_i4(f) == want['sha256']
    assert n == len(want['ids'])
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'flags_fixture.npz'))[key]
    assert np.array_equal(f, gold)


def test_fixture_stages(eng, fixture_cube):
    a, lat, lon = fixture_cube
    st = {}
    oracle.run_contrack(a, lat, lon, 150, '>=', .5, 5, True, stages=st)
    f1, _ = gpu_run(eng, a, lat, lon, 150, '>=', .5, 5, True, stage=1)
    assert same_partition(f1, st['label2d'])
    f2, _ = gpu_run(eng, a, lat, lon, 150, '>=', .5, 5, True, stage=2)
    assert same_partition(f2, st['label2d_seam'])
    f3, _ = gpu_run(eng, a, lat, lon, 150, '>=', .5, 5, True, stage=3)
    assert np.array_equal(f3 > 0, st['filtered'] > 0)
    f4, _ = gpu_run(eng, a, lat, lon, 150, '>=', .5, 5, True, stage=4)
    assert np.array_equal(f4, st['label3d'])


def test_synthetic_golden(eng, golden):
    for r in golden['synthetic']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        lat, lon = regular_grid(H, W)
        f, _ = gpu_run(eng, x, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(f) == r['sha256'], r


def test_stale_box_quirk_vectors(eng, golden):
    lat, lon = regular_grid(24, 16)
    for r in golden['quirk']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        f, _ = gpu_run(eng, x, lat, lon, r['threshold'], '>=', 0.0, r['persistence'], False)
        assert sha_i4(f) == r['sha256'], r['seed']


def test_stale_box_splits(eng):
    lat, lon = regular_grid(24, 16)
    for seed in [1396, 1933, 2136, 2257]:
        x = synth_cube(seed, 12, 24, 16, (1.5, 2, 2))
        f, _ = gpu_run(eng, x, lat, lon, 60, '>=', 0.0, 1, False)
        assert eng.stats()['seam_splits'] > 0
        assert np.array_equal(f, oracle.track_persistence((x >= 60).astype(int), 1)), seed


@pytest.mark.parametrize('shape', [(12, 24, 16), (9, 33, 47), (7, 40, 130), (6, 19, 1031), (5, 64, 2052), (3, 1, 5),
                                   (1, 8, 8), (2, 5, 1)])
def test_random_cubes_odd_shapes(eng, shape):
    T, H, W = shape
    for seed in range(3):
        x = synth_cube(100 * T + seed, T, H, W, (1.0, min(2, H / 4), min(3, W / 4)))
        lat = (60 - np.arange(H) * 1.5).astype(np.float32) if H > 1 else np.array([10.0, 8.5], np.float32)[:H]
        lon = (np.arange(W) * 0.25).astype(np.float32)
        if H < 2 or W < 2:
            # the reference cannot derive a resolution from a single coordinate: feed the weights directly
            w = np.ones(H)
            import torch
            flag, _ = eng.run_contrack(torch.from_numpy(x).cuda(), w, 50, True, 0, 0.5, 1, True)
            assert flag.shape == (T, H, W)
            continue
        for two, gorl, thr in [(True, '>=', 50), (False, '<', -40)]:
            ref = oracle.run_contrack(x, lat, lon, thr, gorl, 0.4, 2, two)
            f, _ = gpu_run(eng, x, lat, lon, thr, gorl, 0.4, 2, two)
            assert np.array_equal(f, ref), (shape, seed, two)


def test_ops_and_compare_precision(eng):
    x = synth_cube(7, 8, 30, 64, (1.0, 2, 3))
    lat, lon = regular_grid(30, 64)
    v = float(x[3, 10, 20])                                      # a threshold that is an actual data value
    for gorl in ['>=', 'ge', '<=', 'le', '>', 'gt', '<', 'lt']:
        ref = oracle.run_contrack(x, lat, lon, v, gorl, 0.3, 1, True)
        f, _ = gpu_run(eng, x, lat, lon, v, gorl, 0.3, 1, True)
        assert np.array_equal(f, ref), gorl
    thr = 150.7                                                  # float32(150.7) > 150.7: float32 vs float64 compare differ
    x2 = x.copy()
    x2[2, 5:9, 5:9] = np.float32(150.7)
    ref32 = oracle.run_contrack(x2, lat, lon, thr, '>=', 0.0, 1, False)
    ref64 = oracle.run_contrack(x2, lat, lon, np.float64(thr), '<=', 0.0, 1, False)
    f32, _ = gpu_run(eng, x2, lat, lon, thr, '>=', 0.0, 1, False, thr_is_f32=True)
    f64, _ = gpu_run(eng, x2, lat, lon, thr, '<=', 0.0, 1, False, thr_is_f32=False)
    assert np.array_equal(f32, ref32) and np.array_equal(f64, ref64)


def test_float64_input_and_nan(eng):
    x = synth_cube(11, 10, 24, 48, (1.0, 2, 3)).astype(np.float64)
    x[4, 3:6, :] = np.nan
    lat, lon = regular_grid(24, 48)
    ref = oracle.run_contrack(x, lat, lon, 40.0, '>=', 0.3, 2, True)
    f, _ = gpu_run(eng, x, lat, lon, 40.0, '>=', 0.3, 2, True)
    assert np.array_equal(f, ref)


def test_per_timestep_thresholds(eng):
    x = synth_cube(13, 10, 24, 48, (1.0, 2, 3))
    lat, lon = regular_grid(24, 48)
    thr = np.linspace(20, 80, 10)
    ref = oracle.run_contrack(x, lat, lon, thr, '>=', 0.3, 2, True)
    f, _ = gpu_run(eng, x, lat, lon, thr, '>=', 0.3, 2, True, thr_is_f32=False)
    assert np.array_equal(f, ref)


def test_empty_and_full_masks(eng):
    lat, lon = regular_grid(24, 48)
    x = np.zeros((6, 24, 48), np.float32)
    f, n = gpu_run(eng, x, lat, lon, 1.0, '>=', 0.5, 1, True)
    assert n == 0 and not f.any()
    ref = oracle.run_contrack(x, lat, lon, -1.0, '>=', 0.5, 1, True)
    f, n = gpu_run(eng, x, lat, lon, -1.0, '>=', 0.5, 1, True)
    assert np.array_equal(f, ref) and n == 1


def test_host_buffers_match_device(eng):
    x = synth_cube(5, 24, 181, 360, (2.0, 4, 6))
    lat, lon = regular_grid(181, 360)
    w = row_weights(lat, lon)
    ref = oracle.run_contrack(x, lat, lon, 150, 'ge', 0.7, 4, True)
    f, n = eng.run_contrack(x, w, 150, True, 0, 0.7, 4, True, chunk_planes=5)      # numpy in -> numpy out, 5 chunks
    assert isinstance(f, np.ndarray) and np.array_equal(f, ref)
    assert n == len(np.unique(ref)) - 1


@pytest.mark.parametrize('sparse,threads', [(1, 0), (1, 1), (1, 3), (0, 0)])
def test_host_buffer_result_formats(eng, sparse, threads):
    """Host-buffer entry point: the flag cube crosses PCIe either as the row-run table that host threads expand
    (host_sparse=1, any thread count) or as the dense cube (host_sparse=0); bytes must be identical, including the pieces
    of components that the stale-box date-line merge splits (overrides), and the result buffer may hold garbage on entry."""
    eng.set_option('host_sparse', sparse)
    eng.set_option('host_threads', threads)
    try:
        x = synth_cube(9, 17, 181, 360, (1.5, 4, 6))
        lat, lon = regular_grid(181, 360)
        ref = oracle.run_contrack(x, lat, lon, 120, '>=', 0.3, 3, True)
        out = np.full(x.shape, -7, np.int32)
        f, n = eng.run_contrack(x, row_weights(lat, lon), 120, True, 0, 0.3, 3, True, out=out, chunk_planes=4)
        assert f is out and np.array_equal(out, ref) and n == len(np.unique(ref)) - 1
        st = eng.stats()
        assert st['host_sparse'] == sparse and st['h2d_bytes'] == x.nbytes
        assert st['d2h_bytes'] == (12 * st['runs'] if sparse else ref.size * 4)
        la, lo = regular_grid(24, 16)
        for seed in [1396, 1933]:
            xs = synth_cube(seed, 12, 24, 16, (1.5, 2, 2))
            out = np.full(xs.shape, 99, np.int32)
            eng.run_contrack(xs, row_weights(la, lo), 60, True, 0, 0.0, 1, False, out=out)
            assert eng.stats()['seam_splits'] > 0
            assert np.array_equal(out, oracle.track_persistence((xs >= 60).astype(int), 1)), seed
        z = np.zeros((3, 8, 8), np.float32)                       # no runs at all
        out = np.full(z.shape, 5, np.int32)
        eng.run_contrack(z, np.ones(8), 1.0, True, 0, 0.5, 1, True, out=out)
        assert not out.any()
    finally:
        eng.set_option('host_sparse', 1)
        eng.set_option('host_threads', 0)


def test_pole_rows_are_special(eng, fixture_cube):
    a, lat, lon = fixture_cube
    gpu_run(eng, a, lat, lon, 150, '>=', .5, 5, True)
    assert eng.stats()['special_rows'] == 2        # cos(float32(pi/2)) rows: not exactly summable with the rest


def test_medium_cube_against_oracle(eng):
    # 0.25-degree rows (721 x 1440) for a few steps: the benchmark grid, oracle still finishes in seconds
    x = synth_cube(2, 12, 721, 1440, (2.5, 24, 40))
    lat = np.linspace(90, -90, 721).astype(np.float32)
    lon = (np.arange(1440) * 0.25).astype(np.float32)
    ref = oracle.run_contrack(x, lat, lon, 160, '>=', 0.5, 5, True, force=True)
    w = oracle.weight_grid(lat, oracle.resolution(lat, True), oracle.resolution(lon, True), 1440)[:, 0].copy()
    import torch
    f, n = eng.run_contrack(torch.from_numpy(x).cuda(), w, 160, True, 0, 0.5, 5, True)
    assert np.array_equal(f.cpu().numpy(), ref)


def test_bulk_copy_threshold_variant(eng, fixture_cube, golden):
    """The cp.async.bulk staged threshold kernel (option tma=1) gives the same bits as the plain-load one."""
    a, lat, lon = fixture_cube
    eng.set_option('tma', 1)
    try:
        for r in golden['fixture']:
            f, _ = gpu_run(eng, a, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
            assert sha_i4(f) == r['sha256'], r['key']
        x = synth_cube(2, 12, 721, 1440, (2.5, 24, 40))
        la = np.linspace(90, -90, 721).astype(np.float32)
        lo = (np.arange(1440) * 0.25).astype(np.float32)
        ref = oracle.run_contrack(x, la, lo, 160, '>=', 0.5, 5, True, force=True)
        w = oracle.weight_grid(la, oracle.resolution(la, True), oracle.resolution(lo, True), 1440)[:, 0].copy()
        import torch
        f, n = eng.run_contrack(torch.from_numpy(x).cuda(), w, 160, True, 0, 0.5, 5, True)
        assert np.array_equal(f.cpu().numpy(), ref)
        xd = synth_cube(11, 10, 24, 48, (1.0, 2, 3)).astype(np.float64)      # float64 rows, 384 bytes each
        la2, lo2 = regular_grid(24, 48)
        ref = oracle.run_contrack(xd, la2, lo2, 40.0, '<', 0.3, 2, True)
        f, _ = gpu_run(eng, xd, la2, lo2, 40.0, '<', 0.3, 2, True)
        assert np.array_equal(f, ref)
    finally:
        eng.set_option('tma', 1)


def test_dense_paint_path(eng, fixture_cube, golden):
    """overlap_zero=0: one dense paint kernel instead of zero fill + sparse paint."""
    a, lat, lon = fixture_cube
    eng.set_option('overlap_zero', 0)
    try:
        r = golden['fixture'][0]
        f, _ = gpu_run(eng, a, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(f) == r['sha256']
    finally:
        eng.set_option('overlap_zero', 1)


def test_near_tie_on_pole_rows_falls_back_to_exact_host_sums(eng, fixture_cube):
    from _common import pole_tie_overlaps
    a, lat, lon = fixture_cube
    flagged = 0
    for thr in (100, 150):
        for ov in pole_tie_overlaps(a, lat, lon, thr):
            for two in (True, False):
                f, _ = gpu_run(eng, a, lat, lon, thr, '>=', ov, 1, two)
                assert np.array_equal(f, oracle.run_contrack(a, lat, lon, thr, '>=', ov, 1, two)), (thr, ov, two)
                flagged += eng.stats().get('neartie_flagged', 0)
    assert flagged > 0


def test_host_table_path_matches_device_table_path(eng, golden):
    """gpu_tables=0: the ordered phase runs entirely on the host (the path the sharded run uses)."""
    r = golden['synthetic'][0]
    T, H, W = r['shape']
    x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
    lat, lon = regular_grid(H, W)
    eng.set_option('gpu_tables', 0)
    try:
        f, _ = gpu_run(eng, x, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(f) == r['sha256']
        assert 'sweeps' not in eng.stats()
    finally:
        eng.set_option('gpu_tables', 1)
    f, _ = gpu_run(eng, x, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
    assert sha_i4(f) == r['sha256'] and eng.stats()['sweeps'] >= 1


@pytest.mark.parametrize('plane', [0, 1])
@pytest.mark.parametrize('chunks', [2, 5, 12])
def test_chunked_table_pipeline(eng, chunks, plane):
    """The table kernels (global-memory ones, or the plane kernel) run per time chunk while later chunks are still being
    thresholded; every chunking must give the bytes of the unchunked run: date-line-heavy cubes with stale-box splits,
    per-timestep thresholds, one-sided mode, and the intermediate stages."""
    eng.set_option('chunks', chunks)
    eng.set_option('fast_chunks', chunks)
    eng.set_option('plane_kernel', plane)
    eng.set_option('chunk_min_planes', 1)
    try:
        la, lo = regular_grid(24, 16)
        for seed in [1396, 1933, 2136, 1011, 1137]:
            xs = synth_cube(seed, 12, 24, 16, (1.5, 2, 2))
            f, _ = gpu_run(eng, xs, la, lo, 60, '>=', 0.0, 1, False)
            assert 2 <= eng.stats()['chunks'] <= chunks
            assert np.array_equal(f, oracle.track_persistence((xs >= 60).astype(int), 1)), seed
        x = synth_cube(11, 23, 91, 180, (1.5, 3, 5))
        lat, lon = regular_grid(91, 180)
        for two, gorl, thr, ov in [(True, '>=', 90, 0.5), (False, '<=', -80, 0.3)]:
            ref = oracle.run_contrack(x, lat, lon, thr, gorl, ov, 3, two)
            f, n = gpu_run(eng, x, lat, lon, thr, gorl, ov, 3, two)
            assert np.array_equal(f, ref) and n == len(np.unique(ref)) - 1, (chunks, two)
        st = {}
        oracle.run_contrack(x, lat, lon, 90, '>=', 0.5, 3, True, stages=st)
        f1, _ = gpu_run(eng, x, lat, lon, 90, '>=', 0.5, 3, True, stage=1)
        assert same_partition(f1, st['label2d'])
        f2, _ = gpu_run(eng, x, lat, lon, 90, '>=', 0.5, 3, True, stage=2)
        assert same_partition(f2, st['label2d_seam'])
        f4, _ = gpu_run(eng, x, lat, lon, 90, '>=', 0.5, 3, True, stage=4)
        assert np.array_equal(f4, st['label3d'])
    finally:
        eng.set_option('chunks', 4)
        eng.set_option('fast_chunks', 1)
        eng.set_option('plane_kernel', 2)
        eng.set_option('chunk_min_planes', 1024)


@pytest.mark.parametrize('opts', [{'tma': 0}, {'overlap_zero': 0}, {'gpu_tables': 0}, {'chunks': 4, 'chunk_min_planes': 1, 'plane_kernel': 0},
                                  {'chunks': 64, 'chunk_min_planes': 1, 'gpu_tables': 0},
                                  {'chunks': 3, 'chunk_min_planes': 2, 'tma': 0, 'plane_kernel': 0}, {'fill_ctas': 0},
                                  {'fill_late': 1, 'plane_kernel': 1}])
def test_kernel_variants_give_identical_results(eng, fixture_cube, golden, opts):
    """Every selectable path (plain-load threshold kernel, dense paint, host table phase, chunked global-memory table
    kernels, uncapped / late zero fill) must produce the same
    bytes."""
    a, lat, lon = fixture_cube
    defaults = {'tma': 1, 'overlap_zero': 1, 'gpu_tables': 1, 'chunks': 4, 'chunk_min_planes': 1024, 'plane_kernel': 2,
                'fill_ctas': 2, 'fill_late': 0}
    for k, v in opts.items():
        eng.set_option(k, v)
    try:
        for r in golden['fixture'][:2]:
            f, _ = gpu_run(eng, a, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
            assert sha_i4(f) == r['sha256'], (opts, r['key'])
        x = synth_cube(3, 6, 721, 1440, (2.0, 12, 20))
        la = np.linspace(90, -90, 721).astype(np.float32)
        lo = (np.arange(1440) * 0.25).astype(np.float32)
        ref = oracle.run_contrack(x, la, lo, 150, '>=', 0.5, 2, True, force=True)
        w = oracle.weight_grid(la, oracle.resolution(la, True), oracle.resolution(lo, True), 1440)[:, 0].copy()
        import torch
        f, n = eng.run_contrack(torch.from_numpy(x).cuda(), w, 150, True, 0, 0.5, 2, True)
        assert np.array_equal(f.cpu().numpy(), ref), opts
    finally:
        for k in opts:
            eng.set_option(k, defaults[k])


def test_event_replay_matches_component_replay(eng, reference_run):
    """Steps 4c/4d (date-line merge through stale boxes + persistence): the label-granular event replay (default) and the
    all-host per-component replay ("gpu_tables" = 0) must agree, on the reference's own outputs for the stale-box cubes and
    on random cubes with many date-line events; the per-component replay must still be reached from the default path when a
    label straddles a stale box."""
    import torch
    fast_used, fallback_used = 0, 0
    cases = [(r['seed'], tuple(r['shape']), tuple(r['sigma']), r['threshold'], r['sha256']) for r in reference_run['quirk']]
    cases += [(s, (16, 24, 16), (1.5, 2, 2), 60, None) for s in range(2000, 2040)]
    for seed, shape, sigma, thr, want in cases:
        x = synth_cube(seed, *shape, sigma)
        lat, lon = regular_grid(shape[1], shape[2])
        w = row_weights(lat, lon)
        xd = torch.from_numpy(x).cuda()
        out = {}
        for gt in (1, 0):
            eng.set_option('gpu_tables', gt)
            try:
                f, n = eng.run_contrack(xd, w, thr, True, 0, 0.0, 1, False)
            finally:
                eng.set_option('gpu_tables', 1)
            out[gt] = (f.cpu().numpy(), n)
            if gt == 1:
                used = eng.stats().get('label_fast', -1)
                fast_used += used == 1
                fallback_used += used == 0
        assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1], seed
        if want is not None:
            assert sha_i4(out[1][0]) == want, seed
        else:
            assert np.array_equal(out[1][0], oracle.run_contrack(x, lat, lon, thr, '>=', 0.0, 1, False)), seed
    assert fast_used > 0 and fallback_used > 0, (fast_used, fallback_used)


@pytest.mark.parametrize('tma', [1, 0])
@pytest.mark.parametrize('W', [1440, 1024, 2048, 96, 33, 1028, 36])
def test_runs_from_threshold_kernel_edge_cases(eng, W, tma):
    """Row-runs come out of the threshold kernel (8 slots per row): rows with more runs than slots (noise), runs that end
    exactly at the row end when the row fills whole 32-word groups (W = 1024, 2048), full rows, alternating cells."""
    import torch
    T, H = 5, 40
    rng = np.random.default_rng(W)
    x = rng.standard_normal((T, H, W)).astype(np.float32) * 100           # white noise: hundreds of runs per row
    x[1, 3, :] = 500                                                       # a full row
    x[1, 5, W - 7:] = 500                                                  # run that reaches the row end
    x[2, 7, ::2] = 500; x[2, 7, 1::2] = -500                               # alternating cells: W / 2 runs
    x[3, :, :] = -500                                                      # empty plane ...
    x[3, 10:14, W - 40 if W > 40 else 0:] = 500                            # ... with one block touching the right border
    x[3, 10:14, :3] = 500                                                  # and the left one (date-line class)
    lat, lon = regular_grid(H, W)
    w = row_weights(lat, lon)
    ref = oracle.run_contrack(x, lat, lon, 120, '>=', 0.2, 1, True)
    eng.set_option('tma', tma)
    try:
        f, n = eng.run_contrack(torch.from_numpy(x).cuda(), w, 120, True, 0, 0.2, 1, True)
        st = eng.stats()
        eng.set_option('plane_kernel', 0)                                  # the global-memory table kernels read the same slots
        f0, n0 = eng.run_contrack(torch.from_numpy(x).cuda(), w, 120, True, 0, 0.2, 1, True)
    finally:
        eng.set_option('tma', 1); eng.set_option('plane_kernel', 2)
    assert np.array_equal(f.cpu().numpy(), ref) and n == len(np.unique(ref)) - 1
    assert np.array_equal(f0.cpu().numpy(), ref)
    if W >= 96:
        assert st.get('slot_overflow', 0) == 1
