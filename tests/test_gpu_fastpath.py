"""The fast table pipeline (one thread block per plane: shared-memory union-find, look-back numbering, pair hash; one
cooperative kernel for steps 3 / 4a / 4b; event replay on the host) against the oracle and against the global-memory table
kernels it falls back to -- needs a B200 (`-m gpu`)."""
import numpy as np
import pytest

from oracle import contrack_oracle as oracle
from _common import row_weights, sha_i4
from _synth import synth_cube, regular_grid
from test_gpu_parity import gpu_run, OPS

pytestmark = pytest.mark.gpu

DEFAULTS = {'plane_kernel': 2, 'max_sweeps': 32, 'plane_smem': 0, 'fast_chunks': 1, 'chunk_min_planes': 1024}


@pytest.fixture()
def eng():
    import torch
    assert torch.cuda.is_available()
    from contrack_b200 import Engine
    e = Engine.get(0)
    yield e
    for k, v in DEFAULTS.items():
        e.set_option(k, v)


def alternating_chain(T, W=1440, H=9):
    """One single-row contour per plane, growing by three cells per step and overlapping its successor by just over half of
    its own length: forward overlap > 0.5, backward overlap < 0.5 whenever the predecessor is kept.  With overlap = 0.5 (twosided)
    the verdicts alternate kept / killed / kept ... along the whole cube: every verdict depends on the one before it, the
    worst case for Jacobi sweeps (contrack.py:706-742 is sequential in time)."""
    x = np.zeros((T, H, W), np.float32)
    a, L = W // 2, 6
    for t in range(T):
        x[t, H // 2, a:a + L] = 100.0
        o = L // 2 + 1
        L2 = L + 3
        a = a + L - o if t % 2 == 0 else a + o - L2             # zigzag: stays within ~2 L of the start
        L = L2
        assert 0 < a and a + L < W
    return x


def test_default_path_is_the_fast_path(eng, fixture_cube, golden):
    a, lat, lon = fixture_cube
    for r in golden['fixture']:
        f, n = gpu_run(eng, a, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(f) == r['sha256'] and n == len(r['ids'])
        st = eng.stats()
        assert st['fast_path'] == 1.0 and st['plane_attempts'] == 1.0
        assert st['kernel_launches'] <= 16, st['kernel_launches']


@pytest.mark.parametrize('opts', [{'plane_kernel': 0}, {'max_sweeps': 1}, {'max_sweeps': 2, 'fast_chunks': 3, 'chunk_min_planes': 2},
                                  {'plane_smem': 200 * 1024}, {'plane_smem': 8 * 1024}])
def test_fast_path_variants_and_fallbacks(eng, fixture_cube, golden, reference_run, opts):
    """global-memory table kernels (plane_kernel = 0), the plane-ordered wavefront after one / two Jacobi sweeps, the large
    shared-memory configuration and a very small one (the quirk cubes' planes still fit; planes that do not are the subject of
    test_noisy_planes_fall_back_and_dense_tables_grow)."""
    for k, v in opts.items():
        eng.set_option(k, v)
    a, lat, lon = fixture_cube
    for r in golden['fixture']:
        f, n = gpu_run(eng, a, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(f) == r['sha256'] and n == len(r['ids']), (opts, r['key'])
    st = eng.stats()
    assert st['fast_path'] == (0.0 if 'plane_kernel' in opts else 1.0)
    if 'max_sweeps' in opts:
        assert st['wavefront_planes'] > 0
    for r in reference_run['quirk'] + reference_run['synthetic'][:4]:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        la, lo = regular_grid(H, W)
        thr = np.float64(r['threshold']) if r.get('threshold_is_np_float64') else r['threshold']
        f, n = gpu_run(eng, x, la, lo, thr, r['gorl'], r['overlap'], r['persistence'], r['twosided'],
                       thr_is_f32=not r.get('threshold_is_np_float64'))
        assert sha_i4(f) == r['sha256'] and n == len(r['ids']), (opts, r['seed'])


@pytest.mark.parametrize('max_sweeps', [32, 4])
def test_alternating_keep_kill_chain(eng, max_sweeps):
    eng.set_option('max_sweeps', max_sweeps)
    T = 120
    x = alternating_chain(T)
    lat = (8 - 2.0 * np.arange(9)).astype(np.float32)
    lon = (np.arange(1440) * 0.25).astype(np.float32)
    st = {}
    ref = oracle.run_contrack(x, lat, lon, 50, '>=', 0.5, 1, True, stages=st)
    kept = [bool(st['filtered'][t].any()) for t in range(T)]
    assert kept[:6] == [True, False, True, False, True, False]             # the chain really alternates
    f, n = gpu_run(eng, x, lat, lon, 50, '>=', 0.5, 1, True)
    assert np.array_equal(f, ref) and n == len(np.unique(ref)) - 1
    s = eng.stats()
    assert s['fast_path'] == 1.0 and s['wavefront_planes'] > 0 and s['sweeps'] <= max_sweeps + 3, s


def test_noisy_planes_fall_back_and_dense_tables_grow():
    """White noise: ~W/4 runs per row.  (a) A fresh context sized for anomaly-like fields reports the exact totals and is
    rebuilt once (capacity retry); (b) 721 x 1440 noise planes hold far more runs than any shared-memory budget: the
    global-memory kernels take over.  Both must give the oracle's bytes."""
    import torch
    from contrack_b200 import Engine
    rng = np.random.default_rng(5)
    e = Engine(0)
    try:
        x = rng.standard_normal((6, 40, 64)).astype(np.float32)
        lat, lon = regular_grid(40, 64)
        ref = oracle.run_contrack(x, lat, lon, 0.3, '>=', 0.2, 2, True)
        f, n = e.run_contrack(torch.from_numpy(x).cuda(), row_weights(lat, lon), 0.3, True, 0, 0.2, 2, True)
        assert np.array_equal(f.cpu().numpy(), ref)
        assert e.stats()['fast_path'] == 1.0
        big = rng.standard_normal((3, 721, 1440)).astype(np.float32)
        la = np.linspace(90, -90, 721).astype(np.float32)
        lo = (np.arange(1440) * 0.25).astype(np.float32)
        ref = oracle.run_contrack(big, la, lo, 0.5, '>=', 0.1, 1, True, force=True)
        w = oracle.weight_grid(la, oracle.resolution(la, True), oracle.resolution(lo, True), 1440)[:, 0].copy()
        f, n = e.run_contrack(torch.from_numpy(big).cuda(), w, 0.5, True, 0, 0.1, 1, True)
        assert np.array_equal(f.cpu().numpy(), ref)
        assert e.stats()['fast_path'] == 0.0 and e.stats()['plane_attempts'] >= 2.0
        # a later anomaly-like cube on the same context uses the fast path again (the large budget stays selected)
        y = synth_cube(2, 6, 721, 1440, (2.5, 24, 40))
        ref = oracle.run_contrack(y, la, lo, 160, '>=', 0.5, 2, True, force=True)
        f, n = e.run_contrack(torch.from_numpy(y).cuda(), w, 160, True, 0, 0.5, 2, True)
        assert np.array_equal(f.cpu().numpy(), ref) and e.stats()['fast_path'] == 1.0
    finally:
        e.lib.ct_destroy(e.handle)
        e.handle = None


def test_capacity_retry_on_a_fresh_context():
    import torch
    from contrack_b200 import Engine
    rng = np.random.default_rng(7)
    e = Engine(0)
    try:
        # many tiny components: one per run (isolated cells on every other row / column of every other plane), far more
        # components than the first-call estimate (runs / 8) allows for
        x = np.full((5, 64, 128), -1.0, np.float32)
        x[::2, ::2, ::2] = 1.0
        x += 0.01 * rng.standard_normal(x.shape).astype(np.float32)
        lat, lon = regular_grid(64, 128)
        ref = oracle.run_contrack(x, lat, lon, 0.5, '>=', 0.5, 2, True)
        f, n = e.run_contrack(torch.from_numpy(x).cuda(), row_weights(lat, lon), 0.5, True, 0, 0.5, 2, True)
        assert np.array_equal(f.cpu().numpy(), ref) and n == len(np.unique(ref)) - 1
        st = e.stats()
        assert st['fast_path'] == 1.0 and st['plane_attempts'] == 2.0, st
    finally:
        e.lib.ct_destroy(e.handle)
        e.handle = None


def test_stages_and_host_buffers_on_fast_tables(eng, fixture_cube):
    """The debug stages and the host-buffer entry point read the tables the plane kernel built."""
    from _common import same_partition
    a, lat, lon = fixture_cube
    st = {}
    ref = oracle.run_contrack(a, lat, lon, 150, '>=', 0.5, 5, True, stages=st)
    f1, _ = gpu_run(eng, a, lat, lon, 150, '>=', 0.5, 5, True, stage=1)
    assert same_partition(f1, st['label2d'])
    f2, _ = gpu_run(eng, a, lat, lon, 150, '>=', 0.5, 5, True, stage=2)
    assert same_partition(f2, st['label2d_seam'])
    f4, _ = gpu_run(eng, a, lat, lon, 150, '>=', 0.5, 5, True, stage=4)
    assert np.array_equal(f4, st['label3d'])
    fh, n = eng.run_contrack(a, row_weights(lat, lon), 150, True, OPS['>='], 0.5, 5, True)
    assert np.array_equal(fh, ref) and eng.stats()['fast_path'] == 1.0
