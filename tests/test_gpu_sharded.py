"""Time-sharded path on the GPU.  The single-process variant (several contexts on cuda:0) runs on one B200 and covers
the ct_shard_* kernels and plumbing; the torchrun variant needs >= 2 GPUs and is skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import contrack_oracle as oracle
from _common import row_weights
from _synth import synth_cube, regular_grid

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_local(x, lat, lon, w, parts, thr, gorl, ov, pers, two):
    import torch
    from contrack_b200 import Engine, sharded
    from contrack_b200._lib import GORL_TO_OP
    engines = [Engine(0) for _ in parts]
    try:
        bounds = np.cumsum([0] + list(parts))
        xs = [torch.from_numpy(np.ascontiguousarray(x[a:b])).cuda() for a, b in zip(bounds[:-1], bounds[1:])]
        outs, n, info = sharded.run_contrack_sharded_local(engines, xs, x.shape[0], w, thr, True, GORL_TO_OP[gorl], ov,
                                                           pers, two)
        torch.cuda.synchronize()
        return np.concatenate([o.cpu().numpy() for o in outs]), n, info
    finally:
        for e in engines:
            e.handle and e.lib.ct_destroy(e.handle)
            e.handle = None


def run_local_dev(x, w, parts, thr, gorl, ov, pers, two, opts=None):
    """Device-table variant (the one bench.py --gpus N uses): tables gathered device to device, merge kernel, global phase
    on the device in a second context."""
    import torch
    from contrack_b200 import Engine, sharded
    from contrack_b200._lib import GORL_TO_OP
    engines = [Engine(0) for _ in parts]
    try:
        for k, v in (opts or {}).items():
            for e in engines + [sharded._global_engine(engines[0])]:      # the global phase runs in engines[0]._global
                e.set_option(k, v)
        bounds = np.cumsum([0] + list(parts))
        xs = [torch.from_numpy(np.ascontiguousarray(x[a:b])).cuda() for a, b in zip(bounds[:-1], bounds[1:])]
        outs, n, info = sharded.run_contrack_sharded_local_dev(engines, xs, x.shape[0], w, thr, True, GORL_TO_OP[gorl],
                                                               ov, pers, two)
        torch.cuda.synchronize()
        return np.concatenate([o.cpu().numpy() for o in outs]), n, info
    finally:
        for e in engines:
            g = getattr(e, '_global', None)
            for h in (e, g):
                if h is not None and h.handle:
                    h.lib.ct_destroy(h.handle)
                    h.handle = None


@pytest.mark.parametrize('parts', [(6, 5), (4, 3, 4), (1, 9, 1), (2, 2, 2, 2, 3), (11,)])
def test_fixture_sharded_device_tables(fixture_cube, reference_run, parts):
    a, lat, lon = fixture_cube
    w = row_weights(lat, lon)
    from _common import sha_i4
    for r in reference_run['fixture']:
        f, n, _ = run_local_dev(a, w, parts, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(f) == r['sha256'], (parts, r['key'])
        assert n == len(r['ids'])


def test_sharded_device_tables_quirks_and_synthetic(reference_run):
    from _common import sha_i4
    for r in reference_run['quirk'] + reference_run['synthetic']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        lat, lon = regular_grid(H, W)
        w = row_weights(lat, lon)
        thr = np.float64(r['threshold']) if r.get('threshold_is_np_float64') else r['threshold']
        f32 = not r.get('threshold_is_np_float64')
        for parts in [(T // 2, T - T // 2), (T // 3, T // 3, T - 2 * (T // 3))]:
            import torch
            from contrack_b200 import Engine, sharded
            from contrack_b200._lib import GORL_TO_OP
            engines = [Engine(0) for _ in parts]
            bounds = np.cumsum([0] + list(parts))
            xs = [torch.from_numpy(np.ascontiguousarray(x[a:b])).cuda() for a, b in zip(bounds[:-1], bounds[1:])]
            outs, n, _ = sharded.run_contrack_sharded_local_dev(engines, xs, T, w, thr, f32, GORL_TO_OP[r['gorl']],
                                                                r['overlap'], r['persistence'], r['twosided'])
            f = np.concatenate([o.cpu().numpy() for o in outs])
            for e in engines:
                for h in (e, getattr(e, '_global', None)):
                    if h is not None and h.handle:
                        h.lib.ct_destroy(h.handle); h.handle = None
            assert sha_i4(f) == r['sha256'], (r['seed'], parts)
            assert n == len(r['ids'])


def test_sharded_device_tables_host_fallbacks(fixture_cube):
    """Near-ties on pole rows (exact host resolver, plane runs served through the callback) and the host table path."""
    from _common import pole_tie_overlaps
    a, lat, lon = fixture_cube
    w = row_weights(lat, lon)
    ovs = pole_tie_overlaps(a, lat, lon, 150)[:6] or [0.5]
    for ov in ovs:
        ref = oracle.run_contrack(a, lat, lon, 150, '>=', ov, 2, True)
        f, n, _ = run_local_dev(a, w, (4, 4, 3), 150, '>=', ov, 2, True)
        assert np.array_equal(f, ref), ov
    ref = oracle.run_contrack(a, lat, lon, 150, '>=', 0.5, 5, True)
    f, n, _ = run_local_dev(a, w, (6, 5), 150, '>=', 0.5, 5, True, opts={'gpu_tables': 0})
    assert np.array_equal(f, ref)


@pytest.mark.parametrize('parts', [(6, 5), (4, 3, 4), (1, 9, 1), (2, 2, 2, 2, 3)])
def test_fixture_sharded_on_one_gpu(fixture_cube, golden, parts):
    a, lat, lon = fixture_cube
    w = row_weights(lat, lon)
    for r in golden['fixture']:
        f, n, _ = run_local(a, lat, lon, w, parts, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        ref = oracle.run_contrack(a, lat, lon, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert np.array_equal(f, ref), (parts, r['key'])
        assert n == len(r['ids'])


def test_sharded_stale_box_split_and_seam_cases():
    lat, lon = regular_grid(24, 16)
    w = row_weights(lat, lon)
    for seed in [1396, 1933, 1003, 1011]:
        x = synth_cube(seed, 12, 24, 16, (1.5, 2, 2))
        for parts in [(6, 6), (5, 4, 3)]:
            f, _, info = run_local(x, lat, lon, w, parts, 60, '>=', 0.0, 1, False)
            assert np.array_equal(f, oracle.track_persistence((x >= 60).astype(int), 1)), (seed, parts)
            f, _, _ = run_local(x, lat, lon, w, parts, 60, '>=', 0.5, 2, True)
            assert np.array_equal(f, oracle.run_contrack(x, lat, lon, 60, '>=', 0.5, 2, True)), (seed, parts)


def test_sharded_benchmark_grid_with_poles():
    x = synth_cube(2, 12, 721, 1440, (2.5, 24, 40))
    lat = np.linspace(90, -90, 721).astype(np.float32)
    lon = (np.arange(1440) * 0.25).astype(np.float32)
    ref = oracle.run_contrack(x, lat, lon, 160, '>=', 0.5, 5, True, force=True)
    w = oracle.weight_grid(lat, oracle.resolution(lat, True), oracle.resolution(lon, True), 1440)[:, 0].copy()
    f, n, _ = run_local(x, lat, lon, w, (5, 4, 3), 160, '>=', 0.5, 5, True)
    assert np.array_equal(f, ref)
    f, n, _ = run_local_dev(x, w, (5, 4, 3), 160, '>=', 0.5, 5, True)
    assert np.array_equal(f, ref)


def test_sharded_torchrun_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29611', os.path.join(ROOT, 'tests', '_shard_gpu_worker.py')]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(' ok') >= 2


def test_sharded_device_tables_empty_shards(fixture_cube):
    """A rank whose planes hold no component at all, components only on one side of a cut, and an empty cube."""
    a, lat, lon = fixture_cube
    w = row_weights(lat, lon)
    x = a.copy()
    x[:5] = -1000.0                                   # the first shard of (5, 6) is empty
    for parts in [(5, 6), (3, 2, 6), (4, 7)]:
        f, n, _ = run_local_dev(x, w, parts, 150, '>=', 0.5, 2, True)
        ref = oracle.run_contrack(x, lat, lon, 150, '>=', 0.5, 2, True)
        assert np.array_equal(f, ref) and n == len(np.unique(ref)) - 1, parts
    x[:] = -1000.0
    f, n, _ = run_local_dev(x, w, (4, 7), 150, '>=', 0.5, 2, True)
    assert n == 0 and not f.any()
    x[:] = 1000.0                                     # one component per plane covering everything
    f, n, _ = run_local_dev(x, w, (6, 5), 150, '>=', 0.5, 2, True)
    assert np.array_equal(f, oracle.run_contrack(x, lat, lon, 150, '>=', 0.5, 2, True))
