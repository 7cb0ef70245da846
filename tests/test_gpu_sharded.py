"""Time-sharded path on the GPU through the ONE C-ABI call per rank (ct_run_contrack_sharded).  On a single B200 the ranks
are host threads with their own contexts and an in-process communicator (same kernels, same merge, same replicated global
phase; only the transport differs); the torchrun variant uses NCCL and needs >= 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import contrack_oracle as oracle
from _common import row_weights, sha_i4
from _synth import synth_cube, regular_grid

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_local(x, w, parts, thr, gorl, ov, pers, two, opts=None, f32=True, engines=None, host=False, comms=None):
    """x split into len(parts) time shards, one in-process rank each (host=True: host buffers through
    ct_run_contrack_sharded_host).  Returns (flag cube, features, [stats])."""
    import torch
    from contrack_b200 import Engine, sharded
    from contrack_b200._lib import GORL_TO_OP
    own = engines is None
    if own:
        engines = [Engine(0) for _ in parts]
    try:
        for k, v in (opts or {}).items():
            for e in engines:
                e.set_option(k, v)
        bounds = np.cumsum([0] + list(parts))
        xs = [np.ascontiguousarray(x[a:b]) for a, b in zip(bounds[:-1], bounds[1:])]
        if not host:
            xs = [torch.from_numpy(a).cuda() for a in xs]
        outs, n, stats = sharded.run_local_group(engines, xs, x.shape[0], w, thr, f32, GORL_TO_OP[gorl], ov, pers, two,
                                                 comms=comms)
        torch.cuda.synchronize()
        return np.concatenate([o if host else o.cpu().numpy() for o in outs]), n, stats
    finally:
        if own:
            for e in engines:
                if e.handle:
                    e.lib.ct_destroy(e.handle)
                    e.handle = None


@pytest.mark.parametrize('parts', [(6, 5), (4, 3, 4), (1, 9, 1), (2, 2, 2, 2, 3), (11,)])
def test_fixture_sharded(fixture_cube, reference_run, parts):
    a, lat, lon = fixture_cube
    w = row_weights(lat, lon)
    for r in reference_run['fixture']:
        f, n, st = run_local(a, w, parts, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'])
        assert sha_i4(f) == r['sha256'], (parts, r['key'])
        assert n == len(r['ids'])
        assert all(s['fast_path'] == 1.0 for s in st)


def test_sharded_quirks_and_synthetic(reference_run):
    """The stale-box quirk cubes need the per-component replay with plane runs served by their owner rank (collective
    fetch); the synthetic ones cover the other operators and float64 thresholds."""
    slow = 0
    for r in reference_run['quirk'] + reference_run['synthetic']:
        T, H, W = r['shape']
        x = synth_cube(r['seed'], T, H, W, tuple(r['sigma']))
        lat, lon = regular_grid(H, W)
        w = row_weights(lat, lon)
        thr = np.float64(r['threshold']) if r.get('threshold_is_np_float64') else r['threshold']
        for parts in [(T // 2, T - T // 2), (T // 3, T // 3, T - 2 * (T // 3))]:
            f, n, st = run_local(x, w, parts, thr, r['gorl'], r['overlap'], r['persistence'], r['twosided'],
                                 f32=not r.get('threshold_is_np_float64'))
            assert sha_i4(f) == r['sha256'], (r['seed'], parts)
            assert n == len(r['ids'])
            slow += st[0]['fast_path'] < 1.0
    assert slow > 0                                  # some quirk cube really took the replay with the collective fetch


@pytest.mark.parametrize('opts', [{'plane_kernel': 0}, {'plane_kernel': 1, 'max_sweeps': 1}, {'plane_kernel': 1, 'plane_smem': 8 * 1024}])
def test_sharded_table_builder_variants(fixture_cube, reference_run, opts):
    a, lat, lon = fixture_cube
    w = row_weights(lat, lon)
    for r in reference_run['fixture'][:2]:
        f, n, _ = run_local(a, w, (4, 3, 4), r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'], opts=opts)
        assert sha_i4(f) == r['sha256'] and n == len(r['ids']), opts
    lat2, lon2 = regular_grid(24, 16)
    for seed in [1396, 1003]:
        x = synth_cube(seed, 12, 24, 16, (1.5, 2, 2))
        f, _, _ = run_local(x, row_weights(lat2, lon2), (5, 4, 3), 60, '>=', 0.5, 2, True, opts=opts)
        assert np.array_equal(f, oracle.run_contrack(x, lat2, lon2, 60, '>=', 0.5, 2, True)), (seed, opts)


def test_sharded_near_ties_use_the_exact_host_resolver(fixture_cube):
    """Near-ties on pole rows: the exact host resolver needs plane runs of planes that live on other ranks."""
    from _common import pole_tie_overlaps
    a, lat, lon = fixture_cube
    w = row_weights(lat, lon)
    ovs = pole_tie_overlaps(a, lat, lon, 150)[:6] or [0.5]
    for ov in ovs:
        ref = oracle.run_contrack(a, lat, lon, 150, '>=', ov, 2, True)
        f, n, _ = run_local(a, w, (4, 4, 3), 150, '>=', ov, 2, True)
        assert np.array_equal(f, ref), ov


def test_sharded_stale_box_split_and_seam_cases():
    lat, lon = regular_grid(24, 16)
    w = row_weights(lat, lon)
    for seed in [1396, 1933, 1003, 1011]:
        x = synth_cube(seed, 12, 24, 16, (1.5, 2, 2))
        for parts in [(6, 6), (5, 4, 3)]:
            f, _, _ = run_local(x, w, parts, 60, '>=', 0.0, 1, False)
            assert np.array_equal(f, oracle.track_persistence((x >= 60).astype(int), 1)), (seed, parts)
            f, _, _ = run_local(x, w, parts, 60, '>=', 0.5, 2, True)
            assert np.array_equal(f, oracle.run_contrack(x, lat, lon, 60, '>=', 0.5, 2, True)), (seed, parts)


@pytest.mark.parametrize('p2p', [1, 0])
def test_sharded_benchmark_grid_with_poles_and_exchange_renegotiation(p2p):
    """721 x 1440 grid with pole rows on three ranks; then, on the SAME contexts and communicators, a cube with many more
    components: the stride negotiated for the first cube is too small, every rank sees it in the headers and the exchange is
    repeated (p2p=1: the peer windows are re-made for the larger stride; p2p=0: the all-gather variant)."""
    from contrack_b200 import Engine, sharded
    lat = np.linspace(90, -90, 721).astype(np.float32)
    lon = (np.arange(1440) * 0.25).astype(np.float32)
    w = oracle.weight_grid(lat, oracle.resolution(lat, True), oracle.resolution(lon, True), 1440)[:, 0].copy()
    engines = [Engine(0) for _ in range(3)]
    comms = sharded.Comm.local_group(3)
    try:
        x = synth_cube(2, 12, 721, 1440, (2.5, 24, 40))
        ref = oracle.run_contrack(x, lat, lon, 160, '>=', 0.5, 5, True, force=True)
        f, n, st = run_local(x, w, (5, 4, 3), 160, '>=', 0.5, 5, True, engines=engines, comms=comms, opts={'p2p': p2p})
        assert np.array_equal(f, ref) and st[0]['shard_attempts'] == 1.0 and st[0]['p2p'] == float(p2p)
        y = synth_cube(3, 12, 721, 1440, (1.0, 3, 4))             # small-scale field: thousands of components per plane
        ref = oracle.run_contrack(y, lat, lon, 100, '>=', 0.3, 2, True, force=True)
        f, n, st = run_local(y, w, (5, 4, 3), 100, '>=', 0.3, 2, True, engines=engines, comms=comms)
        assert np.array_equal(f, ref) and n == len(np.unique(ref)) - 1
        assert st[0]['shard_attempts'] >= 2.0, st[0]
        for seed in (4, 5, 6):                                     # steady state: both halves of the double buffer in use
            z = synth_cube(seed, 12, 721, 1440, (1.0, 3, 4))
            f, n, st = run_local(z, w, (5, 4, 3), 100, '>=', 0.3, 2, True, engines=engines, comms=comms)
            assert np.array_equal(f, oracle.run_contrack(z, lat, lon, 100, '>=', 0.3, 2, True, force=True)), seed
    finally:
        for c in comms:
            c.close()
        for e in engines:
            e.lib.ct_destroy(e.handle)
            e.handle = None


@pytest.mark.parametrize('ctas', [-1, 0, 1, 3])
def test_sharded_result_buffers_full_of_garbage(ctas):
    """Result buffers that arrive full of garbage come back exact, whatever the occupancy cap of the zero fill."""
    import torch
    from contrack_b200 import Engine, sharded
    from contrack_b200._lib import GORL_TO_OP
    lat, lon = regular_grid(91, 180)
    w = row_weights(lat, lon)
    x = synth_cube(11, 13, 91, 180, (1.5, 4, 6))
    ref = oracle.run_contrack(x, lat, lon, 80, '>=', 0.5, 3, True)
    parts = (5, 1, 7)
    engines = [Engine(0) for _ in parts]
    try:
        for e in engines:
            e.set_option('fill_ctas', ctas)
        bounds = np.cumsum([0] + list(parts))
        xs = [torch.from_numpy(np.ascontiguousarray(x[a:b])).cuda() for a, b in zip(bounds[:-1], bounds[1:])]
        outs = [torch.full(tuple(a.shape), -7, dtype=torch.int32, device='cuda') for a in xs]
        res, n, _ = sharded.run_local_group(engines, xs, x.shape[0], w, 80, True, GORL_TO_OP['>='], 0.5, 3, True, outs=outs)
        torch.cuda.synchronize()
        assert np.array_equal(np.concatenate([o.cpu().numpy() for o in res]), ref)
    finally:
        for e in engines:
            e.lib.ct_destroy(e.handle)
            e.handle = None


def test_sharded_empty_shards(fixture_cube):
    """A rank whose planes hold no component at all, components only on one side of a cut, and an empty cube."""
    a, lat, lon = fixture_cube
    w = row_weights(lat, lon)
    x = a.copy()
    x[:5] = -1000.0                                   # the first shard of (5, 6) is empty
    for parts in [(5, 6), (3, 2, 6), (4, 7)]:
        f, n, _ = run_local(x, w, parts, 150, '>=', 0.5, 2, True)
        ref = oracle.run_contrack(x, lat, lon, 150, '>=', 0.5, 2, True)
        assert np.array_equal(f, ref) and n == len(np.unique(ref)) - 1, parts
    x[:] = -1000.0
    f, n, _ = run_local(x, w, (4, 7), 150, '>=', 0.5, 2, True)
    assert n == 0 and not f.any()
    x[:] = 1000.0                                     # one component per plane covering everything
    f, n, _ = run_local(x, w, (6, 5), 150, '>=', 0.5, 2, True)
    assert np.array_equal(f, oracle.run_contrack(x, lat, lon, 150, '>=', 0.5, 2, True))


@pytest.mark.parametrize('parts', [(6, 5), (4, 3, 4), (1, 9, 1)])
def test_sharded_host_buffers(fixture_cube, reference_run, parts):
    """Host shards in, host flag planes out (sparse run-table export per rank), incl. a float64 cube and a stale-box cube."""
    a, lat, lon = fixture_cube
    w = row_weights(lat, lon)
    for r in reference_run['fixture'][:2]:
        f, n, st = run_local(a, w, parts, r['threshold'], r['gorl'], r['overlap'], r['persistence'], r['twosided'], host=True)
        assert sha_i4(f) == r['sha256'] and n == len(r['ids']), (parts, r['key'])
        assert st[0]['d2h_bytes'] < 0.1 * a[:parts[0]].size * 4
    ref = oracle.run_contrack(a.astype(np.float64), lat, lon, 150, '>=', 0.5, 5, True)
    f, n, _ = run_local(a.astype(np.float64), w, parts, 150, '>=', 0.5, 5, True, host=True)
    assert np.array_equal(f, ref)
    lat2, lon2 = regular_grid(24, 16)
    x = synth_cube(1396, 12, 24, 16, (1.5, 2, 2))
    f, _, _ = run_local(x, row_weights(lat2, lon2), (5, 4, 3), 60, '>=', 0.0, 1, False, host=True)
    assert np.array_equal(f, oracle.track_persistence((x >= 60).astype(int), 1))


def test_sharded_per_timestep_thresholds(fixture_cube):
    a, lat, lon = fixture_cube
    w = row_weights(lat, lon)
    thr = np.linspace(120, 180, a.shape[0])
    ref = oracle.run_contrack(a, lat, lon, thr, '>=', 0.5, 3, True)
    f, n, _ = run_local(a, w, (4, 3, 4), thr, '>=', 0.5, 3, True, f32=False)
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
