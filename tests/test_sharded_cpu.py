"""Host-side logic of the time-sharded (multi-GPU) path on CPU: world_size 2 and 3, gloo backend."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_tables_gloo(world):
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='1')
    port = 29500 + world + (os.getpid() % 200)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', '_shard_worker.py')]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(' ok (') == world


def test_pack_roundtrip_and_bounds():
    from contrack_b200 import sharded
    assert sharded.shard_bounds(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert sharded.shard_bounds(10957, 8)[-1][1] == 10957
    rng = np.random.default_rng(0)
    d = dict(planes=5, ncomp=7, halo_comps=2, npair=3, nseg=1, has_prev=1, t_begin=40)
    for name, dt, lk in sharded.ARRAYS:
        n = d['ncomp'] + 1 if lk == 'ncomp+1' else d[lk]
        d[name] = (rng.integers(0, 100, n)).astype(dt)
    e = sharded.unpack_view(sharded.pack_view(d))
    for k in d:
        assert np.array_equal(d[k], e[k]), k


def test_device_path_offsets_match_the_host_merge():
    """comp_offsets (what the device-table path hands to ct_shard_paint_global) = the offsets merge_views computes."""
    from contrack_b200 import sharded
    from _common import row_weights
    from _synth import synth_cube
    from _tables_np import build_tables, legacy_to_view
    T, H, W = 14, 20, 24
    x = synth_cube(11, T, H, W, (1.5, 2, 3))
    lat = (60 - np.arange(H) * 2.0).astype(np.float32)
    lon = (np.arange(W) * 2.0).astype(np.float32)
    w = row_weights(lat, lon)
    for world in (2, 3, 5):
        views = []
        for r, (t0, t1) in enumerate(sharded.shard_bounds(T, world)):
            hp = 1 if r else 0
            views.append(legacy_to_view(build_tables(x[t0 - hp:t1] >= 40, w), hp, t0))
        g, offs = sharded.merge_views(views)
        counts = [[v['t_begin'] - v['has_prev'], v['ncomp'], v['halo_comps'], v['npair'], v['nseg'], 0, 0, 0] for v in views]
        assert sharded.comp_offsets(counts) == offs
        assert g['ncomp'] == sum(v['ncomp'] - v['halo_comps'] for v in views)
