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
    import _shard_np
    assert sharded.shard_bounds(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert sharded.shard_bounds(10957, 8)[-1][1] == 10957
    rng = np.random.default_rng(0)
    d = dict(planes=5, ncomp=7, halo_comps=2, npair=3, nseg=1, has_prev=1, t_begin=40)
    for name, dt, lk in _shard_np.ARRAYS:
        n = d['ncomp'] + 1 if lk == 'ncomp+1' else d[lk]
        d[name] = (rng.integers(0, 100, n)).astype(dt)
    e = _shard_np.unpack_view(_shard_np.pack_view(d))
    for k in d:
        assert np.array_equal(d[k], e[k]), k
