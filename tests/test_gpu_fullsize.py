"""BASELINE.json configs[1] at its full size (2707 x 721 x 1440, threshold 160 >=, overlap 0.5, persistence 5, twosided) on
the synthetic Z500-like field of bench.py: size-independent properties of the result, equality of the single-GPU path, its
kernel variants and the time-sharded path, and the first planes against the oracle.  Needs ~35 GB of HBM."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

T, THR, OV, PERS = 2707, 160, 0.5, 5


@pytest.fixture(scope='module')
def cube():
    import torch
    import bench
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip('needs 40 GB of free device memory')
    a = torch.empty((T, bench.H, bench.W), dtype=torch.float32, device='cuda')
    bench.synth_fill(a, 0, T)
    torch.cuda.synchronize()
    lat, lon = bench.grid()
    return a, bench.reference_weights(lat, lon), lat, lon


@pytest.fixture(scope='module')
def result(cube):
    import torch
    from contrack_b200 import Engine
    a, w, _, _ = cube
    eng = Engine.get(0)
    flag, n = eng.run_contrack(a, w, THR, True, 0, OV, PERS, True)
    torch.cuda.synchronize()
    return flag, n, dict(eng.stats())


def test_properties(cube, result):
    import torch
    a, w, _, _ = cube
    flag, n, stats = result
    assert flag.dtype == torch.int32 and tuple(flag.shape) == tuple(a.shape)
    nz = flag != 0
    assert bool((a[nz] >= THR).all())                                   # flagged cells are cells of the mask
    assert int(flag.min()) == 0
    ids = torch.unique(flag[nz])
    assert len(ids) == n and n > 1000                                   # contrack.py:793
    # persistence (contrack.py:765-772): every surviving id spans at least PERS time steps
    tt = torch.nonzero(nz)[:, 0]
    v = flag[nz].long()
    tmin = torch.full((int(ids.max()) + 1,), T, dtype=torch.long, device='cuda').scatter_reduce(0, v, tt, 'amin')
    tmax = torch.zeros(int(ids.max()) + 1, dtype=torch.long, device='cuda').scatter_reduce(0, v, tt, 'amax')
    assert bool(((tmax[ids.long()] - tmin[ids.long()] + 1) >= PERS).all())
    # first and last plane are never filtered by the overlap test (contrack.py:706: tt in 1 .. T-2) but persistence applies
    assert stats['features'] == n and stats['runs'] > 1e6


def test_variants_and_sharded_agree(cube, result):
    import torch
    from contrack_b200 import Engine, sharded
    a, w, _, _ = cube
    flag, n, _ = result
    eng = Engine.get(0)
    out = torch.empty_like(flag)
    for opts in ({'chunks': 1}, {'plane_kernel': 1}, {'plane_kernel': 1, 'fast_chunks': 3}, {'overlap_zero': 0},
                 {'max_sweeps': 2}, {'tma': 0}):
        defaults = {'chunks': 4, 'plane_kernel': 2, 'fast_chunks': 1, 'overlap_zero': 1, 'max_sweeps': 32, 'tma': 1}
        for k, v in opts.items():
            eng.set_option(k, v)
        try:
            f2, n2 = eng.run_contrack(a, w, THR, True, 0, OV, PERS, True, out=out)
        finally:
            for k in opts:
                eng.set_option(k, defaults[k])
        assert n2 == n and bool(torch.equal(f2, flag)), opts
    del out
    engines = [Engine(0) for _ in range(3)]
    try:
        bounds = sharded.shard_bounds(T, 3)
        outs, n3, _ = sharded.run_local_group(engines, [a[b0:b1] for b0, b1 in bounds], T, w, THR, True, 0, OV, PERS, True)
        assert n3 == n
        for (b0, b1), o in zip(bounds, outs):
            assert bool(torch.equal(o, flag[b0:b1]))
    finally:
        for e in engines:
            if e.handle:
                e.lib.ct_destroy(e.handle); e.handle = None


def test_host_buffer_entry_point_agrees(cube, result):
    import torch
    from contrack_b200 import Engine
    a, w, _, _ = cube
    flag, n, _ = result
    Ts = 300
    sub = a[:Ts].contiguous()
    ref, nref = Engine.get(0).run_contrack(sub, w, THR, True, 0, OV, PERS, True)
    got, ngot = Engine.get(0).run_contrack(sub.cpu().numpy(), w, THR, True, 0, OV, PERS, True)
    assert ngot == nref and np.array_equal(got, ref.cpu().numpy())


def test_first_planes_against_the_oracle(cube):
    from oracle import contrack_oracle as oracle
    from contrack_b200 import Engine
    a, w, lat, lon = cube
    Ts = 48
    sub = a[:Ts].contiguous()
    x = sub.cpu().numpy()
    ref = oracle.run_contrack(x, lat, lon, THR, '>=', OV, PERS, True, force=True)
    got, n = Engine.get(0).run_contrack(sub, w, THR, True, 0, OV, PERS, True)
    assert np.array_equal(got.cpu().numpy(), ref) and n == len(np.unique(ref)) - 1
