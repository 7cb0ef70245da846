"""The C-ABI library loads without a GPU and exports every symbol include/contrack_b200.h declares -- CPU only."""
import ctypes as C
import os
import re

import numpy as np

from contrack_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, 'include', 'contrack_b200.h')).read()
    declared = set(re.findall(r'\b(ct_[a-z0-9_]+)\s*\(', hdr))
    declared -= {'ct_ctx', 'ct_status', 'ct_dtype', 'ct_op', 'ct_stage'}
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.exported_symbols())


def test_version_and_errors_without_gpu():
    lib = _lib.load()
    assert lib.ct_version() >= 100
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        rc = lib.ct_create(0, C.byref(h))
        assert rc == _lib.CT_ERR_CUDA and b'CUDA' in lib.ct_last_error()


def test_numpy_pairwise_sum_matches_numpy():
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n in [0, 1, 7, 8, 9, 127, 128, 129, 1000, 4097, 100003]:
        v = (rng.standard_normal(n) * 10 ** rng.uniform(-8, 8, n)).astype(np.float64)
        cnt = np.ones(n, np.int64)
        got = lib.ct_numpy_pairwise_sum_rle(_lib.ptr(v, _lib._f64p), _lib.ptr(cnt, _lib._i64p), n)
        assert got == np.sum(v), n
    v = np.array([3.36, -5.39e-4, 770.5], np.float64)
    cnt = np.array([1000, 360, 777], np.int64)
    got = lib.ct_numpy_pairwise_sum_rle(_lib.ptr(v, _lib._f64p), _lib.ptr(cnt, _lib._i64p), 3)
    assert got == np.sum(np.repeat(v, cnt))


def test_classify_rows_poles():
    from _common import row_weights
    for H, W in [(181, 360), (721, 1440), (91, 180)]:
        lat = np.linspace(90, -90, H).astype(np.float32)
        lon = (np.arange(W) * (360.0 / W)).astype(np.float32)
        from oracle import contrack_oracle as oracle
        w = oracle.weight_grid(lat, oracle.resolution(lat, True), oracle.resolution(lon, True), W)[:, 0].copy()
        sp = np.zeros(H, np.uint8)
        _lib.load().ct_classify_rows(_lib.ptr(w, _lib._f64p), H, W, _lib.ptr(sp, _lib._u8p))
        assert sp[0] == 1 and sp[-1] == 1 and sp.sum() == 2, (H, W, sp.sum())
        # the exact set really is exact: any order of summation gives the same float64
        rng = np.random.default_rng(1)
        vals = np.repeat(w[1:-1], W)
        s1 = np.sum(vals)
        s2 = np.sum(vals[rng.permutation(len(vals))])
        s3 = float(np.cumsum(vals[::-1])[-1])
        assert s1 == s2 == s3
