"""Shared helpers of the parity tests (test infrastructure; may use the oracle)."""
import hashlib

import numpy as np

from oracle import contrack_oracle as oracle


def sha_i4(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype('<i4').tobytes()).hexdigest()


def row_weights(lat, lon):
    """[H] float64 row weights exactly as the reference computes them (contrack.py:703-704), via the oracle."""
    return oracle.weight_grid(lat, oracle.resolution(lat), oracle.resolution(lon), len(lon))[:, 0].copy()


def same_partition(a, b):
    """True if two label arrays have identical zero sets and identical equivalence classes (ids may be permuted)."""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    if not np.array_equal(a == 0, b == 0):
        return False
    nz = a != 0
    pa, pb = a[nz].astype(np.int64), b[nz].astype(np.int64)
    if pa.size == 0:
        return True
    pairs = np.unique(np.stack([pa, pb], 1), axis=0)
    return len(np.unique(pairs[:, 0])) == len(pairs) and len(np.unique(pairs[:, 1])) == len(pairs)
