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


def class_fractions(a, lat, lon, thr, t):
    """(label, forward fraction, backward fraction, first row, last row) of every date-line class of plane t, computed
    with the reference's own expressions (contrack.py:717-722) -- for t = 1, where plane t-1 is never filtered."""
    from scipy import ndimage
    st = {}
    oracle.run_contrack(a, lat, lon, thr, '>=', 0.0, 1, False, stages=st)
    flag = st['label2d_seam']
    wgrid = oracle.weight_grid(lat, oracle.resolution(lat), oracle.resolution(lon), len(lon))
    out = []
    for label, sl in enumerate(ndimage.find_objects(flag[t]), 1):
        if sl is None:
            continue
        m = flag[t][sl] == label
        ac = np.sum(wgrid[sl][m])
        f = np.sum(wgrid[sl][m & (flag[t + 1][sl] >= 1)])
        b = np.sum(wgrid[sl][m & (flag[t - 1][sl] >= 1)])
        rows = np.nonzero(m.any(1))[0] + sl[0].start
        out.append((label, float((1 / ac) * f), float((1 / ac) * b), int(rows.min()), int(rows.max())))
    return out


def pole_tie_overlaps(a, lat, lon, thr):
    """`overlap` values that sit exactly on (and one ulp around) a fraction of a class that mixes a pole row with
    ordinary rows: the verdict then depends on the last bit of numpy's pairwise sums."""
    vals = []
    for label, ff, fb, y0, y1 in class_fractions(a, lat, lon, thr, 1):
        if (y0 == 0 or y1 == len(lat) - 1) and y1 > y0:
            for f in (ff, fb):
                if 0 < f < 1:
                    vals += [f, float(np.nextafter(f, 1)), float(np.nextafter(f, 0))]
    return vals
