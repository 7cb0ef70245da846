"""Test helper: builds, with numpy/scipy on the CPU, the same tables the CUDA kernels hand to the ordered host phase
(``ct_host_tables``), so that phase can be checked against the oracle without a GPU.  Test infrastructure only."""
import ctypes as C

import numpy as np
from scipy import ndimage

from contrack_b200 import _lib

ONES33 = np.ones((3, 3), int)


def classify_rows(w, W):
    lib = _lib.load()
    w = np.ascontiguousarray(w, np.float64)
    sp = np.zeros(len(w), np.uint8)
    lib.ct_classify_rows(_lib.ptr(w, _lib._f64p), len(w), W, _lib.ptr(sp, _lib._u8p))
    return sp


def build_tables(mask, w):
    """mask [T,H,W] bool; w [H] float64 row weights.  Returns a dict of numpy arrays (see ct_host_tables)."""
    T, H, W = mask.shape
    special = classify_rows(w, W).astype(bool)
    wE = np.where(special, 0.0, w)
    wS = np.where(special, w, 0.0)
    labs, offs = [], [0]
    for t in range(T):
        lab, n = ndimage.label(mask[t], structure=ONES33)
        labs.append(lab)
        offs.append(offs[-1] + n)
    nc = offs[-1]
    comp_t = np.zeros(nc, np.int32)
    y0, y1, x0, x1 = (np.zeros(nc, np.int32) for _ in range(4))
    aE, aS = np.zeros(nc), np.zeros(nc)
    nsp = np.zeros(nc, np.uint32)
    cls = np.arange(nc, dtype=np.uint32)
    seam = []
    pairs = []
    runs = []
    plane_run_ptr = [0]
    for t in range(T):
        lab, o, n = labs[t], offs[t], offs[t + 1] - offs[t]
        comp_t[o:o + n] = t
        for k, sl in enumerate(ndimage.find_objects(lab)):
            y0[o + k], y1[o + k], x0[o + k], x1[o + k] = sl[0].start, sl[0].stop, sl[1].start, sl[1].stop
        flat = lab.ravel()
        aE[o:o + n] = np.bincount(flat, weights=np.repeat(wE, W), minlength=n + 1)[1:]
        aS[o:o + n] = np.bincount(flat, weights=np.repeat(wS, W), minlength=n + 1)[1:]
        nsp[o:o + n] = np.bincount(flat, weights=np.repeat(special.astype(float), W), minlength=n + 1)[1:].astype(np.uint32)
        # same-row date-line classes (contrack.py:691-698): min-id representative
        parent = list(range(n + 1))

        def find(i):
            while parent[i] != i:
                parent[i] = parent[parent[i]]
                i = parent[i]
            return i
        for y in range(H):
            a, b = lab[y, 0], lab[y, -1]
            if a > 0 and b > 0:
                seam.append((t * H + y, o + a - 1, o + b - 1))
                ra, rb = find(a), find(b)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
        for k in range(1, n + 1):
            cls[o + k - 1] = o + find(k) - 1
        if t > 0:
            prev = labs[t - 1]
            both = (lab > 0) & (prev > 0)
            if both.any():
                yy = np.nonzero(both)[0]
                key = (lab[both].astype(np.int64) << 32) | prev[both]
                uk, inv = np.unique(key, return_inverse=True)
                cnt = np.bincount(inv)
                pe = np.bincount(inv, weights=wE[yy])
                ps = np.bincount(inv, weights=wS[yy])
                pn = np.bincount(inv, weights=special[yy].astype(float)).astype(np.uint32)
                for i, k in enumerate(uk):
                    pairs.append((o + (k >> 32) - 1, offs[t - 1] + (k & 0xffffffff) - 1, cnt[i], pn[i], pe[i], ps[i]))
        # row-runs in raster order
        m = mask[t].astype(np.int8)
        d = np.diff(np.concatenate([np.zeros((H, 1), np.int8), m, np.zeros((H, 1), np.int8)], axis=1), axis=1)
        ys, xs = np.nonzero(d == 1)
        ye, xe = np.nonzero(d == -1)
        for y, a, e in zip(ys, xs, xe):
            runs.append((y, a, e, o + lab[y, a] - 1))
        plane_run_ptr.append(len(runs))
    rng = np.random.default_rng(0)
    pairs = [pairs[i] for i in rng.permutation(len(pairs))]        # the GPU hands pairs over in arbitrary order
    tb = dict(T=T, H=H, W=W, w=np.ascontiguousarray(w, np.float64), labs=labs, offs=offs, ncomp=nc,
              comp_t=comp_t, y0=y0, y1=y1, x0=x0, x1=x1, cls=cls, aE=aE, aS=aS, nsp=nsp,
              pair_a=np.array([p[0] for p in pairs], np.uint32), pair_b=np.array([p[1] for p in pairs], np.uint32),
              pair_npix=np.array([p[2] for p in pairs], np.uint32), pair_nsp=np.array([p[3] for p in pairs], np.uint32),
              pair_E=np.array([p[4] for p in pairs], np.float64), pair_S=np.array([p[5] for p in pairs], np.float64),
              seam_row=np.array([s[0] for s in seam], np.uint32), seam_a=np.array([s[1] for s in seam], np.uint32),
              seam_b=np.array([s[2] for s in seam], np.uint32),
              plane_run_ptr=np.array(plane_run_ptr, np.int64),
              run_y=np.array([r[0] for r in runs], np.int32), run_x0=np.array([r[1] for r in runs], np.int32),
              run_x1=np.array([r[2] for r in runs], np.int32), run_comp=np.array([r[3] for r in runs], np.uint32))
    return tb


def host_tables(tb, overlap, persistence, twosided, stage=0, with_runs=True):
    """Calls ct_host_tables and paints the result: returns (flag [T,H,W] int32, stats8)."""
    lib = _lib.load()
    L = _lib
    nc = tb['ncomp']
    val = np.zeros(max(nc, 1), np.int32)
    cap = 1 << 16
    ovr = [np.zeros(cap, np.int32) for _ in range(5)]
    n_ovr = C.c_long(0)
    stats = (C.c_long * 8)()
    rp = tb['plane_run_ptr'] if with_runs else None
    rc = lib.ct_host_tables(
        tb['T'], tb['H'], tb['W'], L.ptr(tb['w'], L._f64p), float(overlap), int(persistence), int(bool(twosided)), stage,
        nc, L.ptr(tb['comp_t'], L._i32p), L.ptr(tb['y0'], L._i32p), L.ptr(tb['y1'], L._i32p), L.ptr(tb['x0'], L._i32p),
        L.ptr(tb['x1'], L._i32p), L.ptr(tb['cls'], L._u32p), L.ptr(tb['aE'], L._f64p), L.ptr(tb['aS'], L._f64p),
        L.ptr(tb['nsp'], L._u32p),
        len(tb['pair_a']), L.ptr(tb['pair_a'], L._u32p), L.ptr(tb['pair_b'], L._u32p), L.ptr(tb['pair_npix'], L._u32p),
        L.ptr(tb['pair_nsp'], L._u32p), L.ptr(tb['pair_E'], L._f64p), L.ptr(tb['pair_S'], L._f64p),
        len(tb['seam_row']), L.ptr(tb['seam_row'], L._u32p), L.ptr(tb['seam_a'], L._u32p), L.ptr(tb['seam_b'], L._u32p),
        L.ptr(rp, L._i64p), L.ptr(tb['run_y'] if with_runs else None, L._i32p),
        L.ptr(tb['run_x0'] if with_runs else None, L._i32p), L.ptr(tb['run_x1'] if with_runs else None, L._i32p),
        L.ptr(tb['run_comp'] if with_runs else None, L._u32p),
        L.ptr(val, L._i32p), cap, *[L.ptr(o, L._i32p) for o in ovr], C.byref(n_ovr), stats)
    L.check(rc)
    T, H, W = tb['T'], tb['H'], tb['W']
    flag = np.zeros((T, H, W), np.int32)
    lut_all = np.concatenate([[0], val[:nc]]).astype(np.int32)
    for t in range(T):
        lab, o = tb['labs'][t], tb['offs'][t]
        lut = np.concatenate([[0], lut_all[1 + o:1 + tb['offs'][t + 1]]])
        flag[t] = lut[lab]
    for i in range(n_ovr.value):
        flag[ovr[0][i], ovr[1][i], ovr[2][i]:ovr[3][i]] = ovr[4][i]
    return flag, list(stats)


def legacy_to_view(tb, has_prev, t_begin):
    """The dict build_tables() returns -> a rank-local view in the layout of ct_shard_view (contrack_b200/sharded.py):
    class sums at the representative, pairs in CSR form over the plane-t component, date-line rows as segments."""
    nc = tb['ncomp']
    cls = tb['cls'].astype(np.int64)
    conE = np.bincount(cls, weights=tb['aE'], minlength=nc) if nc else np.zeros(0)
    conS = np.bincount(cls, weights=tb['aS'], minlength=nc) if nc else np.zeros(0)
    nsp = np.bincount(cls, weights=tb['nsp'].astype(float), minlength=nc).astype(np.uint32) if nc else np.zeros(0, np.uint32)
    pa, pb = tb['pair_a'].astype(np.int64), tb['pair_b'].astype(np.int64)
    rep_b = cls[pb] if len(pb) else np.zeros(0, np.int64)
    fE = np.bincount(rep_b, weights=tb['pair_E'], minlength=nc) if nc else np.zeros(0)
    fS = np.bincount(rep_b, weights=tb['pair_S'], minlength=nc) if nc else np.zeros(0)
    fnsp = np.bincount(rep_b, weights=tb['pair_nsp'].astype(float), minlength=nc).astype(np.uint32) if nc else np.zeros(0, np.uint32)
    order = np.argsort(pa, kind='stable')
    ptr = np.zeros(nc + 1, np.uint32)
    if nc:
        ptr[1:] = np.cumsum(np.bincount(pa, minlength=nc))
    segs = []
    H = tb['H']
    for row, a, b in zip(tb['seam_row'], tb['seam_a'], tb['seam_b']):
        t, y = int(row) // H, int(row) % H
        if segs and segs[-1][0] == t and segs[-1][2] == y and segs[-1][3] == a and segs[-1][4] == b:
            segs[-1][2] = y + 1
        else:
            segs.append([t, y, y + 1, a, b])
    nh = int((tb['comp_t'] == 0).sum()) if has_prev else 0
    d = dict(planes=tb['T'], ncomp=nc, halo_comps=nh, npair=len(pa), nseg=len(segs), has_prev=int(has_prev),
             t_begin=int(t_begin), comp_t=tb['comp_t'], comp_y0=tb['y0'], comp_y1=tb['y1'], comp_x0=tb['x0'],
             comp_x1=tb['x1'], comp_cls=tb['cls'], cls_conE=conE, cls_conS=conS, cls_fE=fE, cls_fS=fS, cls_nsp=nsp,
             cls_fnsp=fnsp, pair_ptr=ptr, pair_b=tb['pair_b'][order], pair_npix=tb['pair_npix'][order],
             pair_nsp=tb['pair_nsp'][order], pair_E=tb['pair_E'][order], pair_S=tb['pair_S'][order],
             seg_t=np.array([s[0] for s in segs], np.int32), seg_y0=np.array([s[1] for s in segs], np.int32),
             seg_y1=np.array([s[2] for s in segs], np.int32), seg_a=np.array([s[3] for s in segs], np.uint32),
             seg_b=np.array([s[4] for s in segs], np.uint32))
    return d
