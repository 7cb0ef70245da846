"""Numpy statement of the table merge of a time-sharded run (test infrastructure).

The product merges the rank tables on the device (contrack_b200/csrc/ct_dist.cu: k_merge_*).  These helpers restate that
renumbering with numpy -- local component i of rank r becomes i + (own components of ranks < r) - halo_r; the halo
components of rank r fall onto the ids rank r-1 gave its last-plane components; forward sums accumulated on halo copies are
added to their owners -- and drive the library's all-host ordered phase (ct_host_tables_fast) with it, so that the merge
logic can be checked on CPU ranks (gloo, tests/test_sharded_cpu.py) against the oracle.
"""
import ctypes as C

import numpy as np

from contrack_b200 import _lib

# arrays of a view: name -> (dtype, length key)
ARRAYS = [('comp_t', np.int32, 'ncomp'), ('comp_y0', np.int32, 'ncomp'), ('comp_y1', np.int32, 'ncomp'),
          ('comp_x0', np.int32, 'ncomp'), ('comp_x1', np.int32, 'ncomp'), ('comp_cls', np.uint32, 'ncomp'),
          ('cls_conE', np.float64, 'ncomp'), ('cls_conS', np.float64, 'ncomp'), ('cls_fE', np.float64, 'ncomp'),
          ('cls_fS', np.float64, 'ncomp'), ('cls_nsp', np.uint32, 'ncomp'), ('cls_fnsp', np.uint32, 'ncomp'),
          ('pair_ptr', np.uint32, 'ncomp+1'), ('pair_b', np.uint32, 'npair'), ('pair_npix', np.uint32, 'npair'),
          ('pair_nsp', np.uint32, 'npair'), ('pair_E', np.float64, 'npair'), ('pair_S', np.float64, 'npair'),
          ('seg_t', np.int32, 'nseg'), ('seg_y0', np.int32, 'nseg'), ('seg_y1', np.int32, 'nseg'),
          ('seg_a', np.uint32, 'nseg'), ('seg_b', np.uint32, 'nseg')]
SCALARS = ['planes', 'ncomp', 'halo_comps', 'npair', 'nseg', 'has_prev', 't_begin']


def pack_view(d):
    """One flat uint8 buffer (scalars as int64 header, then the arrays, each padded to 8 bytes)."""
    parts = [np.array([d[k] for k in SCALARS], np.int64).view(np.uint8)]
    for name, dt, _ in ARRAYS:
        b = np.ascontiguousarray(d[name], dt).view(np.uint8)
        pad = (-len(b)) % 8
        parts.append(b)
        if pad:
            parts.append(np.zeros(pad, np.uint8))
    return np.concatenate(parts)


def unpack_view(buf):
    hdr = buf[:8 * len(SCALARS)].view(np.int64)
    d = {k: int(hdr[i]) for i, k in enumerate(SCALARS)}
    off = 8 * len(SCALARS)
    for name, dt, lk in ARRAYS:
        n = d['ncomp'] + 1 if lk == 'ncomp+1' else d[lk]
        nb = n * np.dtype(dt).itemsize
        d[name] = buf[off:off + nb].view(dt).copy()
        off += nb + ((-nb) % 8)
    return d


def allgather_bytes(buf, group=None, device=None):
    """all_gather of one variable-length uint8 numpy buffer per rank -> list of numpy buffers (rank order)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = device if device is not None else torch.device('cpu')
    n = torch.tensor([len(buf)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    mine = torch.zeros(m, dtype=torch.uint8, device=dev)
    mine[:len(buf)] = torch.from_numpy(buf).to(dev)
    out = [torch.empty(m, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [o[:s].cpu().numpy() for o, s in zip(out, sizes)]


def merge_views(views):
    """Rank-local tables -> global tables in the layout of ct_host_tables_fast.

    Local component i of rank r has the global id  i + off_r,  off_r = (own components of ranks < r) - halo_comps_r:
    own components are numbered consecutively in rank order (= global first-pixel order, because ranks are ordered in
    time), and the halo components of rank r -- the components of rank r-1's last plane, same bit rows, same raster order
    -- fall exactly onto the ids rank r-1 gave them.  Returns (tables dict, offsets list).
    """
    offs, base = [], 0
    for v in views:
        offs.append(base - v['halo_comps'])
        base += v['ncomp'] - v['halo_comps']
    nc = base
    g = {k: [] for k in ('comp_t', 'comp_y0', 'comp_y1', 'comp_x0', 'comp_x1', 'comp_cls', 'pair_cnt', 'pair_b',
                         'pair_npix', 'pair_nsp', 'pair_E', 'pair_S', 'seg_t', 'seg_y0', 'seg_y1', 'seg_a', 'seg_b')}
    conE, conS, fE, fS = (np.zeros(nc) for _ in range(4))
    nsp = np.zeros(nc, np.int64)
    base = 0
    for r, (v, off) in enumerate(zip(views, offs)):
        nh, n = v['halo_comps'], v['ncomp']
        own = slice(nh, n)
        n_own = n - nh
        if nh:
            prev = views[r - 1]
            n_last = int((prev['comp_t'] == prev['planes'] - 1).sum())
            if n_last != nh:
                raise RuntimeError('rank %d sees %d components in its halo plane, rank %d has %d in its last plane'
                                   % (r, nh, r - 1, n_last))
            # forward overlap of rank r-1's last-plane classes with rank r's first plane was accumulated on rank r
            sl = slice(base - nh, base)
            fE[sl] += v['cls_fE'][:nh]
            fS[sl] += v['cls_fS'][:nh]
            nsp[sl] += v['cls_fnsp'][:nh]
        sl = slice(base, base + n_own)
        conE[sl] = v['cls_conE'][own]; conS[sl] = v['cls_conS'][own]
        fE[sl] += v['cls_fE'][own]; fS[sl] += v['cls_fS'][own]
        nsp[sl] += v['cls_nsp'][own].astype(np.int64) + v['cls_fnsp'][own]
        g['comp_t'].append(v['comp_t'][own] - v['has_prev'] + v['t_begin'])
        for k in ('comp_y0', 'comp_y1', 'comp_x0', 'comp_x1'):
            g[k].append(v[k][own])
        g['comp_cls'].append((v['comp_cls'][own].astype(np.int64) + off).astype(np.uint32))
        pp = v['pair_ptr'].astype(np.int64)
        g['pair_cnt'].append(np.diff(pp)[own])
        e0, e1 = int(pp[nh]), int(pp[n])
        g['pair_b'].append((v['pair_b'][e0:e1].astype(np.int64) + off).astype(np.uint32))
        for k in ('pair_npix', 'pair_nsp', 'pair_E', 'pair_S'):
            g[k].append(v[k][e0:e1])
        keep = v['seg_t'] >= v['has_prev']
        g['seg_t'].append(v['seg_t'][keep] - v['has_prev'] + v['t_begin'])
        g['seg_y0'].append(v['seg_y0'][keep]); g['seg_y1'].append(v['seg_y1'][keep])
        g['seg_a'].append((v['seg_a'][keep].astype(np.int64) + off).astype(np.uint32))
        g['seg_b'].append((v['seg_b'][keep].astype(np.int64) + off).astype(np.uint32))
        base += n_own
    cat = lambda k, dt: np.ascontiguousarray(np.concatenate(g[k]) if g[k] else np.zeros(0), dt)   # noqa: E731
    out = dict(ncomp=nc, comp_t=cat('comp_t', np.int32), comp_y0=cat('comp_y0', np.int32), comp_y1=cat('comp_y1', np.int32),
               comp_x0=cat('comp_x0', np.int32), comp_x1=cat('comp_x1', np.int32), comp_cls=cat('comp_cls', np.uint32),
               cls_conE=conE, cls_conS=conS, cls_fE=fE, cls_fS=fS, cls_nsp=nsp.astype(np.uint32),
               pair_b=cat('pair_b', np.uint32), pair_npix=cat('pair_npix', np.uint32), pair_nsp=cat('pair_nsp', np.uint32),
               pair_E=cat('pair_E', np.float64), pair_S=cat('pair_S', np.float64),
               seg_t=cat('seg_t', np.int32), seg_y0=cat('seg_y0', np.int32), seg_y1=cat('seg_y1', np.int32),
               seg_a=cat('seg_a', np.uint32), seg_b=cat('seg_b', np.uint32))
    ptr = np.zeros(nc + 1, np.uint32)
    if nc:
        ptr[1:] = np.cumsum(cat('pair_cnt', np.int64))
    out['pair_ptr'] = ptr
    out['npair'] = len(out['pair_b'])
    out['nseg'] = len(out['seg_t'])
    return out, offs


_FETCH = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_long, C.POINTER(C.c_long), C.POINTER(_lib._i32p), C.POINTER(_lib._i32p),
                     C.POINTER(_lib._i32p), C.POINTER(_lib._u32p))


def host_tables_fast(T, H, W, w, g, overlap, persistence, twosided, stage=0, fetch=None):
    """ct_host_tables_fast on merged tables.  fetch(t) -> (y, x0, x1, comp_global) int32/uint32 arrays, or None.
    Returns (comp_val [ncomp] int32, overrides [(t, y, x0, x1, val)], stats8)."""
    lib = _lib.load()
    L = _lib
    nc = g['ncomp']
    val = np.zeros(max(nc, 1), np.int32)
    cap = 1 << 16
    ovr = [np.zeros(cap, np.int32) for _ in range(5)]
    n_ovr = C.c_long(0)
    stats = (C.c_long * 8)()
    keep = {}

    def _cb(user, t, n, y, x0, x1, comp):
        try:
            arrs = fetch(int(t))
            arrs = (np.ascontiguousarray(arrs[0], np.int32), np.ascontiguousarray(arrs[1], np.int32),
                    np.ascontiguousarray(arrs[2], np.int32), np.ascontiguousarray(arrs[3], np.uint32))
            keep['a'] = arrs
            n[0] = len(arrs[0])
            y[0] = L.ptr(arrs[0], L._i32p); x0[0] = L.ptr(arrs[1], L._i32p); x1[0] = L.ptr(arrs[2], L._i32p)
            comp[0] = L.ptr(arrs[3], L._u32p)
            return 0
        except Exception:                      # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return -1

    cb = _FETCH(_cb) if fetch is not None else C.cast(None, _FETCH)
    w = np.ascontiguousarray(w, np.float64)
    rc = lib.ct_host_tables_fast(
        int(T), int(H), int(W), L.ptr(w, L._f64p), float(overlap), int(persistence), int(bool(twosided)), int(stage), nc,
        L.ptr(g['comp_t'], L._i32p), L.ptr(g['comp_y0'], L._i32p), L.ptr(g['comp_y1'], L._i32p), L.ptr(g['comp_x0'], L._i32p),
        L.ptr(g['comp_x1'], L._i32p), L.ptr(g['comp_cls'], L._u32p), L.ptr(g['cls_conE'], L._f64p),
        L.ptr(g['cls_conS'], L._f64p), L.ptr(g['cls_fE'], L._f64p), L.ptr(g['cls_fS'], L._f64p), L.ptr(g['cls_nsp'], L._u32p),
        L.ptr(g['pair_ptr'], L._u32p), L.ptr(g['pair_b'], L._u32p), L.ptr(g['pair_npix'], L._u32p),
        L.ptr(g['pair_nsp'], L._u32p), L.ptr(g['pair_E'], L._f64p), L.ptr(g['pair_S'], L._f64p), g['nseg'],
        L.ptr(g['seg_t'], L._i32p), L.ptr(g['seg_y0'], L._i32p), L.ptr(g['seg_y1'], L._i32p), L.ptr(g['seg_a'], L._u32p),
        L.ptr(g['seg_b'], L._u32p), cb, None, L.ptr(val, L._i32p), cap, *[L.ptr(o, L._i32p) for o in ovr],
        C.byref(n_ovr), stats)
    L.check(rc)
    k = n_ovr.value
    overrides = [tuple(int(o[i]) for o in ovr) for i in range(k)]
    return val[:nc], overrides, list(stats)


