"""Time-sharded run_contrack over several GPUs (one process per GPU, torch.distributed for the plumbing).

SURVEY.md 8(e): every rank owns a contiguous range of time planes.  Thresholding, 2-D labelling, area tables and painting
are local; what crosses ranks is
  * ONE boundary plane per neighbour pair: the bit rows of a rank's last plane go to the next rank (130 KB at 721x1440),
    so that rank can build the (plane t, plane t-1) pair tables across the cut, and
  * the component / class / pair / date-line tables (a few MB in total), all-gathered DEVICE to DEVICE (NCCL) and merged
    by a kernel into global tables on every rank, which then replays the global part of the path (keep/kill recurrence,
    3-D numbering, stale-box date-line merge, persistence) on its copy and obtains the same global ids -- this is the
    "global relabel" step; no rank-dependent numbering exists.
The cube itself never moves.  `run_contrack_sharded` is that path; `run_contrack_sharded_host` is the older variant that
gathers the tables through the host and replays the ordered phase with ct_host_tables_fast (kept because its merge,
`merge_views`, is the CPU-testable statement of the renumbering the merge kernel performs).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_VIEW_FIELDS = [('planes', C.c_long), ('ncomp', C.c_long), ('halo_comps', C.c_long), ('npair', C.c_long),
                ('nseg', C.c_long), ('nruns', C.c_long)] + \
               [(k, _lib._i32p) for k in ('comp_t', 'comp_y0', 'comp_y1', 'comp_x0', 'comp_x1')] + \
               [('comp_cls', _lib._u32p)] + [(k, _lib._f64p) for k in ('cls_conE', 'cls_conS', 'cls_fE', 'cls_fS')] + \
               [(k, _lib._u32p) for k in ('cls_nsp', 'cls_fnsp', 'pair_ptr', 'pair_b', 'pair_npix', 'pair_nsp')] + \
               [(k, _lib._f64p) for k in ('pair_E', 'pair_S')] + [(k, _lib._i32p) for k in ('seg_t', 'seg_y0', 'seg_y1')] + \
               [(k, _lib._u32p) for k in ('seg_a', 'seg_b')]


class ShardView(C.Structure):
    _fields_ = _VIEW_FIELDS


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


def view_to_dict(v, has_prev, t_begin):
    """Copies of a ct_shard_view's arrays (the view memory is reused by the next library call)."""
    d = dict(planes=int(v.planes), ncomp=int(v.ncomp), halo_comps=int(v.halo_comps), npair=int(v.npair),
             nseg=int(v.nseg), has_prev=int(has_prev), t_begin=int(t_begin))
    for name, dt, lk in ARRAYS:
        n = d['ncomp'] + 1 if lk == 'ncomp+1' else d[lk]
        p = getattr(v, name)
        d[name] = np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n > 0 and p else np.zeros(n, dt)
    if d['ncomp'] == 0:
        d['pair_ptr'] = np.zeros(1, np.uint32)
    return d


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


def shard_bounds(T, world):
    """Contiguous, near-equal time ranges; every rank gets at least one plane (world <= T)."""
    edges = [(T * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


class Shard(object):
    """One rank's part of a time-sharded run: thin wrapper over the ct_shard_* entry points."""

    def __init__(self, engine, anom_local, t_begin, has_prev, out=None):
        import torch
        self.engine, self.lib, self.h = engine, engine.lib, engine.handle
        self.anom = anom_local if anom_local.is_contiguous() else anom_local.contiguous()
        self.dev = self.anom.device
        self.Tl, self.H, self.W = (int(s) for s in self.anom.shape)
        self.t_begin, self.has_prev = int(t_begin), int(bool(has_prev))
        self.dtype = {torch.float32: _lib.CT_F32, torch.float64: _lib.CT_F64}[self.anom.dtype]
        self.out = out if out is not None else torch.empty((self.Tl, self.H, self.W), dtype=torch.int32, device=self.dev)
        self.view = None

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def threshold(self, w, thresholds, thr_is_f32, op):
        w = np.ascontiguousarray(w, np.float64)
        thr = np.ascontiguousarray(np.atleast_1d(thresholds), np.float64)
        _lib.check(self.lib.ct_shard_threshold(self.h, C.c_void_p(self.anom.data_ptr()), self.dtype, self.Tl, self.H,
                                               self.W, _lib.ptr(w, _lib._f64p), _lib.ptr(thr, _lib._f64p), len(thr),
                                               int(thr_is_f32), int(op), self.has_prev, self._stream()))

    def begin(self, w, thresholds, thr_is_f32, op):
        """ct_shard_begin: thresholds only the last own plane and returns its bit rows (int32 CUDA tensor) for the next
        rank; the other planes are thresholded inside tables_dev(), pipelined with the table kernels."""
        import torch
        w = np.ascontiguousarray(w, np.float64)
        thr = np.ascontiguousarray(np.atleast_1d(thresholds), np.float64)
        words = self.H * ((self.W + 31) // 32)
        t = torch.empty(words, dtype=torch.int32, device=self.dev)
        _lib.check(self.lib.ct_shard_begin(self.h, C.c_void_p(self.anom.data_ptr()), self.dtype, self.Tl, self.H, self.W,
                                           _lib.ptr(w, _lib._f64p), _lib.ptr(thr, _lib._f64p), len(thr), int(thr_is_f32),
                                           int(op), self.has_prev, C.c_void_p(t.data_ptr()), self._stream()))
        return t

    def launch_threshold(self):
        """Enqueue the thresholding of the own planes now (optional; tables_dev() does it otherwise)."""
        _lib.check(self.lib.ct_shard_launch_threshold(self.h, C.c_void_p(self.out.data_ptr()), self._stream()))

    def boundary_words(self):
        return int(self.lib.ct_shard_boundary_words(self.h))

    def export_boundary(self):
        import torch
        t = torch.empty(self.boundary_words(), dtype=torch.int32, device=self.dev)
        _lib.check(self.lib.ct_shard_export_boundary(self.h, C.cast(C.c_void_p(t.data_ptr()), _lib._u32p), self._stream()))
        return t

    def import_halo(self, t):
        _lib.check(self.lib.ct_shard_import_halo(self.h, C.cast(C.c_void_p(t.data_ptr()), _lib._u32p), self._stream()))

    def tables(self):
        v = ShardView()
        _lib.check(self.lib.ct_shard_tables(self.h, C.cast(C.c_void_p(self.out.data_ptr()), _lib._i32p), self._stream(),
                                            C.byref(v)))
        self.view = view_to_dict(v, self.has_prev, self.t_begin)
        return self.view

    def tables_dev(self):
        """Table kernels; the tables stay on the device.  Returns (counts8 int64 array, export_bytes)."""
        counts = (C.c_long * 8)()
        nbytes = C.c_long(0)
        _lib.check(self.lib.ct_shard_tables_dev(self.h, C.c_void_p(self.out.data_ptr()), self._stream(), counts,
                                                C.byref(nbytes)))
        k = np.array(list(counts), np.int64)
        k[0] = self.t_begin - self.has_prev
        self.counts = k
        return k, int(nbytes.value)

    def export_tables(self, dst):
        _lib.check(self.lib.ct_shard_export_tables(self.h, C.c_void_p(dst.data_ptr()), int(dst.numel()), self._stream()))

    def paint_global(self, g_handle, off):
        _lib.check(self.lib.ct_shard_paint_global(self.h, g_handle, int(off), self.t_begin,
                                                  C.c_void_p(self.out.data_ptr()), self._stream()))
        return self.out

    def plane_runs(self, t_global, off):
        """Row-runs of one own plane with GLOBAL component ids."""
        n = C.c_long(0)
        py, px0, px1, pc = _lib._i32p(), _lib._i32p(), _lib._i32p(), _lib._u32p()
        _lib.check(self.lib.ct_shard_plane_runs(self.h, t_global - self.t_begin + self.has_prev, C.byref(n), C.byref(py),
                                                C.byref(px0), C.byref(px1), C.byref(pc), self._stream()))
        k = n.value
        take = lambda p, d: np.ctypeslib.as_array(p, shape=(k,)).astype(d, copy=True) if k else np.zeros(0, d)  # noqa
        return take(py, np.int32), take(px0, np.int32), take(px1, np.int32), (take(pc, np.int64) + off).astype(np.uint32)

    def paint(self, val_global, overrides, off):
        v = self.view
        local_val = np.zeros(max(v['ncomp'], 1), np.int32)
        nh = v['halo_comps']
        local_val[nh:v['ncomp']] = val_global[nh + off:v['ncomp'] + off]
        mine = [o for o in overrides if self.t_begin <= o[0] < self.t_begin + self.Tl]
        ov = [np.array([o[i] - (self.t_begin if i == 0 else 0) for o in mine], np.int32) for i in range(5)]
        _lib.check(self.lib.ct_shard_paint(self.h, _lib.ptr(local_val, _lib._i32p), len(mine),
                                           *[_lib.ptr(a, _lib._i32p) for a in ov],
                                           C.cast(C.c_void_p(self.out.data_ptr()), _lib._i32p), self._stream()))
        return self.out


def run_contrack_sharded_host(engine, anom_local, t_begin, T_total, w, thresholds, thr_is_f32, op, overlap, persistence,
                              twosided, out=None, group=None):
    """Collective over `group` (default: the world).  anom_local: torch CUDA tensor [T_local, H, W] holding planes
    [t_begin, t_begin + T_local) of the cube; `thresholds`: one value or T_local values (the local slice).
    Returns (flag_local int32 CUDA tensor, n_features, info dict)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    grank = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    sh = Shard(engine, anom_local, t_begin, rank > 0, out)
    sh.threshold(w, thresholds, thr_is_f32, op)
    # ---- the one halo exchange: last plane's bit rows -> next rank ----
    send = sh.export_boundary()
    recv = torch.empty_like(send)
    ops = []
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, send, grank(rank + 1), group))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, recv, grank(rank - 1), group))
    if ops:
        torch.cuda.current_stream(sh.dev).synchronize()
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank > 0:
        sh.import_halo(recv)
    mine = sh.tables()
    views = [unpack_view(b) for b in allgather_bytes(pack_view(mine), group, sh.dev)]
    g, offs = merge_views(views)
    bounds = [(v['t_begin'], v['t_begin'] + v['planes'] - v['has_prev']) for v in views]

    def fetch(t):                                 # collective: every rank replays the same ordered phase
        owner = next(r for r, (a, b) in enumerate(bounds) if a <= t < b)
        obj = [sh.plane_runs(t, offs[rank])] if rank == owner else [None]
        dist.broadcast_object_list(obj, src=grank(owner), group=group)
        return obj[0]

    val, overrides, stats = host_tables_fast(T_total, sh.H, sh.W, w, g, overlap, persistence, twosided, fetch=fetch)
    sh.paint(val, overrides, offs[rank])
    info = dict(stats8=stats, ncomp_global=g['ncomp'], npair_global=g['npair'], nseg_global=g['nseg'],
                table_bytes=[len(pack_view(v)) for v in views], halo_words=sh.boundary_words())
    return sh.out, int(stats[0]), info


def _global_engine(engine):
    """Second context on the same GPU that holds the merged global tables."""
    g = getattr(engine, '_global', None)
    if g is None:
        g = engine._global = type(engine)(engine.device)
    return g


def comp_offsets(counts):
    """Global id of local component 0 of every rank (see merge_views): own components before it minus its halo count."""
    offs, base = [], 0
    for k in counts:
        offs.append(base - int(k[2]))
        base += int(k[1]) - int(k[2])
    return offs


def global_phase(g_engine, counts, gathered, stride, T_total, H, W, w, overlap, persistence, twosided, fetch=None,
                 stream=None):
    """ct_global_merge + ct_global_phase on `g_engine`.  fetch(t) -> (y, x0, x1, comp_global) arrays.  Returns n_features."""
    lib = g_engine.lib
    flat = np.ascontiguousarray(np.asarray(counts, np.int64).reshape(-1))
    w = np.ascontiguousarray(w, np.float64)
    st = C.c_void_p(stream)
    _lib.check(lib.ct_global_merge(g_engine.handle, len(flat) // 8, flat.ctypes.data_as(_lib._longp),
                                   C.c_void_p(gathered.data_ptr()), int(stride), int(T_total), int(H), int(W),
                                   _lib.ptr(w, _lib._f64p), st))
    keep = {}

    def _cb(user, t, n, y, x0, x1, comp):
        try:
            arrs = fetch(int(t))
            arrs = (np.ascontiguousarray(arrs[0], np.int32), np.ascontiguousarray(arrs[1], np.int32),
                    np.ascontiguousarray(arrs[2], np.int32), np.ascontiguousarray(arrs[3], np.uint32))
            keep['a'] = arrs
            n[0] = len(arrs[0])
            y[0] = _lib.ptr(arrs[0], _lib._i32p); x0[0] = _lib.ptr(arrs[1], _lib._i32p)
            x1[0] = _lib.ptr(arrs[2], _lib._i32p); comp[0] = _lib.ptr(arrs[3], _lib._u32p)
            return 0
        except Exception:                      # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return -1

    cb = _FETCH(_cb) if fetch is not None else C.cast(None, _FETCH)
    nfeat = C.c_long(0)
    _lib.check(lib.ct_global_phase(g_engine.handle, float(overlap), int(persistence), int(bool(twosided)),
                                   C.cast(cb, C.c_void_p), None, C.byref(nfeat), st))
    return int(nfeat.value)


_bufs = {}


def _buffer(key, nbytes, dev):
    import torch
    b = _bufs.get(key)
    if b is None or b.numel() < nbytes or b.device != dev:
        b = _bufs[key] = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=dev)
    return b


_hp_streams = {}
_aux_streams = {}


def aux_stream(device):
    """Second high-priority stream: the halo exchange runs on it beside the thresholding of the own planes."""
    import torch
    s = _aux_streams.get(device)
    if s is None:
        s = _aux_streams[device] = torch.cuda.Stream(device=device, priority=-1)
    return s


def high_priority_stream(device):
    """The sharded step runs on a high-priority stream: its zero fill (a cube-sized grid on a lowest-priority side stream)
    would otherwise sit in front of the small table / merge / paint kernels and of the NCCL kernels in the block scheduler.
    Create the process group with ProcessGroupNCCL.Options(is_high_priority_stream=True) for the same reason."""
    import torch
    s = _hp_streams.get(device)
    if s is None:
        s = _hp_streams[device] = torch.cuda.Stream(device=device, priority=-1)
    return s


def run_contrack_sharded(engine, anom_local, t_begin, T_total, w, thresholds, thr_is_f32, op, overlap, persistence,
                         twosided, out=None, group=None):
    """Collective over `group` (default: the world); see _run_contrack_sharded.  Work is enqueued on a high-priority stream
    that is ordered after the caller's current stream; the call returns after the flag planes of this rank are written."""
    import torch
    dev = anom_local.device
    hp = high_priority_stream(dev)
    hp.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(hp):
        res = _run_contrack_sharded(engine, anom_local, t_begin, T_total, w, thresholds, thr_is_f32, op, overlap,
                                    persistence, twosided, out, group)
    torch.cuda.current_stream(dev).wait_stream(hp)
    return res


def _run_contrack_sharded(engine, anom_local, t_begin, T_total, w, thresholds, thr_is_f32, op, overlap, persistence,
                          twosided, out=None, group=None):
    """Collective over `group` (default: the world); tables travel device to device.  anom_local: torch CUDA tensor
    [T_local, H, W] holding planes [t_begin, t_begin + T_local) of the cube; `thresholds`: one value or T_local values
    (the local slice).  Returns (flag_local int32 CUDA tensor, n_features, info dict)."""
    import time
    import torch
    import torch.distributed as dist
    tm = [('start', time.perf_counter())]
    mark = lambda name: tm.append((name, time.perf_counter()))          # noqa: E731  host wall clock between phases
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    grank = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    sh = Shard(engine, anom_local, t_begin, rank > 0, out)
    g_eng = _global_engine(engine)
    # ---- the one halo exchange: the last plane is thresholded first, its bit rows go to the next rank ----
    send = sh.begin(w, thresholds, thr_is_f32, op)
    main = torch.cuda.current_stream(sh.dev)
    boundary_ready = torch.cuda.Event()
    boundary_ready.record(main)
    sh.launch_threshold()                          # the own planes are being thresholded while the halo travels
    mark('boundary_plane')
    recv = torch.empty_like(send)
    aux = aux_stream(sh.dev)
    with torch.cuda.stream(aux):
        aux.wait_event(boundary_ready)
        ops = []
        if rank + 1 < world:
            ops.append(dist.P2POp(dist.isend, send, grank(rank + 1), group))
        if rank > 0:
            ops.append(dist.P2POp(dist.irecv, recv, grank(rank - 1), group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if rank > 0:
            sh.import_halo(recv)                   # on `aux`; the table kernels wait for its event
    mark('halo_exchange')
    # ---- own planes: threshold chunks + table kernels pipelined (device); all-gather of counts and packed tables ----
    counts, nbytes = sh.tables_dev()
    mark('local_tables')
    mine_k = torch.from_numpy(np.append(counts, nbytes)).to(sh.dev)
    all_k = torch.empty((world, 9), dtype=torch.int64, device=sh.dev)
    dist.all_gather_into_tensor(all_k, mine_k, group=group)
    all_k = all_k.cpu().numpy()
    stride = int(all_k[:, 8].max())
    stride = (stride + 255) // 256 * 256
    mine = _buffer(('mine', engine.device), stride, sh.dev)[:stride]
    gathered = _buffer(('all', engine.device), stride * world, sh.dev)[:stride * world]
    sh.export_tables(mine)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    mark('gather_launch')
    offs = comp_offsets(all_k)
    starts = [int(k[0]) + (1 if r > 0 else 0) for r, k in enumerate(all_k)]          # first own plane of every rank

    def fetch(t):                                 # collective: every rank replays the same global phase
        owner = max(r for r in range(world) if starts[r] <= t)
        obj = [sh.plane_runs(t, offs[rank])] if rank == owner else [None]
        dist.broadcast_object_list(obj, src=grank(owner), group=group)
        return obj[0]

    stream = torch.cuda.current_stream(sh.dev).cuda_stream
    nfeat = global_phase(g_eng, all_k[:, :8], gathered, stride, T_total, sh.H, sh.W, w, overlap, persistence, twosided,
                         fetch=fetch, stream=stream)
    mark('global_phase')
    sh.paint_global(g_eng.handle, offs[rank])
    main.wait_stream(aux)                          # (long finished; keeps `send` / `recv` alive until then)
    mark('paint')
    info = dict(table_bytes=[int(b) for b in all_k[:, 8]], halo_words=sh.boundary_words(),
                ncomp_global=int(sum(k[1] - k[2] for k in all_k)),
                phase_ms={b[0]: 1e3 * (b[1] - a[1]) for a, b in zip(tm[:-1], tm[1:])})
    return sh.out, nfeat, info


def run_contrack_sharded_local_dev(engines, anom_parts, T_total, w, thresholds, thr_is_f32, op, overlap, persistence,
                                   twosided, outs=None):
    """The device-table sharded pipeline inside ONE process: `engines[r]` plays rank r (separate contexts on one GPU), the
    all-gather is a concatenation.  Covers ct_shard_tables_dev / export / ct_global_merge / ct_global_phase /
    ct_shard_paint_global on a single GPU."""
    import torch
    t_begin, shards, edges = 0, [], []
    thr = np.atleast_1d(np.asarray(thresholds, np.float64))
    for r, (e, a) in enumerate(zip(engines, anom_parts)):
        sh = Shard(e, a, t_begin, r > 0, outs[r] if outs is not None else None)
        edges.append(sh.begin(w, thr if len(thr) == 1 else thr[t_begin:t_begin + sh.Tl], thr_is_f32, op))
        t_begin += sh.Tl
        shards.append(sh)
    for r in range(1, len(shards)):
        shards[r].import_halo(edges[r - 1])
    import time
    torch.cuda.synchronize()
    ks, ms_tables = [], []
    for sh in shards:
        t0 = time.perf_counter()
        ks.append(sh.tables_dev())
        ms_tables.append(1e3 * (time.perf_counter() - t0))
    counts = np.stack([k for k, _ in ks])
    stride = (max(n for _, n in ks) + 255) // 256 * 256
    gathered = torch.zeros(stride * len(shards), dtype=torch.uint8, device=shards[0].dev)
    for r, sh in enumerate(shards):
        sh.export_tables(gathered[r * stride:(r + 1) * stride])
    torch.cuda.synchronize()
    offs = comp_offsets(counts)
    starts = [sh.t_begin for sh in shards]

    def fetch(t):
        owner = max(r for r in range(len(shards)) if starts[r] <= t)
        return shards[owner].plane_runs(t, offs[owner])

    g_eng = _global_engine(engines[0])
    stream = torch.cuda.current_stream(shards[0].dev).cuda_stream
    t0 = time.perf_counter()
    nfeat = global_phase(g_eng, counts, gathered, stride, T_total, shards[0].H, shards[0].W, w, overlap, persistence,
                         twosided, fetch=fetch, stream=stream)
    ms_global = 1e3 * (time.perf_counter() - t0)
    outs, ms_paint = [], []
    for r, sh in enumerate(shards):
        t0 = time.perf_counter()
        outs.append(sh.paint_global(g_eng.handle, offs[r]))
        ms_paint.append(1e3 * (time.perf_counter() - t0))
    return outs, nfeat, dict(counts=counts, offsets=offs, stride=stride, ms_tables=ms_tables, ms_global=ms_global,
                             ms_paint=ms_paint, global_stats=g_eng.stats())


def run_contrack_sharded_local(engines, anom_parts, T_total, w, thresholds, thr_is_f32, op, overlap, persistence,
                               twosided):
    """The same sharded pipeline inside ONE process: `engines[r]` (separate contexts, possibly on the same GPU) plays
    rank r, the exchanges are plain tensor hand-overs.  Used by the single-GPU tests of the sharded kernels' plumbing."""
    t_begin, shards = 0, []
    thr = np.atleast_1d(np.asarray(thresholds, np.float64))
    for r, (e, a) in enumerate(zip(engines, anom_parts)):
        sh = Shard(e, a, t_begin, r > 0)
        sh.threshold(w, thr if len(thr) == 1 else thr[t_begin:t_begin + sh.Tl], thr_is_f32, op)
        t_begin += sh.Tl
        shards.append(sh)
    for r in range(1, len(shards)):
        shards[r].import_halo(shards[r - 1].export_boundary())
    views = [unpack_view(pack_view(sh.tables())) for sh in shards]
    g, offs = merge_views(views)
    bounds = [(sh.t_begin, sh.t_begin + sh.Tl) for sh in shards]

    def fetch(t):
        owner = next(r for r, (a, b) in enumerate(bounds) if a <= t < b)
        return shards[owner].plane_runs(t, offs[owner])

    val, overrides, stats = host_tables_fast(T_total, shards[0].H, shards[0].W, w, g, overlap, persistence, twosided,
                                             fetch=fetch)
    outs = [sh.paint(val, overrides, offs[r]) for r, sh in enumerate(shards)]
    return outs, int(stats[0]), dict(stats8=stats, views=views)
