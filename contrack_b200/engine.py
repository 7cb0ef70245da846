"""Host-side driver of the C-ABI library: owns one ``ct_ctx`` per GPU and moves numpy / torch buffers across the ABI.

PyTorch is used only as plumbing (device memory, streams); every computation happens in the library's CUDA kernels and
its native host table phase.  There is no CPU fallback: without the library or without a B200 the calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ContrackLibError

_STAT_KEYS = ('runs', 'comps2d', 'pairs', 'seam_rows', 'kept_comps', 'labels3d', 'features', 'seam_events',
              'seam_splits', 'neartie_resolved', 'override_runs', 'special_rows', 'kernel_launches', 'ms_threshold',
              'ms_tables_gpu', 'ms_tables_host_roundtrip', 'ms_host_tables', 'ms_paint', 'ms_total', 'ms_h2d_threshold',
              'ms_tables', 'ms_paint_d2h', 'ms_zero_fill', 'seam_segments', 'sweeps', 'neartie_flagged', 'h2d_bytes',
              'd2h_bytes', 'host_sparse', 'host_threads', 'chunks', 'ms_h_chunks', 'ms_h_global', 'ms_g_sweeps', 'ms_g_link',
              'ms_g_labels_d2h', 'ms_tables_after_threshold', 'moved_comps', 'ms_lc_rows', 'ms_lc_tables_sums', 'lc_rows',
              'lc_rolled', 'ms_ht_init', 'ms_ht_events', 'ms_ht_persist', 'ht_walked', 'label_fast', 'slot_overflow', 'fast_path', 'plane_attempts', 'wavefront_planes',
              'ms_g_kernel', 'event_segments', 'ms_h_tables', 'shard_attempts', 'exchange_bytes', 'exchange_negotiated', 'ms_global_kernel', 'ms_plane_kernel', 'ms_exchange', 'p2p') + tuple('ms_h_c%d' % i for i in range(8)) + tuple(
    'ms_t_' + k for k in ('row_scans', 'extract_runs', 'ccl_union', 'ccl_flatten', 'root_scan', 'comp_accumulate', 'seam_segs',
                          'class_sums', 'pairs_accumulate', 'pairs_csr'))


def _is_torch(x):
    return hasattr(x, 'data_ptr') and hasattr(x, 'device')


class Engine(object):
    """One context (scratch memory, streams) on one GPU."""

    _cache = {}

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.ct_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)

    @classmethod
    def get(cls, device=0):
        e = cls._cache.get(device)
        if e is None:
            e = cls._cache[device] = Engine(device)
        return e

    def close(self):
        if self.handle:
            self.lib.ct_destroy(self.handle)
            self.handle = None
            Engine._cache.pop(self.device, None)

    def set_option(self, key, value):
        _lib.check(self.lib.ct_set_option(self.handle, key.encode(), int(value)))

    def stats(self):
        out = {}
        for k in _STAT_KEYS:
            v = self.lib.ct_get_stat(self.handle, k.encode())
            if v >= 0:
                out[k] = v
        return out

    # ------------------------------------------------------------------------------------------------------------
    def run_contrack(self, anom, w, thresholds, thr_is_f32, op, overlap, persistence, twosided, stage=0, out=None,
                     chunk_planes=0):
        """anom: [T,H,W] C-contiguous float32/float64, numpy (host) or torch CUDA tensor (device).
        Returns (flag, n_features); flag is int32 of the same kind (numpy / torch CUDA) as the input."""
        T, H, W = (int(s) for s in anom.shape)
        w = np.ascontiguousarray(w, np.float64)
        thr = np.ascontiguousarray(np.atleast_1d(thresholds), np.float64)
        if w.shape != (H,):
            raise ValueError('weights must have shape (H,)')
        nfeat = C.c_long(0)
        if _is_torch(anom):
            import torch
            if not anom.is_cuda:
                raise ValueError('torch input must live on a CUDA device (pass numpy for host data)')
            if anom.device.index != self.device:
                raise ValueError('input is on %s, engine on cuda:%d' % (anom.device, self.device))
            if not anom.is_contiguous():
                anom = anom.contiguous()
            dt = {torch.float32: _lib.CT_F32, torch.float64: _lib.CT_F64}.get(anom.dtype)
            if dt is None:
                raise TypeError('input dtype must be float32 or float64')
            if out is None:
                out = torch.empty((T, H, W), dtype=torch.int32, device=anom.device)
            stream = torch.cuda.current_stream(anom.device).cuda_stream
            rc = self.lib.ct_run_contrack(self.handle, C.c_void_p(anom.data_ptr()), dt, T, H, W,
                                          _lib.ptr(w, _lib._f64p), _lib.ptr(thr, _lib._f64p), len(thr), int(thr_is_f32),
                                          int(op), float(overlap), int(persistence), int(bool(twosided)),
                                          C.c_void_p(out.data_ptr()), C.byref(nfeat), int(stage), C.c_void_p(stream))
            _lib.check(rc)
            return out, nfeat.value
        a = np.asarray(anom)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        a = np.ascontiguousarray(a)
        dt = _lib.CT_F32 if a.dtype == np.float32 else _lib.CT_F64
        if stage != 0:
            raise ValueError('debug stages are only available for device-resident input')
        if out is None:
            # np.empty + the library's zeroing pass: a fresh np.zeros (option "host_out_zeroed") skips the pass but pays for
            # it in page faults when the runs are expanded (measured 145 ms instead of 15 ms at 2707 x 721 x 1440)
            out = np.empty((T, H, W), np.int32)
        rc = self.lib.ct_run_contrack_host(self.handle, C.c_void_p(a.ctypes.data), dt, T, H, W, _lib.ptr(w, _lib._f64p),
                                           _lib.ptr(thr, _lib._f64p), len(thr), int(thr_is_f32), int(op), float(overlap),
                                           int(persistence), int(bool(twosided)), C.c_void_p(out.ctypes.data),
                                           C.byref(nfeat), int(chunk_planes))
        _lib.check(rc)
        return out, nfeat.value


    # ------------------------------------------------------------------------------------------------------------
    def _to_device(self, x, like=None):
        """(cuda tensor, was_host).  float32 and float64 keep their precision (xarray computes a float64 cube in float64:
        climatology, anomaly and quantiles of packed ERA5 decoded with scale/offset are float64 in the reference); anything
        else is converted to float64, numpy's result type for integer input.  `like`: another tensor whose dtype wins if it
        is the wider one (numpy promotion of `x - clim`)."""
        import torch
        if _is_torch(x):
            if not x.is_cuda:
                raise ValueError('torch input must live on a CUDA device (pass numpy for host data)')
            t, was_host = x, False
        else:
            a = np.asarray(x)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            t, was_host = torch.from_numpy(np.ascontiguousarray(a)).to('cuda:%d' % self.device), True
        if t.dtype not in (torch.float32, torch.float64):
            t = t.double()
        if like is not None and like.dtype == torch.float64 and t.dtype != torch.float64:
            t = t.double()
        return t.contiguous(), was_host

    @staticmethod
    def _ct_dtype(t):
        import torch
        return _lib.CT_F64 if t.dtype == torch.float64 else _lib.CT_F32

    def calc_clim(self, z, group_index, ngroups, window):
        """z [T,H,W] float32 (numpy or torch CUDA); group_index [T] int 0..ngroups-1.  Returns clim [G,H,W] of the same
        kind as z (contrack.py:483-489)."""
        import torch
        zd, was_host = self._to_device(z)
        T, H, W = (int(s) for s in zd.shape)
        g = np.ascontiguousarray(group_index, np.int32)
        if g.shape != (T,):
            raise ValueError('group_index must have shape (T,)')
        clim = torch.empty((int(ngroups), H, W), dtype=zd.dtype, device=zd.device)
        stream = torch.cuda.current_stream(zd.device).cuda_stream
        _lib.check(self.lib.ct_calc_clim_t(self.handle, C.c_void_p(zd.data_ptr()), self._ct_dtype(zd), T, H, W,
                                           _lib.ptr(g, _lib._i32p), int(ngroups), int(window), C.c_void_p(clim.data_ptr()),
                                           C.c_void_p(stream)))
        return clim.cpu().numpy() if was_host else clim

    def calc_anom(self, z, group_index, ngroups, clim, smooth, out=None):
        """anom [T,H,W] = centred rolling mean (window `smooth`) of z[t] - clim[group_index[t]] (contrack.py:568-570)."""
        import torch
        zd, was_host = self._to_device(z)
        cd, _ = self._to_device(clim, like=zd)
        if cd.dtype != zd.dtype:                       # float32 z against a float64 climatology: numpy promotes to float64
            zd = zd.double()
        T, H, W = (int(s) for s in zd.shape)
        g = np.ascontiguousarray(group_index, np.int32)
        if out is None or was_host or out.dtype != zd.dtype:
            outd = torch.empty((T, H, W), dtype=zd.dtype, device=zd.device)
        else:
            outd = out
        stream = torch.cuda.current_stream(zd.device).cuda_stream
        _lib.check(self.lib.ct_calc_anom_t(self.handle, C.c_void_p(zd.data_ptr()), self._ct_dtype(zd), T, H, W,
                                           _lib.ptr(g, _lib._i32p), int(ngroups), C.c_void_p(cd.data_ptr()), int(smooth),
                                           C.c_void_p(outd.data_ptr()), C.c_void_p(stream)))
        return outd.cpu().numpy() if was_host else outd

    # ---- callers either side of the path (ct_extras.cu) -----------------------------------------------
    def quantile_time(self, x, q, y0=0, y1=None, comm=None):
        """np.nanquantile(x[:, y0:y1, :], q, axis=0) ('linear') of a float32 / float64 cube (numpy or torch CUDA): float64
        [len(q), y1-y0, W], same kind as x (README.rst:150-151).  With `comm` (contrack_b200.sharded.Comm) x holds this
        rank's time steps of a time-sharded cube and every rank receives the quantiles over ALL time steps (collective)."""
        import torch
        xd, was_host = self._to_device(x)
        T, H, W = (int(s) for s in xd.shape)
        y1 = H if y1 is None else int(y1)
        qa = np.ascontiguousarray(np.atleast_1d(q), np.float64)
        if ((qa < 0) | (qa > 1) | np.isnan(qa)).any():
            raise ValueError('Quantiles must be in the range [0, 1]')
        out = torch.empty((len(qa), y1 - int(y0), W), dtype=torch.float64, device=xd.device)
        stream = torch.cuda.current_stream(xd.device).cuda_stream
        _lib.check(self.lib.ct_quantile_time_t(self.handle, comm.handle if comm is not None else None,
                                               C.c_void_p(xd.data_ptr()), self._ct_dtype(xd), T, H, W, int(y0), y1,
                                               _lib.ptr(qa, _lib._f64p), len(qa), C.c_void_p(out.data_ptr()),
                                               C.c_void_p(stream)))
        return out.cpu().numpy() if was_host else out

    def flag_count(self, flag, greater_than=1):
        """(flag > greater_than).sum(axis=0) as int32 [H, W] (README.rst:161)."""
        import torch
        if _is_torch(flag):
            if not flag.is_cuda:
                raise ValueError('torch input must live on a CUDA device (pass numpy for host data)')
            fd, was_host = flag.contiguous().to(torch.int32), False
        else:
            fd, was_host = torch.from_numpy(np.ascontiguousarray(np.asarray(flag), np.int32)).to('cuda:%d' % self.device), True
        T, H, W = (int(s) for s in fd.shape)
        out = torch.empty((H, W), dtype=torch.int32, device=fd.device)
        stream = torch.cuda.current_stream(fd.device).cuda_stream
        _lib.check(self.lib.ct_flag_count(self.handle, C.c_void_p(fd.data_ptr()), T, H, W, int(greater_than),
                                          C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
        return out.cpu().numpy() if was_host else out

    def divide(self, x, divisor):
        """x / float32(divisor) in float32 (contrack.py:417-419)."""
        import torch
        xd, was_host = self._to_device(x)
        if xd.dtype != torch.float32:
            raise TypeError('divide() is the float32 kernel of calculate_gph_from_gp; float64 data divide on the caller side')
        out = torch.empty_like(xd)
        stream = torch.cuda.current_stream(xd.device).cuda_stream
        _lib.check(self.lib.ct_divide_f32(self.handle, C.c_void_p(xd.data_ptr()), int(xd.numel()), float(np.float32(divisor)),
                                          C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
        return out.cpu().numpy() if was_host else out

    def gather_planes(self, src, iy, ix):
        """dst[g, y, x] = src[g, iy[y], ix[x]] (nearest-neighbour regrid with the caller's index maps, contrack.py:565)."""
        import torch
        sd, was_host = self._to_device(src)
        G, Hs, Ws = (int(s) for s in sd.shape)
        iy = np.ascontiguousarray(iy, np.int32)
        ix = np.ascontiguousarray(ix, np.int32)
        out = torch.empty((G, len(iy), len(ix)), dtype=sd.dtype, device=sd.device)
        stream = torch.cuda.current_stream(sd.device).cuda_stream
        _lib.check(self.lib.ct_gather_planes_t(self.handle, C.c_void_p(sd.data_ptr()), self._ct_dtype(sd), G, Hs, Ws,
                                               _lib.ptr(iy, _lib._i32p), _lib.ptr(ix, _lib._i32p), len(iy), len(ix),
                                               C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
        return out.cpu().numpy() if was_host else out

    # ------------------------------------------------------------------------------------------------------------
    def run_lifecycle(self, flag, var, w):
        """flag [T,H,W] integer, var [T,H,W] float32/float64 (numpy or torch CUDA, C-order time/lat/lon), w [H] float64 row
        weights.  Returns a dict of numpy arrays with one entry per (time step, flag id) in no particular order -- see
        ct_run_lifecycle in include/contrack_b200.h (contrack.py:799-907)."""
        import torch
        dev = 'cuda:%d' % self.device

        def to_dev(x, dtypes, default):
            if _is_torch(x):
                if not x.is_cuda:
                    raise ValueError('torch input must live on a CUDA device (pass numpy for host data)')
                t = x
            else:
                a = np.asarray(x)
                t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            if t.dtype not in dtypes:
                t = t.to(default)
            return t.contiguous()
        fd = to_dev(flag, (torch.int32,), torch.int32)
        vd = to_dev(var, (torch.float32, torch.float64), torch.float64)
        if fd.shape != vd.shape or fd.dim() != 3:
            raise ValueError('flag and variable must both have shape (time, lat, lon)')
        T, H, W = (int(s) for s in fd.shape)
        w = np.ascontiguousarray(w, np.float64)
        if w.shape != (H,):
            raise ValueError('weights must have shape (H,)')
        n = C.c_long(0)
        stream = torch.cuda.current_stream(fd.device).cuda_stream
        _lib.check(self.lib.ct_run_lifecycle(self.handle, C.c_void_p(fd.data_ptr()), C.c_void_p(vd.data_ptr()),
                                             _lib.CT_F64 if vd.dtype == torch.float64 else _lib.CT_F32, T, H, W,
                                             _lib.ptr(w, _lib._f64p), C.byref(n), C.c_void_p(stream)))
        n = n.value
        out = {k: np.zeros(n, np.int32) for k in ('t', 'label', 'npix', 'roll')}
        out.update({k: np.zeros(n, np.float64) for k in ('area', 'wsum', 'norm', 'sy', 'sx')})
        _lib.check(self.lib.ct_lifecycle_fetch(self.handle, n, *[_lib.ptr(out[k], _lib._i32p) for k in
                                                                  ('t', 'label', 'npix', 'roll')],
                                               *[_lib.ptr(out[k], _lib._f64p) for k in
                                                 ('area', 'wsum', 'norm', 'sy', 'sx')]))
        return out


__all__ = ['Engine', 'ContrackLibError']
