"""Time-sharded run_contrack over several GPUs: one process (or host thread) per GPU, one collective call per rank.

SURVEY.md 8(e): every rank owns a contiguous range of time planes.  Thresholding, 2-D labelling, area tables and painting
are local; what crosses ranks is
  * ONE boundary plane per neighbour pair: the bit rows of a rank's last plane go to the next rank (130 KB at 721x1440),
    so that rank can build the (plane t, plane t-1) pair tables across the cut, and
  * the component / class / pair / date-line tables (a few MB in total), ALL-GATHERED device to device with one collective
    and merged by a kernel into global tables on every rank, which then replays the global part of the path (keep/kill
    recurrence, 3-D numbering, stale-box date-line merge, persistence) on its copy and obtains the same global ids -- the
    "global relabel" step; no rank-dependent numbering exists.
The cube itself never moves.  All of it happens inside the library (``ct_run_contrack_sharded``, csrc/ct_dist.cu; NCCL bound
with dlopen): this module only creates the communicator and passes buffers across the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def shard_bounds(T, world):
    """Contiguous, near-equal time ranges; every rank gets at least one plane (world <= T)."""
    edges = [(T * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


class Comm(object):
    """A ``ct_comm``: the collectives of the sharded entry points (NCCL, or an in-process group for single-GPU tests)."""

    def __init__(self, handle, rank, size):
        self.handle, self.rank, self.size = handle, int(rank), int(size)

    @classmethod
    def nccl(cls, rank, size, device, bcast):
        """NCCL communicator.  ``bcast(bytes_or_None) -> bytes`` distributes rank 0's 128-byte unique id to all ranks
        (any side channel will do: torch.distributed, MPI, a file)."""
        lib = _lib.load()
        uid = (C.c_ubyte * 128)()
        if rank == 0:
            _lib.check(lib.ct_nccl_unique_id(uid))
        raw = bcast(bytes(uid) if rank == 0 else None)
        uid = (C.c_ubyte * 128).from_buffer_copy(raw)
        h = C.c_void_p()
        _lib.check(lib.ct_comm_init_nccl(uid, int(rank), int(size), int(device), C.byref(h)))
        return cls(h, rank, size)

    @classmethod
    def from_torch(cls, device, group=None):
        """NCCL communicator spanning the ranks of a torch.distributed group (the unique id travels through the group)."""
        import torch.distributed as dist
        rank, size = dist.get_rank(group), dist.get_world_size(group)
        src = dist.get_global_rank(group, 0) if group is not None else 0

        def bcast(b):
            obj = [b]
            dist.broadcast_object_list(obj, src=src, group=group)
            return obj[0]
        return cls.nccl(rank, size, device, bcast)

    @classmethod
    def local_group(cls, size):
        """``size`` communicators of an in-process group (one host thread per rank): single-GPU tests of the sharded path."""
        lib = _lib.load()
        arr = (C.c_void_p * size)()
        _lib.check(lib.ct_comm_init_local(int(size), arr))
        return [cls(C.c_void_p(arr[r]), r, size) for r in range(size)]

    def close(self):
        if self.handle:
            _lib.load().ct_comm_destroy(self.handle)
            self.handle = None


_comms = {}


def default_comm(engine, group=None):
    """The (cached) NCCL communicator of this process for a torch.distributed group."""
    key = (engine.device, id(group))
    c = _comms.get(key)
    if c is None:
        c = _comms[key] = Comm.from_torch(engine.device, group)
    return c


def run_contrack_sharded(engine, anom_local, t_begin, T_total, w, thresholds, thr_is_f32, op, overlap, persistence,
                         twosided, out=None, comm=None, group=None):
    """Collective over ``comm`` (default: an NCCL communicator over the torch.distributed world / ``group``).
    anom_local: torch CUDA tensor (device resident) or numpy array (host buffers, ideally page-locked) [T_local, H, W] holding
    planes [t_begin, t_begin + T_local) of the cube; ``thresholds``: one value or T_local values (the local slice).  Returns
    (flag_local int32, same kind as the input; n_features; stats dict)."""
    if comm is None:
        comm = default_comm(engine, group)
    w = np.ascontiguousarray(w, np.float64)
    thr = np.ascontiguousarray(np.atleast_1d(thresholds), np.float64)
    nfeat = C.c_long(0)
    if isinstance(anom_local, np.ndarray):
        # host buffers: the shard streams through the GPU, the flag planes come back as a row-run table (see the header)
        a = anom_local if anom_local.dtype in (np.float32, np.float64) else anom_local.astype(np.float64)
        a = np.ascontiguousarray(a)
        Tl, H, W = (int(s) for s in a.shape)
        if out is None:
            out = np.empty((Tl, H, W), np.int32)
        _lib.check(engine.lib.ct_run_contrack_sharded_host(
            engine.handle, comm.handle, C.c_void_p(a.ctypes.data), _lib.CT_F32 if a.dtype == np.float32 else _lib.CT_F64, Tl,
            int(t_begin), int(T_total), H, W, _lib.ptr(w, _lib._f64p), _lib.ptr(thr, _lib._f64p), len(thr), int(thr_is_f32),
            int(op), float(overlap), int(persistence), int(bool(twosided)), C.c_void_p(out.ctypes.data), C.byref(nfeat), 0))
        return out, int(nfeat.value), engine.stats()
    import torch
    x = anom_local if anom_local.is_contiguous() else anom_local.contiguous()
    if not x.is_cuda or x.device.index != engine.device:
        raise ValueError('the shard must live on cuda:%d' % engine.device)
    dt = {torch.float32: _lib.CT_F32, torch.float64: _lib.CT_F64}.get(x.dtype)
    if dt is None:
        raise TypeError('input dtype must be float32 or float64')
    Tl, H, W = (int(s) for s in x.shape)
    if out is None:
        out = torch.empty((Tl, H, W), dtype=torch.int32, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(engine.lib.ct_run_contrack_sharded(
        engine.handle, comm.handle, C.c_void_p(x.data_ptr()), dt, Tl, int(t_begin), int(T_total), H, W,
        _lib.ptr(w, _lib._f64p), _lib.ptr(thr, _lib._f64p), len(thr), int(thr_is_f32), int(op), float(overlap),
        int(persistence), int(bool(twosided)), C.c_void_p(out.data_ptr()), C.byref(nfeat), C.c_void_p(stream)))
    return out, int(nfeat.value), engine.stats()


def quantile_time_sharded(engine, x_local, q, y0=0, y1=None, comm=None, group=None):
    """Quantiles over ALL time steps of a time-sharded cube (README.rst:150-151): collective, every rank gets the result."""
    return engine.quantile_time(x_local, q, y0, y1, comm=comm if comm is not None else default_comm(engine, group))


def run_local_group(engines, anom_parts, T_total, w, thresholds, thr_is_f32, op, overlap, persistence, twosided, comms=None,
                    outs=None):
    """The sharded call on an in-process group: engines[r] plays rank r from its own host thread (the contexts may share one
    GPU).  Returns ([flag_r], n_features, [stats_r]).  Single-GPU tests and debugging.  `comms`: communicators of an earlier
    Comm.local_group(n) to reuse (the negotiated exchange capacities and the peer windows live with the communicator)."""
    import threading
    import torch
    n = len(engines)
    own_comms = comms is None
    if own_comms:
        comms = Comm.local_group(n)
    thr = np.atleast_1d(np.asarray(thresholds, np.float64))
    bounds = np.cumsum([0] + [int(a.shape[0]) for a in anom_parts])
    res, errs = [None] * n, [None] * n

    def work(r):
        try:
            torch.cuda.set_device(engines[r].device)
            with torch.cuda.stream(torch.cuda.Stream(device=engines[r].device)):
                t = thr if len(thr) == 1 else thr[bounds[r]:bounds[r + 1]]
                res[r] = run_contrack_sharded(engines[r], anom_parts[r], int(bounds[r]), T_total, w, t, thr_is_f32, op,
                                              overlap, persistence, twosided, comm=comms[r],
                                              out=None if outs is None else outs[r])
                torch.cuda.current_stream().synchronize()
        except BaseException as e:                      # noqa: BLE001  (reported by the caller; a dead rank would hang the rest)
            errs[r] = e
    threads = [threading.Thread(target=work, args=(r,)) for r in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if own_comms:
        for c in comms:
            c.close()
    for e in errs:
        if e is not None:
            raise e
    feats = {r[1] for r in res}
    if len(feats) != 1:
        raise RuntimeError('ranks disagree on the number of features: %s' % sorted(feats))
    return [r[0] for r in res], res[0][1], [r[2] for r in res]
