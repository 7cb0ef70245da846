"""Times the ordered host table phase (ct_host_tables) on tables dumped from a GPU run (CT_DUMP_TABLES=file)."""
import ctypes as C, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrack_b200 import _lib as L


def load(path):
    raw = open(path, 'rb').read()
    hdr = np.frombuffer(raw, np.int64, 8)
    T, H, W, nc, npair, ns = (int(v) for v in hdr[:6])
    off = [64]

    def take(dt, n):
        a = np.frombuffer(raw, dt, n, off[0]).copy()
        off[0] += a.nbytes
        return a
    d = dict(T=T, H=H, W=W, nc=nc, np=npair, ns=ns, w=take(np.float64, H))
    for k in ('t', 'y0', 'y1', 'x0', 'x1'):
        d[k] = take(np.int32, nc)
    d['cls'] = take(np.uint32, nc); d['E'] = take(np.float64, nc); d['S'] = take(np.float64, nc); d['nsp'] = take(np.uint32, nc)
    for k in ('pa', 'pb', 'pn', 'pnsp'):
        d[k] = take(np.uint32, npair)
    d['pE'] = take(np.float64, npair); d['pS'] = take(np.float64, npair)
    for k in ('sr', 'sa', 'sb'):
        d[k] = take(np.uint32, ns)
    return d


def run(d, overlap=0.5, persistence=5, twosided=1, reps=5):
    lib = L.load()
    val = np.zeros(d['nc'] + 1, np.int32)
    ovr = [np.zeros(1024, np.int32) for _ in range(5)]
    n_ovr = C.c_long(0)
    stats = (C.c_long * 8)()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = lib.ct_host_tables(d['T'], d['H'], d['W'], L.ptr(d['w'], L._f64p), overlap, persistence, twosided, 0,
                                d['nc'], *[L.ptr(d[k], L._i32p) for k in ('t', 'y0', 'y1', 'x0', 'x1')],
                                L.ptr(d['cls'], L._u32p), L.ptr(d['E'], L._f64p), L.ptr(d['S'], L._f64p), L.ptr(d['nsp'], L._u32p),
                                d['np'], *[L.ptr(d[k], L._u32p) for k in ('pa', 'pb', 'pn', 'pnsp')], L.ptr(d['pE'], L._f64p),
                                L.ptr(d['pS'], L._f64p), d['ns'], *[L.ptr(d[k], L._u32p) for k in ('sr', 'sa', 'sb')],
                                None, None, None, None, None, L.ptr(val, L._i32p), 1024, *[L.ptr(o, L._i32p) for o in ovr],
                                C.byref(n_ovr), stats)
        best = min(best, time.perf_counter() - t0)
        if rc != 0:
            print('rc', rc, lib.ct_last_error())
            break
    return best, list(stats), val


if __name__ == '__main__':
    d = load(sys.argv[1])
    print({k: d[k] for k in ('T', 'nc', 'np', 'ns')})
    best, st, val = run(d)
    print('host phase best %.3f ms' % (best * 1e3), st)
    import hashlib
    print('comp_val sha', hashlib.sha256(val.tobytes()).hexdigest()[:16])
