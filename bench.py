#!/usr/bin/env python3
"""bench.py -- timesteps/s of run_contrack on a synthetic 721x1440 Z500-anomaly cube (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--T 10957] [--impl reference]

One "step" = one complete run_contrack pass (threshold -> 2-D labelling -> overlap filter -> 3-D tracking -> persistence
-> int32 flag cube) over the whole [T, 721, 1440] cube.  `value` = T*K / (CUDA-event time of K steps), inputs and
outputs resident in HBM; `e2e` = the same pass through the host-buffer entry point (pinned host float32 in, int32 out,
copies inside the timed region); `roofline` = the dominant kernel against MEASURED_PEAKS.json; `cpu_baseline` = the
reference algorithm (oracle/: same scipy calls, same loops) on a bounded sub-cube on this box's host, 1 thread.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# NCCL_DEBUG / NCCL_DEBUG_FILE are left exactly as the caller set them.  Under NCCL_DEBUG=INFO|VERSION without a
# NCCL_DEBUG_FILE NCCL logs to stdout: in a multi-rank run file descriptor 1 is therefore pointed at stderr while the
# job runs and the ONE JSON line is written to the real stdout at the end (see main()).

H, W = 721, 1440
SIGMA = (2.5, 24.0, 40.0)        # (2.5 steps, 6 deg, 10 deg) at 0.25 deg
THRESHOLD, GORL, OVERLAP, PERSISTENCE, TWOSIDED = 160, '>=', 0.5, 5, True
SEED = 2


def grid():
    lat = np.linspace(90, -90, H).astype(np.float32)
    lon = (np.arange(W) * (360.0 / W)).astype(np.float32)
    return lat, lon


def reference_weights(lat, lon):
    """contrack.py:703-704 with dlat = dlon = 0.25 (set_up(force=True) on the float32 linspace grid)."""
    weight_lat = np.cos(lat * np.pi / 180)
    w = np.array((111 * np.float32(0.25) * 111 * np.float32(0.25) * weight_lat)).astype(np.float32)
    return w.astype(np.float64)


_synth = None


def synth_lib():
    global _synth
    if _synth is None:
        _synth = C.CDLL(os.path.join(ROOT, 'bench_support', 'libct_synth.so'))
        _synth.ct_synth_fill.restype = C.c_int
        _synth.ct_synth_fill.argtypes = [C.c_void_p, C.c_ulonglong, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int,
                                         C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_void_p]
        _synth.ct_checksum_i32.restype = C.c_int
        _synth.ct_checksum_i32.argtypes = [C.c_void_p, C.c_size_t, C.c_ulonglong, C.c_void_p, C.c_void_p]
    return _synth


def synth_fill(out, t0, T_total, seed=SEED, season=False):
    """Fill the CUDA float32 tensor out[nt, H, W] with planes [t0, t0+nt) of the synthetic cube (bench_support/)."""
    import torch
    synth_lib()
    nt, h, w = out.shape
    rc = _synth.ct_synth_fill(C.c_void_p(out.data_ptr()), seed, t0, nt, T_total, h, w, SIGMA[0],
                              SIGMA[1] * h / 721.0, SIGMA[2] * w / 1440.0, 100.0, 60.0 if season else 0.0, 365.25,
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError('ct_synth_fill failed: cudaError %d' % rc)
    return out


def flag_checksum(flag, index0):
    """64-bit position-weighted checksum of an int32 CUDA tensor taken as cells index0.. of the whole cube (additive over
    disjoint parts: the sum over the time shards modulo 2^64 is the checksum of the cube)."""
    import torch
    synth_lib()
    out = torch.zeros(1, dtype=torch.int64, device=flag.device)
    rc = _synth.ct_checksum_i32(C.c_void_p(flag.data_ptr()), flag.numel(), index0, C.c_void_p(out.data_ptr()),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError('ct_checksum_i32 failed: cudaError %d' % rc)
    return out


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')] + [time.time()])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(2)
            except Exception:
                self.proc.kill()

    def summary(self, t0=None, t1=None):
        """Samples inside [t0, t1] (the timed region); when the region is shorter than the sampling period, the samples of
        the whole window the sampler was running (warm-up + timed steps, the same load) and `window` says so."""
        rows, window = self.rows, 'sampler lifetime'
        if t0 is not None:
            inside = [r for r in self.rows if t0 - 0.02 <= r[-1] <= t1 + 0.02]
            rows, window = (inside, 'timed region') if inside else (self.rows, 'warm-up + timed region (timed region '
                                                                    'shorter than the 20 ms sampling period)')
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0, 'window': window}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm), 'window': window}


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


def ncu_traffic(kernel, cells):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/summarise_ncu.py), scaled from the captured launch to `cells` cells."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            d = json.load(f)
        return d['kernels'][kernel]['bytes_per_cell'] * cells, d['source']
    except Exception:
        return None, None


def cpu_reference_run(x, lat, lon):
    """The reference algorithm on the host (oracle restatement: same scipy.ndimage calls, same Python loops)."""
    from oracle import contrack_oracle as oracle
    t0 = time.perf_counter()
    f = oracle.run_contrack(x, lat, lon, THRESHOLD, GORL, OVERLAP, PERSISTENCE, TWOSIDED, force=True)
    return f, time.perf_counter() - t0


def make_sample(T_sub, use_gpu, season=False):
    """[T_sub, H, W] float32 host array of the benchmark's synthetic field (first T_sub planes of a T_sub-long cube)."""
    if use_gpu:
        import torch
        d = torch.empty((T_sub, H, W), dtype=torch.float32, device='cuda')
        synth_fill(d, 0, T_sub, season=season)
        torch.cuda.synchronize()
        return d.cpu().numpy()
    from _synth import synth_cube
    x = synth_cube(SEED, T_sub, H, W, SIGMA)
    if season:
        lat = np.linspace(90, -90, H)[None, :, None] * np.pi / 180
        x = (x + 5500.0 + 60.0 * np.sin(lat) * np.cos(2 * np.pi * np.arange(T_sub)[:, None, None] / 365.25)).astype(np.float32)
    return x


def run_reference_arm(args):
    """--impl reference: the reference's CPU path (oracle port; the reference itself cannot be imported here: xarray is
    missing) on bounded samples of the same workload, one sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    try:
        import torch
        use_gpu = torch.cuda.is_available()
    except Exception:
        use_gpu = False
    T_sub = args.cpu_T
    lat, lon = grid()
    x = make_sample(T_sub, use_gpu, season=(args.config == 3))

    def one():
        if args.config != 3:
            return cpu_reference_run(x, lat, lon)[1]
        from oracle import contrack_oracle as oracle
        doy, _, _ = day_groups(T_sub)
        t0 = time.perf_counter()
        a = oracle.calc_anom(x, doy, window=31, smooth=2)
        return time.perf_counter() - t0 + cpu_reference_run(a, lat, lon)[1]
    for _ in range(args.warmup):
        one()
    t = 0.0
    for _ in range(args.steps):
        t += one()
    v = T_sub * args.steps / t
    line = {'impl': 'reference', 'metric': 'timesteps/sec (721x1440 grid) run_contrack', 'value': v,
            'unit': 'timesteps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * t / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32 compare / f64 areas / int32 labels', 'data': 'synthetic',
            'config': workload_config(args.T, args.gpus, args.config),
            'cpu_baseline': {'value': v, 'unit': 'timesteps/s', 'cores': 1, 'kind': 'port',
                             'host_cores': os.cpu_count(),
                             'sample': '%d consecutive steps of the %dx%d cube as a standalone cube per step; the path '
                                       'is single-threaded (1 core used of %d)' % (T_sub, H, W, os.cpu_count())},
            'e2e': {'value': v, 'unit': 'timesteps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))
    return 0


def workload_config(T, n, config=0):
    pre = 'calc_anom(smooth=2, window=31, groupby=dayofyear) + ' if config == 3 else ''
    return {'workload': '%srun_contrack on synthetic %dx%dx%d Z500 %s (seed %d, sigma %s cells), threshold=%s %s '
                        'overlap=%.1f persistence=%d twosided=%s' % (pre, T, H, W, 'height (anomaly + seasonal cycle)'
                                                                     if config == 3 else 'anomaly', SEED, SIGMA,
                                                                     THRESHOLD, GORL, OVERLAP, PERSISTENCE, TWOSIDED),
            'baseline_config': {0: 'configs[2]/[3] (10957 steps)', 2: 'configs[1]', 3: 'configs[2]', 5: 'configs[4]'}[config],
            'T': T, 'H': H, 'W': W, 'sharding': 'time x%d' % n, 'threshold': THRESHOLD, 'overlap': OVERLAP,
            'persistence': PERSISTENCE,
            'l2': 'inputs (%.1f GB) and outputs far exceed the 126 MB L2; no explicit flush' % (T * H * W * 4 / 1e9)}


def day_groups(T):
    """(group index per time step, number of groups) of a daily axis that starts on 1981-01-01 (README.rst:112-117)."""
    from contrack_b200.contrack import time_group_keys
    times = (np.datetime64('1981-01-01') + np.arange(T).astype('timedelta64[D]')).astype('datetime64[ns]')
    doy = time_group_keys(times, 'dayofyear')
    uniq, gidx = np.unique(doy, return_inverse=True)
    return doy, gidx.astype(np.int32), len(uniq)


def parity_block(eng, world, rank, w, lat, lon, run_shard):
    """Bit-exactness against the oracle on a standalone cube that is cut at EVERY rank boundary: Tp = max(128, 32 x ranks)
    planes of the benchmark field, sharded over all ranks exactly like the timed cube (same entry points, same
    collectives), gathered on rank 0 and compared with the oracle's flag cube (contrack.py:646-772)."""
    import torch
    import torch.distributed as dist
    Tp = max(128, 32 * world)
    Tp -= Tp % world
    per = Tp // world
    x = torch.empty((per, H, W), dtype=torch.float32, device='cuda')
    synth_fill(x, rank * per, Tp)
    f, n = run_shard(x, rank * per, Tp)
    torch.cuda.synchronize()
    if world > 1:
        xs = torch.empty((Tp, H, W), dtype=torch.float32, device='cuda') if rank == 0 else None
        fs = torch.empty((Tp, H, W), dtype=torch.int32, device='cuda') if rank == 0 else None
        dist.gather(x, list(xs.split(per)) if rank == 0 else None, dst=0)
        dist.gather(f.contiguous(), list(fs.split(per)) if rank == 0 else None, dst=0)
    else:
        xs, fs = x, f
    if rank != 0:
        return None
    t0 = time.perf_counter()
    ref = oracle_run(xs.cpu().numpy(), lat, lon)
    dt = time.perf_counter() - t0
    got = fs.cpu().numpy()
    return {'bit_exact_vs_oracle': bool(np.array_equal(got, ref)), 'features': int(n),
            'features_oracle': int(len(np.unique(ref)) - 1), 'planes': Tp, 'cuts': world - 1, 'oracle_seconds': dt,
            'sample': 'standalone %d-step cube of the benchmark field, %d planes per rank, through the same entry points '
                      'as the timed steps; oracle = CPU restatement of contrack.py:646-772 on rank 0' % (Tp, per)}


def oracle_run(x, lat, lon):
    from oracle import contrack_oracle as oracle
    return oracle.run_contrack(x, lat, lon, THRESHOLD, GORL, OVERLAP, PERSISTENCE, TWOSIDED, force=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--T', type=int, default=0, help='time steps of the cube (default 10957: BASELINE.json configs[2..3])')
    ap.add_argument('--config', type=int, default=0, choices=[0, 2, 3, 5],
                    help='BASELINE.json configs[i-1]: 2 = 2707 steps; 3 = calc_anom(smooth=2, window=31) + run_contrack at '
                         '10957 steps (1 GPU); 5 = 43828 steps, overlap 0.7, persistence 20, threshold = 90th percentile of '
                         'the 80N-50N band over ALL time steps (needs >= 4 GPUs)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-T', type=int, default=0, help='time steps of the CPU-baseline sample')
    ap.add_argument('--e2e-T', type=int, default=0, help='time steps of the end-to-end (host buffer) measurement')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--tma', type=int, default=-1)
    ap.add_argument('--opt', action='append', default=[], help='engine option key=value (ct_set_option), repeatable')
    args = ap.parse_args()
    global THRESHOLD, OVERLAP, PERSISTENCE
    if not args.T:
        args.T = {0: 10957, 2: 2707, 3: 10957, 5: 43828}[args.config]
    if args.config == 5:
        OVERLAP, PERSISTENCE = 0.7, 20
    if args.impl == 'reference':
        if not args.cpu_T:
            args.cpu_T = 64
        return run_reference_arm(args)
    if not args.cpu_T:
        args.cpu_T = 256

    import torch
    import torch.distributed as dist
    from contrack_b200 import Engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the product path has no CPU fallback)')
    if args.config == 3 and world > 1:
        raise SystemExit('--config 3 is the single-GPU configuration of BASELINE.json (configs[2])')
    torch.cuda.set_device(local)
    real_stdout = None
    if world > 1:
        # NCCL (under NCCL_DEBUG=INFO without NCCL_DEBUG_FILE) and c10d log to stdout; the driver wants exactly one JSON
        # line there: stdout points at stderr while the job runs, the line goes to the saved descriptor
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        opts = None
        try:                                   # NCCL kernels must not queue behind the cube-sized zero-fill grid
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            pass
        dist.init_process_group('nccl', device_id=torch.device('cuda', local), pg_options=opts)
    from contrack_b200 import sharded

    T = args.T
    lat, lon = grid()
    w = reference_weights(lat, lon)
    eng = Engine.get(local)
    if args.tma >= 0:
        eng.set_option('tma', args.tma)
    for kv in args.opt:
        k, v = kv.split('=')
        eng.set_option(k, int(v))
    t_lo, t_hi = sharded.shard_bounds(T, world)[rank]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    anom = torch.empty((t_hi - t_lo, H, W), dtype=torch.float32, device='cuda')
    t_gen = time.perf_counter()
    synth_fill(anom, t_lo, T, season=(args.config == 3))
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    flag = torch.empty((t_hi - t_lo, H, W), dtype=torch.int32, device='cuda')
    thr_note = None
    zcube = gidx = None
    if args.config == 3:
        # the generated cube is the geopotential HEIGHT z (anomaly field + seasonal cycle); the anomaly is computed per step
        zcube, anom = anom, torch.empty_like(anom)
        _, gidx, G = day_groups(T)
    if args.config == 5:
        # README.rst:150-151: anom.sel(latitude=slice(80, 50)).quantile([0.9], dim='time').mean(), once before the timed
        # region.  The cube is time-sharded: the exact order statistics over ALL time steps come from the distributed
        # radix select (per-column histograms all-reduced over the ranks), bit-identical to np.nanquantile on the
        # gathered cube (tests/test_gpu_extras.py)
        y0, y1 = int(round((90 - 80) / 0.25)), int(round((90 - 50) / 0.25)) + 1
        q = sharded.quantile_time_sharded(eng, anom, [0.9], y0, y1) if world > 1 else eng.quantile_time(anom, [0.9], y0, y1)
        THRESHOLD = float(np.nanmean(q.cpu().numpy()))
        thr_note = ('float(np.nanmean(quantile(0.9, dim=time) of the 80N-50N band)) over all %d time steps (distributed exact '
                    'radix select)' % T)

    shard_info = []
    # the library's own NCCL communicator over the ranks of this job (unique id distributed through torch.distributed);
    # created before the timed region
    comm = sharded.default_comm(eng) if world > 1 else None

    def run_shard(x, t0, T_total, out=None):
        """one run_contrack pass over this rank's planes [t0, t0 + len(x)) of a T_total-step cube -> (flag, features)"""
        if world == 1:
            return eng.run_contrack(x, w, THRESHOLD, True, 0, OVERLAP, PERSISTENCE, TWOSIDED, out=out)
        f, n, st = sharded.run_contrack_sharded(eng, x, t0, T_total, w, THRESHOLD, True, 0, OVERLAP, PERSISTENCE,
                                                TWOSIDED, out=out, comm=comm)
        shard_info.append(st)
        return f, n

    ms_stage = {'calc_clim': [], 'calc_anom': []}
    ev_s = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def step():
        if args.config == 3:
            ev_s[0].record()
            clim = eng.calc_clim(zcube, gidx, G, 31)
            ev_s[1].record()
            eng.calc_anom(zcube, gidx, G, clim, 2, out=anom)
            ev_s[2].record()
        return run_shard(anom, t_lo, T, out=flag)

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms_thr, ms_paint, ms_host, ms_tab, ms_zero = [], [], [], [], []
    with ClockSampler(local) as clocks:
        for _ in range(args.warmup):
            step()
        barrier()
        t_timed0 = time.time()
        ev0.record()
        for _ in range(args.steps):
            _, nfeat = step()
            st = lambda k: max(0.0, eng.lib.ct_get_stat(eng.handle, k))          # noqa: E731  (five cheap look-ups per step)
            ms_thr.append(st(b'ms_threshold')); ms_paint.append(st(b'ms_paint')); ms_host.append(st(b'ms_host_tables'))
            ms_tab.append(st(b'ms_tables_gpu') + st(b'ms_tables_host_roundtrip'))
            ms_zero.append(st(b'ms_zero_fill'))
            if args.config == 3:                 # run_contrack ends with a stream synchronize: the events are complete
                ms_stage['calc_clim'].append(ev_s[0].elapsed_time(ev_s[1]))
                ms_stage['calc_anom'].append(ev_s[1].elapsed_time(ev_s[2]))
        ev1.record()
        barrier()
        t_timed1 = time.time()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t_ms = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        ms = float(t_ms.item())
    stats = eng.stats()
    value = T * args.steps / (ms / 1e3)
    peak, peak_kind = measured_peak()
    cells = (t_hi - t_lo) * H * W                      # cells one launch of the cube-sized kernels processes (this rank)
    thr_ms, paint_ms = float(np.mean(ms_thr)), float(np.mean(ms_paint))
    bytes_per_cell = 4
    dom = ('threshold_bits', thr_ms) if thr_ms >= paint_ms else ('paint', paint_ms)
    if args.config == 3 and float(np.mean(ms_stage['calc_anom'])) > dom[1]:
        dom, bytes_per_cell = ('anom_chunks', float(np.mean(ms_stage['calc_anom']))), 8
    achieved = cells * bytes_per_cell / (dom[1] / 1e3) / 1e9
    path_bytes = 20 if args.config == 3 else 8
    path_gbs = T * H * W * path_bytes / (ms / args.steps / 1e3) / 1e9
    traffic, traffic_src = ncu_traffic(dom[0], cells)
    # the same kernel timed alone (no table kernels beside it): two extra, untimed-for-`value` steps without the pipeline
    iso = None
    if world == 1 and dom[0] == 'threshold_bits':
        eng.set_option('chunks', 1)
        t_iso = []
        for _ in range(2):
            run_shard(anom, t_lo, T, out=flag)
            t_iso.append(eng.stats()['ms_threshold'])
        eng.set_option('chunks', 4)
        iso = cells * 4 / (min(t_iso) / 1e3) / 1e9
    line = {'metric': 'timesteps/sec (721x1440 grid) run_contrack', 'value': value, 'unit': 'timesteps/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32 compare / f64 areas / int32 labels', 'data': 'synthetic',
            'config': workload_config(T, world, args.config),
            'features': int(nfeat), 'gpu_launches': int(stats['kernel_launches']) * args.steps * world,
            'clocks': clocks.summary(t_timed0, t_timed1),
            'roofline': {'bound': 'hbm', 'kernel': dom[0], 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'peak_kind': peak_kind, 'traffic': traffic,
                         'traffic_note': 'DRAM read+write bytes per cell of this kernel in %s x %d cells' % (traffic_src, cells),
                         'algorithmic_bytes_per_launch': cells * bytes_per_cell,
                         'note': '%d B/cell x %d cells (rank 0) / CUDA-event time of that kernel inside the timed steps '
                                 '(threshold_bits: float32 read, the table kernels of the previous time chunk run beside it, '
                                 '%d chunk launches, first launch start -> last launch end; paint: int32 write; anom_chunks: '
                                 'float32 read + write)' % (bytes_per_cell, cells, int(stats.get('chunks', 1))),
                         'achieved_alone': iso, 'frac_alone': iso / peak if iso else None,
                         'peak_note': 'peak = copy bandwidth of MEASURED_PEAKS.json (a read+write mix); a pure read stream such '
                                      'as the threshold kernel alone can exceed it',
                         'note_alone': 'same kernel as one launch with nothing beside it (option chunks=1), CUDA events'},
            'roofline_path': {'achieved': path_gbs, 'frac': path_gbs / (peak * world), 'unit': 'GB/s',
                              'bytes_per_cell': path_bytes,
                              'note': ('%d B/cell x all cells / whole step time (run_contrack: read anomaly once + write flag '
                                       'once = 8; config 3 adds calc_clim 4 + calc_anom 8); frac against %d x the per-GPU peak'
                                       % (path_bytes, world))},
            'breakdown_ms': {'threshold_bits': thr_ms, 'paint': paint_ms, 'tables_gpu_and_host': float(np.mean(ms_tab)),
                             'host_table_phase': float(np.mean(ms_host)),
                             'zero_fill_overlapped_with_tables': float(np.mean(ms_zero)),
                             'plane_kernel': stats.get('ms_plane_kernel'), 'global_kernel': stats.get('ms_global_kernel')},
            'tables': {k: int(stats[k]) for k in ('runs', 'comps2d', 'pairs', 'seam_rows', 'kept_comps', 'labels3d',
                                                  'seam_events', 'seam_splits', 'neartie_resolved') if k in stats},
            'synth_seconds': t_gen}
    if args.config == 3:
        clim_ms, anom_ms = float(np.mean(ms_stage['calc_clim'])), float(np.mean(ms_stage['calc_anom']))
        rc_ms = ms / args.steps - clim_ms - anom_ms
        line['stages'] = {
            'calc_clim': {'ms': clim_ms, 'bytes_per_cell': 4, 'gbs': cells * 4 / clim_ms / 1e6, 'frac': cells * 4 / clim_ms / 1e6 / peak},
            'calc_anom': {'ms': anom_ms, 'bytes_per_cell': 8, 'gbs': cells * 8 / anom_ms / 1e6, 'frac': cells * 8 / anom_ms / 1e6 / peak},
            'run_contrack': {'ms': rc_ms, 'bytes_per_cell': 8, 'gbs': cells * 8 / rc_ms / 1e6, 'frac': cells * 8 / rc_ms / 1e6 / peak}}
    if thr_note:
        line['config']['threshold_value'] = THRESHOLD
        line['config']['threshold_note'] = thr_note
    if world > 1 and shard_info:
        last = shard_info[-args.steps:]
        keys = ('ms_threshold', 'ms_zero_fill', 'ms_tables_after_threshold', 'ms_plane_kernel', 'ms_exchange', 'ms_global_kernel', 'ms_host_tables',
                'ms_paint', 'ms_total')
        mine = {k: round(float(np.mean([i.get(k, 0.0) for i in last])), 3) for k in keys}
        mine.update({k: last[-1].get(k) for k in ('exchange_bytes', 'shard_attempts', 'fast_path', 'sweeps', 'kernel_launches', 'p2p')})
        every = [None] * world
        dist.all_gather_object(every, mine)
        line['shard_ms_all_ranks'] = every
        line['shard_ms_note'] = ('per rank, CUDA events of the sharded call: threshold (own planes), zero fill (beside the table '
                                 'phase), tables_after_threshold = plane kernel + table exchange (p2p 1: peer-memory stores by the pack kernel, 0: all-gather) + merge + global kernel + '
                                 'host replay, paint; ms_total = the whole call on the device')
    # ---- parity: checksum of the timed run's flag cube (identical at every N) + oracle on a cube cut at every rank ----
    if not args.no_parity:
        cs = flag_checksum(flag, t_lo * H * W)
        if world > 1:
            dist.all_reduce(cs)                                  # int64 sum wraps modulo 2^64
        line['parity'] = {'checksum': '%016x' % (int(cs.item()) & 0xffffffffffffffff),
                          'checksum_note': 'sum over all cells of flag * (mix64(t*H*W + y*W + x) | 1) mod 2^64 of the flag cube of '
                                           'the last timed step, summed over the ranks: independent of the sharding'}
        shard_info_keep = list(shard_info)
        pb = parity_block(eng, world, rank, w, lat, lon, run_shard)
        del shard_info[:]
        shard_info.extend(shard_info_keep)
        if pb:
            line['parity'].update(pb)

    # ---- CPU baseline + parity on a bounded sample (rank 0 of a single-GPU run only) ---------------------------------
    if not args.no_cpu and world == 1:
        from oracle import contrack_oracle as oracle
        Ts = min(args.cpu_T, T)
        if args.config == 3:
            doy, _, _ = day_groups(Ts)
            zs = zcube[:Ts].contiguous()
            zh = zs.cpu().numpy()
            t0 = time.perf_counter()
            xa = oracle.calc_anom(zh, doy, window=31, smooth=2)
            dt_anom = time.perf_counter() - t0
            _, gi, Gs = day_groups(Ts)
            ga = eng.calc_anom(zs, gi, Gs, eng.calc_clim(zs, gi, Gs, 31), 2)
            anom_close = bool(np.allclose(ga.cpu().numpy(), xa, rtol=1e-5, atol=4e-3, equal_nan=True))
            sub = ga
            x = sub.cpu().numpy()                                # the oracle tracks the GPU anomaly: flag parity is bit-exact
        else:
            sub = anom[:Ts].contiguous()
            x = sub.cpu().numpy()
            dt_anom = 0.0
        cpu_reference_run(np.ascontiguousarray(x[:min(8, Ts)]), lat, lon)          # warm-up (imports, allocator)
        ref, dt = cpu_reference_run(x, lat, lon)
        got, _ = eng.run_contrack(sub, w, THRESHOLD, True, 0, OVERLAP, PERSISTENCE, TWOSIDED)
        line['cpu_baseline'] = {'value': Ts / (dt + dt_anom), 'unit': 'timesteps/s', 'cores': 1, 'kind': 'port',
                                'host_cores': os.cpu_count(), 'seconds': dt + dt_anom,
                                'sample': 'first %d steps of the cube as a standalone cube (the reference holds int64 '
                                          'cubes and cannot run all %d steps on the host); single-threaded path: 1 core '
                                          'used of %d' % (Ts, T, os.cpu_count()),
                                'bit_exact_vs_gpu': bool(np.array_equal(got.cpu().numpy(), ref))}
        if args.config == 3:
            line['cpu_baseline']['calc_anom_seconds'] = dt_anom
            line['cpu_baseline']['calc_anom_within_tolerance'] = anom_close
            line['cpu_baseline']['tolerance'] = 'rtol 1e-5, atol 4e-3 (float32 values of magnitude 5500)'
        del sub, got
        # the reference's step 3 walks every label slot of the CUBE for every plane (O(T^2)): a shorter sample is faster
        # per step, the full cube would be slower than either
        Ts2 = min(64, T)
        _, dt2 = cpu_reference_run(np.ascontiguousarray(x[:Ts2]), lat, lon)
        line['cpu_baseline']['shorter_sample'] = {'steps': Ts2, 'value': Ts2 / dt2}

    # ---- end to end: host buffers in, host buffers out, copies inside the timed region --------------------------------
    if not args.no_e2e:
        del flag
        avail = 0
        try:
            for ln in open('/proc/meminfo'):
                if ln.startswith('MemAvailable'):
                    avail = int(ln.split()[1]) * 1024
        except Exception:
            pass
        Te = args.e2e_T or min(T, 2707)
        per_plane = H * W * 8
        if avail:
            Te = max(16 * world, min(Te, int(0.45 * avail // per_plane)))
        n_e2e = max(1, min(args.steps, 3))
        # the SAME standalone Te-step cube at every N: rank r holds planes shard_bounds(Te, N)[r] of it in pinned host memory
        e_lo, e_hi = sharded.shard_bounds(Te, world)[rank]
        dev_in = (zcube if args.config == 3 else anom)[:e_hi - e_lo]
        synth_fill(dev_in, e_lo, Te, season=(args.config == 3))
        xin = torch.empty((e_hi - e_lo, H, W), dtype=torch.float32, pin_memory=True)
        xin.copy_(dev_in)
        fout = torch.empty((e_hi - e_lo, H, W), dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()
        if args.config == 3:
            line['e2e'] = e2e_config3(xin.numpy(), Te, n_e2e, local)
        elif world == 1:
            xin_np, fout_np = xin.numpy(), fout.numpy()
            eng.run_contrack(xin_np, w, THRESHOLD, True, 0, OVERLAP, PERSISTENCE, TWOSIDED, out=fout_np)      # warm-up
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                _, nf = eng.run_contrack(xin_np, w, THRESHOLD, True, 0, OVERLAP, PERSISTENCE, TWOSIDED, out=fout_np)
            dt = time.perf_counter() - t0
            s = eng.stats()
            extra = {'ms_h2d_threshold': s.get('ms_h2d_threshold'), 'ms_tables': s.get('ms_tables'),
                     'ms_paint_d2h': s.get('ms_paint_d2h'), 'host_threads': int(s.get('host_threads', 0)),
                     'note': 'ct_run_contrack_host: pinned host float32 cube in (chunked copies overlapped with the '
                             'threshold kernel), int32 flag cube out on the host; the result crosses PCIe as the row-run '
                             'table (12 B per run) and host threads expand it into the zero-filled host cube; wall clock '
                             'around the call'}
            d2h = int(s.get('d2h_bytes', Te * H * W * 4))
            extra['checksum'] = '%016x' % (int(flag_checksum(torch.from_numpy(fout_np).cuda(), 0).item()) & 0xffffffffffffffff)
        else:
            # every rank holds its shard of the Te-step cube in pinned host memory and calls the host-buffer entry point: the
            # shard streams through the GPU under the threshold kernel, the flag planes come back as a row-run table that host
            # threads expand (ct_run_contrack_sharded_host)
            xin_np, fout_np = xin.numpy(), fout.numpy()

            def e2e_step():
                _, n, _ = sharded.run_contrack_sharded(eng, xin_np, e_lo, Te, w, THRESHOLD, True, 0, OVERLAP, PERSISTENCE,
                                                       TWOSIDED, out=fout_np, comm=comm)
                return n
            barrier()                                    # (rank 0 may come late: it ran the CPU oracle of the parity block)
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                nf = e2e_step()
            barrier()
            dt = time.perf_counter() - t0
            t_dt = torch.tensor([dt], dtype=torch.float64, device='cuda')
            dist.all_reduce(t_dt, op=dist.ReduceOp.MAX)
            dt = float(t_dt.item())
            s = eng.stats()
            d2h_t = torch.tensor([float(s.get('d2h_bytes', 0.0))], dtype=torch.float64, device='cuda')
            dist.all_reduce(d2h_t)
            # parity of the host path: checksum of the flag planes that arrived on the hosts
            cs_e = flag_checksum(torch.from_numpy(fout_np).cuda(), e_lo * H * W)
            dist.all_reduce(cs_e)
            extra = {'host_threads': int(s.get('host_threads', 0)), 'checksum': '%016x' % (int(cs_e.item()) & 0xffffffffffffffff),
                     'note': 'ct_run_contrack_sharded_host per rank: pinned host shard in (chunked copies under the threshold '
                             'kernel), flag planes out on the host as a row-run table expanded by host threads; wall clock, max '
                             'over ranks; host memory bandwidth (reading the shards for PCIe, zeroing the results) is shared by '
                             'all ranks of the box'}
            d2h = int(d2h_t.item())
        if args.config != 3:
            line['e2e'] = dict({'value': Te * n_e2e / dt, 'unit': 'timesteps/s', 'h2d_bytes_per_step': Te * H * W * 4,
                                'd2h_bytes_per_step': d2h, 'T': Te, 'steps': n_e2e, 'features': int(nf),
                                'cube': 'standalone %d-step cube of the benchmark field, the same at every N' % Te}, **extra)
    if rank == 0:
        if real_stdout is not None:
            sys.stdout.flush()
            os.write(real_stdout, (json.dumps(line) + '\n').encode())
        else:
            print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def e2e_config3(z_host, Te, n, device):
    """Config 3 end to end through the class API (the call a user of the reference makes): host z cube in a Dataset ->
    calc_anom(smooth=2, window=31) -> run_contrack -> ds['flag'] on the host."""
    from contrack_b200 import contrack
    from contrack_b200 import dataset as ds_mod
    lat, lon = grid()
    times = (np.datetime64('1981-01-01') + np.arange(Te).astype('timedelta64[D]')).astype('datetime64[ns]')

    def once():
        ds = ds_mod.Dataset({'z': ds_mod.Variable(('time', 'latitude', 'longitude'), z_host,
                                                  {'units': 'm', 'long_name': 'Geopotential Height'})},
                            coords={'time': times, 'latitude': lat, 'longitude': lon})
        c = contrack(ds=ds, device=device)
        c.set_up(force=True, write=False)
        c.calc_anom('z', window=31, smooth=2)
        c.run_contrack('anom', THRESHOLD, GORL, OVERLAP, PERSISTENCE, TWOSIDED)
        return c
    once()
    t0 = time.perf_counter()
    for _ in range(n):
        c = once()
    dt = time.perf_counter() - t0
    f = np.asarray(c['flag'].data)
    cells = Te * H * W
    return {'value': Te * n / dt, 'unit': 'timesteps/s', 'h2d_bytes_per_step': cells * 8, 'd2h_bytes_per_step': cells * 4 + 0,
            'T': Te, 'steps': n, 'features': int(len(np.unique(f)) - 1),
            'note': 'contrack(ds).calc_anom(z, window=31, smooth=2) + run_contrack(anom) on a host Dataset: z crosses PCIe for '
                    'calc_anom, the float32 anomaly comes back (4 B/cell) and is streamed in again by run_contrack, the flag '
                    'returns as a row-run table; wall clock around the three calls'}


if __name__ == '__main__':
    sys.exit(main())
