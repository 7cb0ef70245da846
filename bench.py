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

# NCCL prints its version banner on stdout under NCCL_DEBUG=VERSION/INFO; the driver wants ONE JSON line there
if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'INFO', ''):
    os.environ['NCCL_DEBUG'] = 'WARN'

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


def synth_fill(out, t0, T_total, seed=SEED, season=False):
    """Fill the CUDA float32 tensor out[nt, H, W] with planes [t0, t0+nt) of the synthetic cube (bench_support/)."""
    global _synth
    import torch
    if _synth is None:
        _synth = C.CDLL(os.path.join(ROOT, 'bench_support', 'libct_synth.so'))
        _synth.ct_synth_fill.restype = C.c_int
        _synth.ct_synth_fill.argtypes = [C.c_void_p, C.c_ulonglong, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int,
                                         C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_void_p]
    nt, h, w = out.shape
    rc = _synth.ct_synth_fill(C.c_void_p(out.data_ptr()), seed, t0, nt, T_total, h, w, SIGMA[0],
                              SIGMA[1] * h / 721.0, SIGMA[2] * w / 1440.0, 100.0, 60.0 if season else 0.0, 365.25,
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError('ct_synth_fill failed: cudaError %d' % rc)
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


def make_sample(T_sub, use_gpu):
    """[T_sub, H, W] float32 host array of the benchmark's synthetic field (first T_sub planes of a T_sub-long cube)."""
    if use_gpu:
        import torch
        d = torch.empty((T_sub, H, W), dtype=torch.float32, device='cuda')
        synth_fill(d, 0, T_sub)
        torch.cuda.synchronize()
        return d.cpu().numpy()
    from _synth import synth_cube
    return synth_cube(SEED, T_sub, H, W, SIGMA)


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
    x = make_sample(T_sub, use_gpu)
    for _ in range(args.warmup):
        cpu_reference_run(x, lat, lon)
    t = 0.0
    for _ in range(args.steps):
        _, dt = cpu_reference_run(x, lat, lon)
        t += dt
    v = T_sub * args.steps / t
    line = {'impl': 'reference', 'metric': 'timesteps/sec (721x1440 grid) run_contrack', 'value': v,
            'unit': 'timesteps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * t / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32 compare / f64 areas / int32 labels', 'data': 'synthetic',
            'config': workload_config(args.T, args.gpus),
            'cpu_baseline': {'value': v, 'unit': 'timesteps/s', 'cores': 1, 'kind': 'port',
                             'host_cores': os.cpu_count(),
                             'sample': '%d consecutive steps of the %dx%d cube as a standalone cube per step; the path '
                                       'is single-threaded (1 core used of %d)' % (T_sub, H, W, os.cpu_count())},
            'e2e': {'value': v, 'unit': 'timesteps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))
    return 0


def workload_config(T, n):
    return {'workload': 'run_contrack on synthetic %dx%dx%d Z500 anomaly (seed %d, sigma %s cells), threshold=%d %s '
                        'overlap=%.1f persistence=%d twosided=%s' % (T, H, W, SEED, SIGMA, THRESHOLD, GORL, OVERLAP,
                                                                     PERSISTENCE, TWOSIDED),
            'T': T, 'H': H, 'W': W, 'sharding': 'time x%d' % n, 'threshold': THRESHOLD, 'overlap': OVERLAP,
            'persistence': PERSISTENCE,
            'l2': 'inputs (%.1f GB) and outputs far exceed the 126 MB L2; no explicit flush' % (T * H * W * 4 / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--T', type=int, default=0, help='time steps of the cube (default 10957: BASELINE.json configs[2..3])')
    ap.add_argument('--config', type=int, default=0, choices=[0, 2, 5],
                    help='BASELINE.json configs[i-1]: 2 = 2707 steps; 5 = 43828 steps, overlap 0.7, persistence 20, threshold = '
                         '90th percentile of the 80N-50N band (needs >= 4 GPUs)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-T', type=int, default=0, help='time steps of the CPU-baseline sample')
    ap.add_argument('--e2e-T', type=int, default=0, help='time steps of the end-to-end (host buffer) measurement')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--tma', type=int, default=-1)
    ap.add_argument('--opt', action='append', default=[], help='engine option key=value (ct_set_option), repeatable')
    args = ap.parse_args()
    global THRESHOLD, OVERLAP, PERSISTENCE
    if not args.T:
        args.T = {0: 10957, 2: 2707, 5: 43828}[args.config]
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
    torch.cuda.set_device(local)
    real_stdout = None
    if world > 1:
        # NCCL / c10d print banners on stdout ("NCCL version ..."); the driver wants exactly one JSON line there
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
    synth_fill(anom, t_lo, T)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    flag = torch.empty((t_hi - t_lo, H, W), dtype=torch.int32, device='cuda')
    thr_note = None
    if args.config == 5:
        # README.rst:150-151: anom.sel(latitude=slice(80, 50)).quantile([0.9], dim='time').mean(), computed once before the
        # timed region.  The cube is time-sharded, so every rank takes the quantile over ITS time steps and the ranks' band
        # means are averaged: a harness-side stand-in for the global quantile (the field is stationary in time).
        y0, y1 = int(round((90 - 80) / 0.25)), int(round((90 - 50) / 0.25)) + 1
        q = eng.quantile_time(anom, [0.9], y0, y1)
        m = torch.nanmean(q).reshape(1)
        if world > 1:
            dist.all_reduce(m)
            m /= world
        THRESHOLD = float(m.item())
        thr_note = '90th percentile over time per grid point of the 80N-50N band, band mean, averaged over the time shards'

    shard_info = []

    def step():
        if world == 1:
            return eng.run_contrack(anom, w, THRESHOLD, True, 0, OVERLAP, PERSISTENCE, TWOSIDED, out=flag)
        f, n, info = sharded.run_contrack_sharded(eng, anom, t_lo, T, w, THRESHOLD, True, 0, OVERLAP, PERSISTENCE, TWOSIDED,
                                                  out=flag)
        shard_info.append(info)
        return f, n

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
    dom = ('threshold_bits', thr_ms) if thr_ms >= paint_ms else ('paint', paint_ms)
    achieved = cells * 4 / (dom[1] / 1e3) / 1e9
    path_gbs = T * H * W * 8 / (ms / args.steps / 1e3) / 1e9
    traffic, traffic_src = ncu_traffic(dom[0], cells)
    # the same kernel timed alone (no table kernels beside it): two extra, untimed-for-`value` steps without the pipeline
    iso = None
    if world == 1:
        eng.set_option('chunks', 1)
        t_iso = []
        for _ in range(2):
            step()
            t_iso.append(eng.stats()['ms_threshold'])
        eng.set_option('chunks', 4)
        iso = cells * 4 / (min(t_iso) / 1e3) / 1e9
    line = {'metric': 'timesteps/sec (721x1440 grid) run_contrack', 'value': value, 'unit': 'timesteps/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32 compare / f64 areas / int32 labels', 'data': 'synthetic',
            'config': workload_config(T, world),
            'features': int(nfeat), 'gpu_launches': int(stats['kernel_launches']) * args.steps * world,
            'clocks': clocks.summary(t_timed0, t_timed1),
            'roofline': {'bound': 'hbm', 'kernel': dom[0], 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'peak_kind': peak_kind, 'traffic': traffic,
                         'traffic_note': 'DRAM read+write bytes per cell of this kernel in %s x %d cells' % (traffic_src, cells),
                         'algorithmic_bytes_per_launch': cells * 4,
                         'note': '4 B/cell (float32 read for threshold_bits, int32 write for paint) x %d cells (rank 0) / '
                                 'CUDA-event time of that kernel inside the timed steps, where the table kernels of the '
                                 'previous time chunk run beside it (the cube is thresholded in %d chunk launches; '
                                 'the time is first launch start -> last launch end)' % (cells, int(stats.get('chunks', 1))),
                         'achieved_alone': iso, 'frac_alone': iso / peak if iso else None,
                         'peak_note': 'peak = copy bandwidth of MEASURED_PEAKS.json (a read+write mix); a pure read stream such '
                                      'as this kernel alone can exceed it',
                         'note_alone': 'same kernel as one launch with nothing beside it (option chunks=1), CUDA events'},
            'roofline_path': {'achieved': path_gbs, 'frac': path_gbs / (peak * world), 'unit': 'GB/s',
                              'note': '8 B/cell (read anomaly once + write flag once) x all cells / whole step time; '
                                      'frac against %d x the per-GPU peak' % world},
            'breakdown_ms': {'threshold_bits': thr_ms, 'paint': paint_ms, 'tables_gpu_and_host': float(np.mean(ms_tab)),
                             'host_table_phase': float(np.mean(ms_host)),
                             'zero_fill_overlapped_with_tables': float(np.mean(ms_zero))},
            'tables': {k: int(stats[k]) for k in ('runs', 'comps2d', 'pairs', 'seam_rows', 'kept_comps', 'labels3d',
                                                  'seam_events', 'seam_splits', 'neartie_resolved') if k in stats},
            'synth_seconds': t_gen}
    if thr_note:
        line['config']['threshold_value'] = THRESHOLD
        line['config']['threshold_note'] = thr_note
    if world > 1 and shard_info:
        last = shard_info[-args.steps:]
        line['shard_ms'] = {k: float(np.mean([i['phase_ms'][k] for i in last])) for k in last[0]['phase_ms']}
        line['shard_ms']['note'] = ('rank 0 host wall clock per phase of the sharded step (phases end where the host has to '
                                    'wait: halo exchange, table counts, gathered counts, global phase, paint)')
        line['shard_table_bytes'] = last[-1]['table_bytes']
        every = [None] * world
        dist.all_gather_object(every, {k: round(v, 3) for k, v in line['shard_ms'].items() if k != 'note'})
        line['shard_ms_all_ranks'] = every

    # ---- CPU baseline + parity on a bounded sample (rank 0 of a single-GPU run only) ---------------------------------
    if not args.no_cpu and world == 1:
        Ts = min(args.cpu_T, T)
        sub = anom[:Ts].contiguous()
        x = sub.cpu().numpy()
        cpu_reference_run(np.ascontiguousarray(x[:min(8, Ts)]), lat, lon)          # warm-up (imports, allocator)
        ref, dt = cpu_reference_run(x, lat, lon)
        got, _ = eng.run_contrack(sub, w, THRESHOLD, True, 0, OVERLAP, PERSISTENCE, TWOSIDED)
        line['cpu_baseline'] = {'value': Ts / dt, 'unit': 'timesteps/s', 'cores': 1, 'kind': 'port',
                                'host_cores': os.cpu_count(), 'seconds': dt,
                                'sample': 'first %d steps of the cube as a standalone cube (the reference holds int64 '
                                          'cubes and cannot run all %d steps on the host); single-threaded path: 1 core '
                                          'used of %d' % (Ts, T, os.cpu_count()),
                                'bit_exact_vs_gpu': bool(np.array_equal(got.cpu().numpy(), ref))}
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
        if world == 1:
            xin = torch.empty((Te, H, W), dtype=torch.float32, pin_memory=True)
            xin.copy_(anom[:Te])
            fout = torch.empty((Te, H, W), dtype=torch.int32, pin_memory=True)
            xin_np, fout_np = xin.numpy(), fout.numpy()
            torch.cuda.synchronize()
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
        else:
            # every rank holds its shard of a Te-step cube in pinned host memory; per step: H2D of the shard, the sharded
            # run, D2H of the shard's flag planes
            e_lo, e_hi = sharded.shard_bounds(Te, world)[rank]
            dev_in = anom[:e_hi - e_lo]
            synth_fill(dev_in, e_lo, Te)
            xin = torch.empty((e_hi - e_lo, H, W), dtype=torch.float32, pin_memory=True)
            xin.copy_(dev_in)
            fout = torch.empty((e_hi - e_lo, H, W), dtype=torch.int32, pin_memory=True)
            dev_out = torch.empty((e_hi - e_lo, H, W), dtype=torch.int32, device='cuda')

            def e2e_step():
                dev_in.copy_(xin, non_blocking=True)
                _, n, _ = sharded.run_contrack_sharded(eng, dev_in, e_lo, Te, w, THRESHOLD, True, 0, OVERLAP, PERSISTENCE,
                                                       TWOSIDED, out=dev_out)
                fout.copy_(dev_out, non_blocking=True)
                torch.cuda.synchronize()
                return n
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
            extra = {'note': 'per rank: pinned host shard -> device, sharded run, flag shard -> pinned host; wall clock, '
                             'max over ranks'}
            d2h = Te * H * W * 4
        line['e2e'] = dict({'value': Te * n_e2e / dt, 'unit': 'timesteps/s', 'h2d_bytes_per_step': Te * H * W * 4,
                            'd2h_bytes_per_step': d2h, 'T': Te, 'steps': n_e2e, 'features': int(nf)}, **extra)
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


if __name__ == '__main__':
    sys.exit(main())
