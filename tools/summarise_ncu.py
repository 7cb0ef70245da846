#!/usr/bin/env python3
"""Summaries of the ncu outputs that tools/profile_step.sh leaves in gpurun_out/ -> profiles/ (run on the build box).

  python tools/summarise_ncu.py <tag>   e.g. r1c -> profiles/<tag>_ncu_full.csv, profiles/<tag>_launches_T10957.csv,
                                                    profiles/<tag>_launch_shares_T10957.txt, profiles/ncu_traffic.json
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')
METRICS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
           'lts__t_sectors_srcunit_tex_op_write.sum']
T_PROF, H, W = 1461, 721, 1440        # tools/profile_step.sh


def to_bytes(v, unit):
    return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def short(name):
    return name.split('(')[0].replace('void ', '').replace('<unnamed>::', '')


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1c'
    raw = subprocess.run(['ncu', '-i', os.path.join(OUT, '%s_full.ncu-rep' % tag), '--page', 'raw', '--csv'],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    metrics = [m for m in METRICS if m in head]
    idx = [head.index(m) for m in metrics]
    # the target runs every kernel twice (warm-up pass + second pass): keep the second instance of each kernel
    last = collections.OrderedDict()
    for r in rows[2:]:
        last[short(r[head.index('Kernel Name')])] = r
    with open(os.path.join(PROF, '%s_ncu_full.csv' % tag), 'w', newline='') as f:
        wr = csv.writer(f)
        wr.writerow(['# ncu --set full --clock-control none, tools/prof_target.py %d (cube %dx%dx%d, %.2f GB float32); second '
                     'instance of every kernel' % (T_PROF, T_PROF, H, W, T_PROF * H * W * 4 / 1e9)])
        wr.writerow(metrics)
        wr.writerow([units[i] for i in idx])
        for r in last.values():
            wr.writerow([short(r[i]) if i == idx[0] else r[i] for i in idx])
    cells = T_PROF * H * W
    traffic = {}
    ir, iw, it = head.index('dram__bytes_read.sum'), head.index('dram__bytes_write.sum'), head.index('gpu__time_duration.sum')
    names = {'k_threshold': 'threshold_bits', 'k_zero_fill': 'zero_fill', 'k_paint': 'paint', 'k_anom': 'calc_anom',
             'k_group_mean': 'group_mean', 'k_clim_smooth': 'clim_smooth', 'k_flag_count': 'flag_count',
             'k_compact_runs': 'compact_runs', 'k_plane_tables': 'plane_tables', 'k_global_phase': 'global_phase'}
    for k, r in last.items():
        name = next((v for p, v in names.items() if p in k), k)
        rd, wrb = to_bytes(r[ir], units[ir]), to_bytes(r[iw], units[iw])
        traffic[name] = {'kernel': k, 'cells': cells, 'dram_read_bytes': rd, 'dram_write_bytes': wrb,
                         'bytes_per_cell': (rd + wrb) / cells, 'duration_ms_under_ncu': float(r[it]),
                         'dram_gbs_under_ncu': (rd + wrb) / float(r[it]) / 1e6}
    with open(os.path.join(PROF, 'ncu_traffic.json'), 'w') as f:
        json.dump({'source': '%s_ncu_full.csv (ncu --set full, T=%d x %d x %d)' % (tag, T_PROF, H, W), 'kernels': traffic}, f,
                  indent=1)
    # launch list: keep the csv, add a per-kernel share table of the LAST step (the first is the warm-up)
    src = os.path.join(OUT, '%s_launches_T10957.csv' % tag)
    lines = [ln for ln in open(src) if not ln.startswith('==')]
    with open(os.path.join(PROF, '%s_launches_T10957.csv' % tag), 'w') as f:
        f.writelines(lines)
    rows = list(csv.reader(lines))
    h = rows[0]
    kn, mv = h.index('Kernel Name'), h.index('Metric Value')
    body = [r for r in rows[1:] if len(r) > mv]
    # bench.py: warm-up step, timed step, then two single-launch steps (chunks=1) for `achieved_alone`: the timed step is
    # the second group of threshold launches (4 chunk launches each)
    thr = [i for i, r in enumerate(body) if 'k_threshold' in r[kn]]
    step = body[thr[4]:thr[8]] if len(thr) >= 9 else body[thr[-1]:] if thr else body
    agg = collections.OrderedDict()
    for r in step:
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(',', ''))
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, '%s_launch_shares_T10957.txt' % tag), 'w') as f:
        f.write('one run_contrack step at T=10957 (ncu --metrics gpu__time_duration.sum --clock-control none; serialised, '
                'cold cache: compare shares, not absolutes)\n')
        f.write('%-48s %6s %12s %7s\n' % ('kernel', 'count', 'ms', 'share'))
        for k, a in agg.items():
            f.write('%-48s %6d %12.3f %6.1f%%\n' % (k[:48], a[0], a[1] / 1e6, 100 * a[1] / tot))
        f.write('%-48s %6d %12.3f\n' % ('total', sum(a[0] for a in agg.values()), tot / 1e6))
    print(open(os.path.join(PROF, '%s_launch_shares_T10957.txt' % tag)).read())
    src = os.path.join(OUT, '%s_launches_T1370.csv' % tag)
    if os.path.exists(src):                       # shard-sized cube: launches of the last step (from its threshold launch on)
        lines = [ln for ln in open(src) if not ln.startswith('==')]
        rows = list(csv.reader(lines))
        h = rows[0]
        kn, mv = h.index('Kernel Name'), h.index('Metric Value')
        body = [r for r in rows[1:] if len(r) > mv]
        thr = [i for i, r in enumerate(body) if 'k_threshold' in r[kn]]
        step = body[thr[1]:thr[2]] if len(thr) >= 3 else body
        with open(os.path.join(PROF, '%s_launch_list_T1370.txt' % tag), 'w') as f:
            f.write('one run_contrack step at T=1370 (the shard of one rank of an 8-GPU run), every launch in order; ncu '
                    '--metrics gpu__time_duration.sum --clock-control none: serialised, cold cache\n')
            for r in step:
                f.write('%-48s %10.1f us\n' % (short(r[kn])[:48], float(r[mv].replace(',', '')) / 1e3))
            f.write('%d launches, %.3f ms\n' % (len(step), sum(float(r[mv].replace(',', '')) for r in step) / 1e6))
        print(open(os.path.join(PROF, '%s_launch_list_T1370.txt' % tag)).read())
    print(json.dumps(traffic, indent=1))


if __name__ == '__main__':
    main()
