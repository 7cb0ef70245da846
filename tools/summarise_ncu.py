#!/usr/bin/env python3
"""Summaries of the ncu outputs that tools/profile_step.sh leaves in gpurun_out/ -> profiles/ (run on the build box).

  python tools/summarise_ncu.py <tag>      e.g. r1b  ->  profiles/<tag>_ncu_full_T512.csv, profiles/<tag>_launches_T10957.csv,
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
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
           'launch__block_size', 'smsp__inst_executed.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
           'lts__t_sectors_srcunit_tex_op_write.sum']


def to_bytes(v, unit):
    return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
    T_prof, H, W = 512, 721, 1440
    raw = subprocess.run(['ncu', '-i', os.path.join(OUT, 'prof_T%d.ncu-rep' % T_prof), '--page', 'raw', '--csv'],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    idx = [head.index(m) for m in METRICS]
    with open(os.path.join(PROF, '%s_ncu_full_T%d.csv' % (tag, T_prof)), 'w', newline='') as f:
        wr = csv.writer(f)
        wr.writerow(METRICS)
        wr.writerow([units[i] for i in idx])
        for r in rows[2:]:
            wr.writerow([r[i] for i in idx])
    cells = T_prof * H * W
    traffic = {}
    ir, iw, ik = head.index('dram__bytes_read.sum'), head.index('dram__bytes_write.sum'), head.index('Kernel Name')
    it = head.index('gpu__time_duration.sum')
    for r in rows[2:]:
        name = 'threshold_bits' if 'k_threshold' in r[ik] else 'zero_fill' if 'k_zero_fill' in r[ik] else 'paint'
        rd, wrb = to_bytes(r[ir], units[ir]), to_bytes(r[iw], units[iw])
        traffic[name] = {'kernel': r[ik].split('(')[0].replace('void ', ''), 'cells': cells, 'dram_read_bytes': rd,
                         'dram_write_bytes': wrb, 'bytes_per_cell': (rd + wrb) / cells,
                         'duration_us_under_ncu': float(r[it])}
    with open(os.path.join(PROF, 'ncu_traffic.json'), 'w') as f:
        json.dump({'source': '%s_ncu_full_T%d.csv (ncu --set full, T=%d x %d x %d)' % (tag, T_prof, T_prof, H, W),
                   'kernels': traffic}, f, indent=1)
    # launch list: keep the csv, add a per-kernel share table of the LAST step (the first is the warm-up)
    src = os.path.join(OUT, 'launches_T10957.csv')
    lines = [ln for ln in open(src) if not ln.startswith('==')]
    with open(os.path.join(PROF, '%s_launches_T10957.csv' % tag), 'w') as f:
        f.writelines(lines)
    rows = list(csv.reader(lines))
    h = rows[0]
    kn, mv = h.index('Kernel Name'), h.index('Metric Value')
    body = [r for r in rows[1:] if len(r) > mv]
    thr = [i for i, r in enumerate(body) if 'k_threshold' in r[kn]]
    step = body[thr[-1]:] if thr else body
    agg = collections.OrderedDict()
    for r in step:
        k = r[kn].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
        a = agg.setdefault(k, [0, 0.0])
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
    print(json.dumps(traffic, indent=1))


if __name__ == '__main__':
    main()
