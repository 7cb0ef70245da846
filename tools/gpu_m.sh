#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity "$@" > gpurun_out/r2m_$tag.json 2> gpurun_out/r2m_$tag.err; echo "$tag rc=$?"; }
run T1370_c4_late1 --T 1370 --opt fill_ctas=4 --opt fill_late=1
run T1370_c4_late0 --T 1370 --opt fill_ctas=4 --opt fill_late=0
run T1370_c2_late0 --T 1370 --opt fill_ctas=2 --opt fill_late=0
run T1370_c2_late1 --T 1370 --opt fill_ctas=2 --opt fill_late=1
run T2739_c4_late0 --T 2739 --opt fill_ctas=4 --opt fill_late=0
run T2739_c4_late1 --T 2739 --opt fill_ctas=4 --opt fill_late=1
run T10957_c4 --opt fill_ctas=4
run T10957_c2 --opt fill_ctas=2
run T10957_c8 --opt fill_ctas=8
run T10957_plane_c4 --opt fill_ctas=4 --opt plane_kernel=1 --opt fill_late=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_T*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        b=d['breakdown_ms']
        print('%-32s %8d ts/s %7.3f ms  thr %.2f fill %.2f after_thr %.2f plane %s global %s paint %.2f' % (f.split('/')[-1], d['value'], d['ms_per_step'], b['threshold_bits'], b['zero_fill_overlapped_with_tables'], b['tables_gpu_and_host'], b.get('plane_kernel'), b.get('global_kernel'), b['paint']))
    except Exception as e: print(f,'ERR',e)
PY
