#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity "$@" > gpurun_out/r2n_$tag.json 2> gpurun_out/r2n_$tag.err; echo "$tag rc=$?"; }
run T10957_ch1_c2 --opt chunks=1 --opt fill_ctas=2
run T10957_ch1_c1 --opt chunks=1 --opt fill_ctas=1
run T10957_ch4_c1 --opt chunks=4 --opt fill_ctas=1
run T10957_ch2_c2 --opt chunks=2 --opt fill_ctas=2
run T10957_ch1_c3 --opt chunks=1 --opt fill_ctas=3
run T1370_c1_late0 --T 1370 --opt fill_ctas=1 --opt fill_late=0
run T1370_c3_late0 --T 1370 --opt fill_ctas=3 --opt fill_late=0
run T2739_c2_late0 --T 2739 --opt fill_ctas=2 --opt fill_late=0
run T2739_c1_late0 --T 2739 --opt fill_ctas=1 --opt fill_late=0
run T5479_ch1_c2 --T 5479 --opt chunks=1 --opt fill_ctas=2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2n_T*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        b=d['breakdown_ms']
        print('%-32s %8d ts/s %7.3f ms  thr %.2f fill %.2f after_thr %.2f plane %s global %s paint %.2f' % (f.split('/')[-1], d['value'], d['ms_per_step'], b['threshold_bits'], b['zero_fill_overlapped_with_tables'], b['tables_gpu_and_host'], b.get('plane_kernel'), b.get('global_kernel'), b['paint']))
    except Exception as e: print(f,'ERR',e)
PY
