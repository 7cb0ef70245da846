#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fastpath.py tests/test_gpu_sharded.py -x -q > gpurun_out/r2l_tests.log 2>&1; rc=$?; echo "tests rc=$rc"
tail -4 gpurun_out/r2l_tests.log
[ $rc -eq 124 ] && exit 1
run() { tag=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity "$@" > gpurun_out/r2l_$tag.json 2> gpurun_out/r2l_$tag.err; echo "$tag rc=$?"; }
run T1370_late1 --T 1370 --opt plane_kernel=1 --opt fill_late=1
run T1370_late0 --T 1370 --opt plane_kernel=1 --opt fill_late=0
run T1370_chunks4 --T 1370 --opt plane_kernel=1 --opt fill_late=1 --opt fast_chunks=4 --opt chunk_min_planes=256
run T2739_late1 --T 2739 --opt plane_kernel=1 --opt fill_late=1
run T2739_chunks4 --T 2739 --opt plane_kernel=1 --opt fill_late=1 --opt fast_chunks=4 --opt chunk_min_planes=256
run T5479_late1 --T 5479 --opt plane_kernel=1 --opt fill_late=1
run T5479_classic --T 5479 --opt plane_kernel=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l_T*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        b=d['breakdown_ms']
        print('%-40s %8d ts/s %7.3f ms  thr %.2f fill %.2f after_thr %.2f plane %s global %s paint %.2f' % (f.split('/')[-1], d['value'], d['ms_per_step'], b['threshold_bits'], b['zero_fill_overlapped_with_tables'], b['tables_gpu_and_host'], b.get('plane_kernel'), b.get('global_kernel'), b['paint']))
    except Exception as e: print(f,'ERR',e)
PY
