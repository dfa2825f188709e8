#!/bin/bash
# A/B of two builds of the library on the same box: tools/ab_lib.sh path/to/alt.so
for v in "" "$1" "" "$1"; do
  env PHOREGEN_B200_LIB=$v timeout 90 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('lib=${v:-default}', round(d['ms_per_step'],2), {k:round(x,2) for k,x in d['roofline']['ms_per_step_by_kernel_class'].items()})"
done
