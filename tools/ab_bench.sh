#!/bin/bash
# A/B of an experiment switch on the same box: tools/ab_bench.sh VAR valueA valueB  -> per-class ms/step of both
for v in "$2" "$3" "$2" "$3"; do
  env $1=$v timeout 90 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$1=$v', round(d['ms_per_step'],2), {k:round(x,2) for k,x in d['roofline']['ms_per_step_by_kernel_class'].items()})"
done
