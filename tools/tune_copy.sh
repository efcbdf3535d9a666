#!/bin/bash
# copy-kernel variants on the C3 bench (development aid): prints roofline of each
for cfg in "3 6 32" "3 2 32" "3 1 32" "3 3 12" "3 2 6" "3 1 3" "4 6 32" "4 2 12" "5 6 32" "5 2 8" "1 6 32" "1 2 8" "6 6 32" "0 6 32"; do
  set -- $cfg
  SLAMRS_COPY_VARIANT=$1 SLAMRS_COPY_K=$2 SLAMRS_COPY_GRID=$3 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-full-copy --steps 16 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('variant/k/grid $cfg', 'ms/step %.4f' % d['ms_per_step'], 'copy ms %.4f' % r['ms_per_launch'], 'GB/s %.0f' % r['achieved'], 'frac %.3f' % r['frac'])
"
done
