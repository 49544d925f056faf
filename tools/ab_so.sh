#!/bin/bash
# A/B of whole libraries built under different names (MGW_SO_NAME): tools/ab_so.sh so1 so2 ...  -> ablate lines + bench step
for so in "$@"; do
  echo "== $so"
  MGW_SO_NAME=$so timeout 200 python tools/ablate.py 2>&1 | grep -E "^(fwd: out \+ black|bwd: dU \+ dHs, with|mesh fwd|mesh bwd|tile)"
  MGW_SO_NAME=$so timeout 300 python bench.py --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('step us %.1f  fwd us %.1f  bwd us %.1f  dbl-buffered %.1f' % (d['ms_per_step']*1e3, d['roofline_kernels']['warp_fwd']['us_per_launch'], d['roofline_kernels']['warp_bwd']['us_per_launch'], d['dU_double_buffered']['ms_per_step']*1e3))"
done
