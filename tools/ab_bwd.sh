#!/bin/bash
# A/B of backward variants (MGW_BWD=pipe): tools/ab_bwd.sh so1 so2 ...
for so in "$@"; do
  echo "== $so"
  MGW_BWD=pipe MGW_SO_NAME=$so bash tools/ncu_quick.sh gpurun_out/ab.csv warp_bwd_pipe | cut -c1-220
done
