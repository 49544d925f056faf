#!/bin/bash
# A/B of forward variants: tools/ab_fwd.sh so1 so2 ... -> ncu duration of warp_fwd_pipe + ablate fwd line, per library
for so in "$@"; do
  echo "== $so"
  MGW_SO_NAME=$so bash tools/ncu_quick.sh gpurun_out/ab.csv warp_fwd_pipe | cut -c1-200
  MGW_SO_NAME=$so timeout 100 python tools/ablate.py 2>&1 | grep -E "fwd: out \+ black|mesh fwd"
done
