#!/bin/bash
# DRAM bytes of the step IN PLACE: ncu without cache control (no flush, no replay: one pass for the three metrics), the kernels of
# five consecutive steps over rotating inputs; summarised per step by tools/ncu_step_traffic.py
# usage: tools/ncu_step_traffic.sh <tag>
tag=${1:-steptraffic}
cd "$(dirname "$0")/.."
PROF_NOFLUSH=1 timeout 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  -k regex:"solve_h|warp_|fill_zero" --csv --log-file gpurun_out/${tag}.csv python tools/prof_step.py 32 8 > gpurun_out/${tag}.log 2>&1
python tools/ncu_step_traffic.py gpurun_out/${tag}.csv
