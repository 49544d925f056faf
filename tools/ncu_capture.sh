#!/bin/bash
# ncu --set full capture of the warp kernels at config #2 + CSV exports (raw and source pages) into gpurun_out/<tag>_*
# usage: tools/ncu_capture.sh <tag> [kernel-regex] [extra env assignments...]
tag=${1:-cap}; rx=${2:-warp_.*kernel}
cd "$(dirname "$0")/.."
python tools/prof_step.py 32 2 > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$rx -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-2} -o gpurun_out/${tag} -f python tools/prof_step.py 32 2 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${tag}.ncu-rep > gpurun_out/${tag}_summary.txt 2>&1
ls -la gpurun_out/${tag}*
