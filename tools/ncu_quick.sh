#!/bin/bash
# quick per-launch metrics of kernels matching a regex: tools/ncu_quick.sh <out.csv> <regex> [env...]
out=$1; rx=$2
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv -k regex:$rx --log-file $out python tools/prof_step.py 32 2 > /dev/null 2>&1
python - $out <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
d={}
for r in rows[h+1:]: d.setdefault(r[0],{})[r[-3]]=r[-1]
for k,v in d.items(): print(k, v)
PY
