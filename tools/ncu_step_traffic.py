#!/usr/bin/env python
"""per-step DRAM bytes from the CSV of tools/ncu_step_traffic.sh: the last five steps (K1, K2, fill, K3, K4), median per kernel;
prints a JSON fragment for profiles/traffic.json"""
import csv
import json
import statistics as st
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
per = {}
for r in rows[h + 1:]:
    name = r[ix['Kernel Name']]
    key = ('solve_h_fwd' if 'solve_h_fwd' in name else 'solve_h_bwd' if 'solve_h_bwd' in name else 'fill_zero' if 'fill_zero' in name else
           'warp_fwd' if 'warp_fwd' in name else 'warp_bwd')
    v = float(r[ix['Metric Value']].replace(',', ''))
    u = r[ix['Metric Unit']]
    v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(u, 1)
    per.setdefault((int(r[ix['ID']]), key), {})[r[ix['Metric Name']]] = v
by = {}
for (i, key), m in sorted(per.items()):
    by.setdefault(key, []).append(m)
out = {}
for key, ms in by.items():
    ms = ms[3:]                      # the first three steps warm the rotating sets up
    out[key] = {'dram_read': st.median(m['dram__bytes_read.sum'] for m in ms), 'dram_write': st.median(m['dram__bytes_write.sum'] for m in ms),
                'us': st.median(m['gpu__time_duration.sum'] for m in ms), 'launches': len(ms)}
tot = sum(v['dram_read'] + v['dram_write'] for v in out.values())
fb = sum(out[k]['dram_read'] + out[k]['dram_write'] for k in ('fill_zero', 'warp_bwd') if k in out)
print(json.dumps({'step_in_place': {'per_kernel': out, 'step_total': tot, 'step_over_algorithmic_377.5MB': tot / 377487360.0,
                                    'fill_plus_bwd': fb, 'fill_plus_bwd_over_algorithmic_207.6MB': fb / 207618048.0,
                                    'source': 'tools/ncu_step_traffic.sh: ncu --cache-control none, kernels serialised but caches left as '
                                              'the previous kernel left them, medians over the last five of eight steps'}}, indent=1))
