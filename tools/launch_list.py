#!/usr/bin/env python
"""print an `ncu --metrics gpu__time_duration.sum --csv` launch list as  <us>  <kernel name>  lines + the total"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[h]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = 0.0
for r in rows[h + 1:]:
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    tot += v
    print('%8.1f  %s' % (v, r[ki][:120]))
print('total %.1f us in %d launches' % (tot, len(rows) - h - 1))
