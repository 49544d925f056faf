"""profiles/traffic.json from ncu raw-page CSVs (dram__bytes_read.sum + dram__bytes_write.sum per launch of the two hot kernels).
usage: python tools/ncu_traffic.py <fwd_raw.csv> <bwd_raw.csv> > profiles/traffic.json   (the CSVs come from tools/ncu_capture.sh)"""
import csv, json, sys


def traffic(path, needle):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if needle in r[ix['Kernel Name']]:
            def val(name):
                v, u = float(r[ix[name]].replace(',', '')), units[ix[name]]
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
            return val('dram__bytes_read.sum') + val('dram__bytes_write.sum'), r[ix['Kernel Name']][:90], float(r[ix['gpu__time_duration.sum']])
    raise SystemExit('no kernel matching %r in %s' % (needle, path))


fwd, fname, fus = traffic(sys.argv[1], 'warp_fwd')
bwd, bname, bus = traffic(sys.argv[2], 'warp_bwd')
print(json.dumps({'warp_fwd': fwd, 'warp_bwd': bwd,
                  'source': 'ncu --set full raw pages %s, %s: dram__bytes_read.sum + dram__bytes_write.sum per launch at config #2 '
                            '(%s: %.1f us; %s: %.1f us); caches flushed before each profiled launch, so the backward figure includes the '
                            're-read of the zero-filled dU lines that stay in L2 in the real step' % (sys.argv[1], sys.argv[2], fname, fus, bname, bus)}, indent=1))
