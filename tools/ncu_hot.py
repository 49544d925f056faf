"""Hottest SASS lines (by stall samples) of one kernel of an .ncu-rep.  usage: ncu_hot.py rep kernel-regex [top]"""
import csv, subprocess, sys, io, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > ix['# Samples'] and r[ix['# Samples']].isdigit()]
tot = sum(int(r[ix['# Samples']]) for r in body) or 1
texec = sum(int(r[ix['Instructions Executed']]) for r in body)
print('static SASS %d, warp instr executed %d, samples %d' % (len(body), texec, tot))
best = sorted(enumerate(body), key=lambda t: -int(t[1][ix['# Samples']]))[:top]
for i, r in sorted(best):
    print('%5d %5.1f%% exec %9s  %s' % (i, 100 * int(r[ix['# Samples']]) / tot, r[ix['Instructions Executed']], r[ix['Source']].strip()[:90]))
hist = collections.Counter()
for r in body:
    t = r[ix['Source']].strip().split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    hist[op] += int(r[ix['Instructions Executed']])
print('opcode mix: ' + ', '.join('%s %.1f%%' % (o, 100 * c / max(texec, 1)) for o, c in hist.most_common(18)))
