"""Per-region instruction and stall accounting from an ncu source-page CSV (SASS view): consecutive SASS lines with the same
execution count form a region.  usage: ncu_regions.py source.csv <kernel-substring> [min-share]"""
import csv, sys, io, collections
path, want = sys.argv[1], sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
txt = open(path).read()
blocks = txt.split('"Kernel Name",')
for b in blocks[1:]:
    lines = b.split('\n')
    name = lines[0]
    if want not in name:
        continue
    rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[1:] if len(r) > ix['Instructions Executed'] and r[ix['Instructions Executed']].isdigit()]
    tot = sum(int(r[ix['Instructions Executed']]) for r in body)
    tots = sum(int(r[ix['# Samples']]) for r in body) or 1
    totw = sum(int(r[ix['L1 Wavefronts Shared']] or 0) for r in body)
    print(name[:120]); print('SASS lines %d, warp-instr %d, samples %d, smem wavefronts %d' % (len(body), tot, tots, totw))
    regs = []; cur = None
    for i, r in enumerate(body):
        e = int(r[ix['Instructions Executed']])
        if cur is None or e != cur[2]:
            cur = [i, i, e, 0, 0, collections.Counter(), 0, 0]; regs.append(cur)
        cur[1] = i; cur[3] += e; cur[4] += int(r[ix['# Samples']])
        t = r[ix['Source']].strip().split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        cur[5][op] += 1
        cur[6] += int(r[ix['L1 Wavefronts Shared']] or 0); cur[7] += int(r[ix['L1 Wavefronts Shared Ideal']] or 0)
    for a, z, e, n, s, ops, wf, wfi in regs:
        if 100.0 * n / tot >= minshare or 100.0 * s / tots >= minshare:
            print('  [%5d-%5d] exec/line %9d  lines %4d  instr %5.1f%%  samples %5.1f%%  smem wf %8d (ideal %8d)  %s'
                  % (a, z, e, z - a + 1, 100.0 * n / tot, 100.0 * s / tots, wf, wfi, ' '.join('%s:%d' % kv for kv in ops.most_common(6))))
    break
