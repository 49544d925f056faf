"""Opcode histogram per kernel of libmgw_b200.so (cuobjdump -sass): the committed evidence for the TMA / mbarrier / shared-atomic /
reduction instructions DESIGN.md cites.  usage: python tools/sass_counts.py [so] > profiles/r02_sass_counts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'deep-online-video-stabilization_b200', 'libmgw_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
WATCH = ['UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTMACCTL', 'SYNCS', 'ATOMS', 'REDG', 'ATOMG', 'RED', 'LDS', 'STS', 'LDG', 'STG', 'FFMA2', 'FMUL2', 'FFMA',
         'FMUL', 'FADD', 'MUFU', 'SHFL', 'BAR', 'CCTL', 'ACQBULK', 'HMMA', 'UTCMMA', 'STL', 'LDL']
kern, hist, total = None, {}, collections.Counter()
for ln in txt.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace('(anonymous namespace)::', '').replace('void ', '').replace('mgw::', '')
        kern = re.sub(r'\(.*', '', kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
        total[m.group(1)] += 1
print('# static SASS opcode counts per kernel, %s (sm_100a), cuobjdump -sass' % os.path.basename(so))
print('# whole library: ' + ', '.join('%s %d' % (k, total[k]) for k in WATCH if total[k]))
print('# tensor-core opcodes (HMMA / UTCMMA ...): %d -- the path is a gather, not a contraction' % sum(v for k, v in total.items() if 'MMA' in k))
for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
    h = hist[k]
    if sum(h.values()) < 40:
        continue
    print('%-78s %5d instr | %s' % (k[:78], sum(h.values()), ' '.join('%s:%d' % (o, h[o]) for o in WATCH if h[o])))
