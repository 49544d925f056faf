"""Summarise an .ncu-rep (raw page + hottest SASS lines) into text.  usage: ncu_summary.py rep [kernel-regex]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]; units = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    if name in seen: continue
    seen.add(name)
    print('=' * 100)
    for w in want:
        if w in idx and r[idx[w]] not in ('', '0'):
            print('%-78s %s %s' % (w[:78], r[idx[w]][:100], units[idx[w]]))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + name.split('(')[0].split('<')[0].split()[-1], '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) < 3: continue
    sh = srows[1]; six = {h: i for i, h in enumerate(sh)}
    body = [x for x in srows[2:] if len(x) > six['# Samples'] and x[six['# Samples']].isdigit()]
    body = body[:len(body) // 2] if len(body) > 3000 and body[0][six['Source']] == body[len(body) // 2][six['Source']] else body
    tot = sum(int(x[six['# Samples']]) for x in body) or 1
    texec = sum(int(x[six['Instructions Executed']]) for x in body)
    print('--- static SASS %d, warp instr executed %d, samples %d; hottest instructions (by stall samples):' % (len(body), texec, tot))
    top = sorted(enumerate(body), key=lambda t: -int(t[1][six['# Samples']]))[:22]
    for i, x in sorted(top):
        print('  %5d %5.1f%%  exec %9s  %s' % (i, 100 * int(x[six['# Samples']]) / tot, x[six['Instructions Executed']], x[six['Source']].strip()[:80]))
    hist = collections.Counter()
    for x in body:
        t = x[six['Source']].strip().split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        hist[op] += int(x[six['Instructions Executed']])
    print('--- opcode mix: ' + ', '.join('%s %.1f%%' % (o, 100 * c / max(texec, 1)) for o, c in hist.most_common(16)))
