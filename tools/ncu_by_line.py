"""Aggregate an ncu SASS-page CSV (ncu -i X.ncu-rep --page source --csv --print-source sass) by source line, using
the line table of the same cubin (nvdisasm -g -c).  Usage: ncu_by_line.py sass.csv kernel.dis [top]"""
import csv, re, sys, collections
sass_csv, dis, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
outer = len(sys.argv) > 4 and sys.argv[4] == 'outer'   # attribute to the outermost call site (needs nvdisasm -gi)
addr2line = {}; cur = None; fresh = True
for l in open(dis):
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)(?: inlined at ".*?([^/"]+)", line (\d+))?', l)
    if m:
        inner = '%s:%s' % (m.group(1), m.group(2))
        outr = '%s:%s' % (m.group(3), m.group(4)) if m.group(3) else inner
        if outer: cur = outr
        elif fresh: cur = inner
        fresh = False
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
    if m: addr2line[int(m.group(1), 16)] = cur; fresh = True
rows = list(csv.reader(open(sass_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; ci = {n: i for i, n in enumerate(hdr)}
ex = collections.Counter(); st = collections.Counter(); tot_ex = 0; tot_st = 0
stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
stall_tot = collections.Counter(); per_line_stall = collections.defaultdict(collections.Counter)
base = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[ci['Address']], 16)
    if base is None: base = a
    line = addr2line.get(a - base, '?')
    e = int(r[ci['Instructions Executed']] or 0); s = int(r[ci['Warp Stall Sampling (All Samples)']] or 0)
    ex[line] += e; st[line] += s; tot_ex += e; tot_st += s
    for n in stall_cols:
        v = int(r[ci[n]] or 0); stall_tot[n] += v; per_line_stall[line][n] += v
print('total executed warp-instr %d, samples %d' % (tot_ex, tot_st))
print('stall totals:', ', '.join('%s %.1f%%' % (k[6:], 100.0 * v / max(tot_st, 1)) for k, v in stall_tot.most_common(8)))
print('%-32s %12s %6s %10s %6s  top stalls' % ('line', 'executed', '%', 'samples', '%'))
for line, s in st.most_common(top):
    tops = ', '.join('%s %d' % (k[6:], v) for k, v in per_line_stall[line].most_common(3))
    print('%-32s %12d %6.2f %10d %6.2f  %s' % (line, ex[line], 100.0 * ex[line] / tot_ex, s, 100.0 * s / tot_st, tops))
