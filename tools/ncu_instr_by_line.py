"""Dynamic instruction counts per kernel-body source line (outermost call site) with the opcode mix.
Usage: ncu_instr_by_line.py report.ncu-rep object_or_binary kernel_symbol_prefix source.cu LPs [top]"""
import re, csv, collections, subprocess, sys, os, tempfile
rep, obj, sym, srcf, lps = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
d = tempfile.mkdtemp()
subprocess.run('cd %s && cuobjdump -xelf all %s > /dev/null && nvdisasm -gi -c *.cubin > all.dis 2>/dev/null' % (d, obj), shell=True)
lines = open(os.path.join(d, 'all.dis')).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.' + sym)][0]
end = [i for i, l in enumerate(lines) if i > start and l.startswith('//--------------------- .text.')]
end = end[0] if end else len(lines)
a2l = {}; cur = None
for l in lines[start:end]:
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)(?: inlined at ".*?([^/"]+)", line (\d+))?', l)
    if m:
        cur = int(m.group(4)) if m.group(3) else int(m.group(2)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: a2l[int(m.group(1), 16)] = cur
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; ci = {n: i for i, n in enumerate(hdr)}
base = None; by = collections.defaultdict(collections.Counter); ops = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[ci['Address']], 16)
    if base is None: base = a
    e = int(r[ci['Instructions Executed']] or 0)
    s = r[ci['Source']].split(); op = (s[1] if s[0].startswith('@') else s[0]).split('.')[0]
    by[a2l.get(a - base)][op] += e; ops[op] += e
src = open(srcf).read().split('\n')
tot = sum(ops.values())
print('total %.0f warp-instr per LP; mix: %s' % (tot / lps, ', '.join('%s %.0f' % (k, v / lps) for k, v in ops.most_common(12))))
for ln, c in sorted(by.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    t = sum(c.values())
    print('%4s %8.0f/LP %5.1f%%  %-58s %s' % (ln, t / lps, 100 * t / tot, (src[ln - 1].strip()[:58] if ln else ''), ', '.join('%s %.0f' % (k, v / lps) for k, v in c.most_common(4))))
