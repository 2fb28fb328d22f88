"""Extract per-launch DRAM traffic of one kernel from an ncu report and write profiles/ncu_traffic.json.
Usage: ncu_traffic.py report.ncu-rep kernel_substring LPs_per_launch m n out.json"""
import csv, json, subprocess, sys
rep, sub, lps, m, n, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
ci = {name: i for i, name in enumerate(h)}
best = None
for r in rows[2:]:
    if sub in r[ci['Kernel Name']]:
        best = r
if best is None:
    raise SystemExit('kernel not found')
def val(name):
    return float(best[ci[name]].replace(',', ''))
unit_r = rows[1][ci['dram__bytes_read.sum']]; unit_w = rows[1][ci['dram__bytes_write.sum']]
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
rd = val('dram__bytes_read.sum') * scale[unit_r]; wr = val('dram__bytes_write.sum') * scale[unit_w]
res = {'m': m, 'n': n, 'lps_per_launch': lps, 'dram_bytes_read': rd, 'dram_bytes_write': wr,
       'dram_bytes_per_lp': (rd + wr) / lps, 'kernel': best[ci['Kernel Name']],
       'duration_us': val('gpu__time_duration.sum') * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}[rows[1][ci['gpu__time_duration.sum']]],
       'source': 'ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum), %s, %d LPs per launch' % (rep.split('/')[-1], lps)}
json.dump(res, open(out, 'w'), indent=1)
print(json.dumps(res))
