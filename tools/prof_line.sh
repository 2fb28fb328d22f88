#!/bin/bash
# usage: tools/prof_line.sh <ncu-rep> <kernel-symbol-prefix> [top]   -- per-source-line ncu summary of one kernel
REP=$1; SYM=$2; TOP=${3:-40}; MODE=${4:-inner}
OBJ=${5:-/root/repo/deep_dantzig_b200/csrc/simplex_rowreg.o}
mkdir -p /tmp/sass && cd /tmp/sass && rm -f *.cubin && cuobjdump -xelf all $OBJ >/dev/null && nvdisasm -gi -c *.cubin > all.dis 2>/dev/null
python - "$SYM" <<'PY'
import sys
sym=sys.argv[1]
lines=open('/tmp/sass/all.dis').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.'+sym)][0]
end=[i for i,l in enumerate(lines) if i>start and l.startswith('//--------------------- .text.')]
end=end[0] if end else len(lines)
open('/tmp/sass/kern.dis','w').write('\n'.join(lines[start:end]))
PY
ncu -i $REP --page source --csv --print-source sass 2>/dev/null > /tmp/kern_sass.csv && python /root/repo/tools/ncu_by_line.py /tmp/kern_sass.csv /tmp/sass/kern.dis $TOP $MODE
