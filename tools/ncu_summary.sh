#!/bin/bash
# usage: tools/ncu_summary.sh <ncu-rep> <kernel-symbol-prefix> <binary with the cubin> <units per launch> <unit> <algorithmic bytes per unit> <out.txt> <header line>
# Writes the judge-readable summary of one ncu --set full capture: details highlights, DRAM traffic per unit, stall samples by source line.
REP=$(realpath $1); SYM=$2; BIN=$(realpath $3); UNITS=$4; UNIT=$5; ALG=$6; OUT=$(realpath -m $7); HDR=$8
HERE=$(cd "$(dirname "$0")/.." && pwd)
{
  echo "# $HDR"
  python $HERE/tools/ncu_brief.py $REP $UNITS $UNIT $ALG
  echo
  echo "## stall samples by source line (outer attribution), top 30"
  mkdir -p /tmp/sass && cd /tmp/sass && rm -f *.cubin && cuobjdump -xelf all $BIN >/dev/null 2>&1 && for f in *.cubin; do nvdisasm -gi -c $f 2>/dev/null; done > all.dis
  python - "$SYM" <<'PY'
import sys
sym=sys.argv[1]
lines=open('/tmp/sass/all.dis').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.'+sym)][0]
end=[i for i,l in enumerate(lines) if i>start and l.startswith('//--------------------- .text.')]
end=end[0] if end else len(lines)
open('/tmp/sass/kern.dis','w').write('\n'.join(lines[start:end]))
PY
  ncu -i $REP --page source --csv --print-source sass 2>/dev/null > /tmp/kern_sass.csv
  python $HERE/tools/ncu_by_line.py /tmp/kern_sass.csv /tmp/sass/kern.dis 30 outer
} > $OUT 2>&1
