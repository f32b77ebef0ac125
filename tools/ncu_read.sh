#!/bin/bash
# Read a capture made by tools/ncu_capture.sh here (no GPU): raw.csv + per-kernel instruction mix / stall sites.
# usage: tools/ncu_read.sh <dir> <kernel name fragment>...
d=$1; shift
ncu -i $d/prof.ncu-rep --page raw --csv > $d/raw.csv 2>/dev/null
python - "$d" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1] + '/raw.csv'))); H = rows[0]
for r in rows[2:]:
    d = dict(zip(H, r))
    g = lambda k: d.get(k, '?')
    print(g('Kernel Name')[:56].ljust(56), 'us', g('gpu__time_duration.sum'), '| issue%', g('sm__issue_active.avg.pct_of_peak_sustained_elapsed'),
          '| dram%', g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), '| tensor%', g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
          '| MB', round((float(g('dram__bytes_read.sum').replace(',', '')) + float(g('dram__bytes_write.sum').replace(',', ''))) , 1), H and rows[1][H.index('dram__bytes_read.sum')],
          '| inst', g('smsp__inst_executed.sum'))
PY
for k in "$@"; do
  ncu -i $d/prof.ncu-rep --page source --csv --kernel-name regex:$k > $d/src_$k.csv 2>/dev/null
  python tools/ncu_source_hot.py $d/src_$k.csv 22
done
