#!/bin/bash
# A/B of library variants on tracking + scan: tools/scan_ab.sh name=libpath ...
for spec in "$@"; do
  name=${spec%%=*}; lib=${spec#*=}
  NIS_LIB=$lib python bench.py --cpu-frames 0 --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); l=d['loop_closure']; print('$name value %.0f e2e %.0f scan cand/s %.0f q/s %.2f' % (d['value'], d['e2e']['value'], l['candidates_per_sec'], l['value']))"
done
