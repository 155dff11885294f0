#!/bin/bash
# batch/lane sweep of the tracking bench (run on the GPU box): tools/sweep.sh "14,3 14,4 7,6"
mkdir -p gpurun_out
for bl in $1; do
  b=${bl%,*}; l=${bl#*,}
  python bench.py --db 0 --cpu-frames 0 --batch $b --lanes $l --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('batch $b lanes $l value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
done
