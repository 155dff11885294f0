#!/bin/bash
# batch x lanes sweep on the GPU box: tools/sweep.sh "14 28 56 112" "1 2 3 4"
mkdir -p gpurun_out
for b in $1; do for l in $2; do
  python bench.py --db 0 --cpu-frames 0 --no-cfg4 --no-extras --steps 5 --warmup 3 --batch $b --lanes $l > gpurun_out/sw_${b}_${l}.json 2> gpurun_out/sw_${b}_${l}.err
  python - $b $l <<PY
import json,sys
b,l=sys.argv[1:3]
try:
    d=json.load(open("gpurun_out/sw_%s_%s.json"%(b,l))); print("batch",b,"lanes",l, round(d["value"]), round(d["e2e"]["value"]))
except Exception as e: print(b,l,"ERR",e, open("gpurun_out/sw_%s_%s.err"%(b,l)).read()[-300:])
PY
done; done
