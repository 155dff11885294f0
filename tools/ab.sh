#!/bin/bash
# A/B of env-selected variants on the GPU box: tools/ab.sh "NAME=ENV=VAL,ENV2=VAL2 NAME2=..."  (NAME=- for no env)
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%=*}; envs=${spec#*=}; [ "$envs" = "-" ] && envs=""
  env $(echo $envs | tr ',' ' ') python bench.py --db 0 --cpu-frames 0 --no-cfg4 --no-extras --steps 6 --warmup 3 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<PY
import json,sys
n=sys.argv[1]
try:
    d=json.load(open("gpurun_out/ab_%s.json"%n)); print(n, round(d["value"]), round(d["e2e"]["value"]), {k:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()})
except Exception as e: print(n, "ERR", e, open("gpurun_out/ab_%s.err"%n).read()[-500:])
PY
done
