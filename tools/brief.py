"""Runs bench.py with the given arguments and prints a one-line digest (dev helper for gpurun sweeps).
usage: python tools/brief.py [ENV=VAL ...] -- <bench.py args>"""
import json
import os
import subprocess
import sys

args = sys.argv[1:]
env = dict(os.environ)
if "--" in args:
    i = args.index("--")
    for kv in args[:i]:
        k, v = kv.split("=", 1)
        env[k] = v
    args = args[i + 1:]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(root, "bench.py")] + args, capture_output=True, text=True, env=env)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
except Exception:
    print("FAILED", out.stdout[-2000:], out.stderr[-2000:])
    sys.exit(1)
k = d.get("roofline", {}).get("kernels", {})
top = {n: (v["share"], round(v["ms"] / max(v["launches"], 1) * 1e3, 1)) for n, v in list(k.items())[:8]}
lc = d.get("loop_closure") or {}
print(" ".join(sys.argv[1:]), "| value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2),
      "W", (d.get("clocks") or {}).get("power_w_max"), "| cand/s", round(lc.get("candidates_per_sec", 0)), "|", top)
