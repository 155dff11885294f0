"""Registers / spills / static smem per kernel: python tools/regs.py nis_col.cu [filter] [extra nvcc flags...]"""
import re, subprocess, sys, os
HERE = os.path.dirname(os.path.abspath(__file__))
unit = sys.argv[1]; flt = sys.argv[2] if len(sys.argv) > 2 else ""; extra = sys.argv[3:]
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-c",
       os.path.join(HERE, "..", "ni_slam_b200", "csrc", unit), "-o", "/tmp/regs_tmp.o"] + extra
err = subprocess.run(cmd, capture_output=True, text=True).stderr
cur = None; rows = []
for l in err.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", l)
    if m: cur = [m.group(1), "", ""]; rows.append(cur); continue
    if cur is None: continue
    if "spill" in l: cur[2] = re.sub(r"\s+", " ", l.strip())
    m = re.search(r"Used (\d+) registers", l)
    if m: cur[1] = m.group(1)
    if "error" in l: print(l)
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
for n, r in sorted(zip(names, rows)):
    n = re.sub(r"\(int\)|nis::|void ", "", n); n = re.sub(r"\(Src.*|\(Pro.*|\(.*", "", n)
    if flt in n: print("%-100s regs %4s  %s" % (n[:100], r[1], "" if "0 bytes spill stores" in r[2] else r[2]))
