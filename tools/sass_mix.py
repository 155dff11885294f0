"""Static SASS instruction mix per kernel of the built objects (python tools/sass_mix.py [filter])."""
import collections, re, subprocess, sys, os
HERE = os.path.dirname(os.path.abspath(__file__))
flt = sys.argv[1:] or [""]
for unit in ("nis_row", "nis_col"):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(HERE, "..", "ni_slam_b200", "build", unit + ".o")], capture_output=True, text=True).stdout
    cur = None; d = {}
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m: cur = m.group(1); d[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m and cur: d[cur][m.group(2).split('.')[0]] += 1
    for k, c in d.items():
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void nis::", "")
        if not all(f in name for f in flt): continue
        tot = sum(c.values())
        fp1 = sum(c[x] for x in ("FADD", "FMUL", "FFMA")); fp2 = sum(c[x] for x in ("FADD2", "FMUL2", "FFMA2"))
        print("%-70s tot %5d fp %4d fp2 %4d MOV %3d LDS %3d STS %3d LDG %3d" % (name[:70], tot, fp1, fp2, c["MOV"], c["LDS"], c["STS"], c["LDG"]))
