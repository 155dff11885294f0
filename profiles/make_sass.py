"""SASS evidence for the production kernels: python profiles/make_sass.py [tag]

cuobjdump -sass of ni_slam_b200/lib/libnislam.so, one gzipped listing per kernel instantiation the 640x480 / 1280x960 paths launch
(profiles/sass_<tag>/<kernel>.sass.gz) and profiles/sass_<tag>_summary.md: instructions per kernel and the counts of the opcodes
that show what the hardware is asked to do (packed f32x2 math, TMA, mbarrier, shared/global memory, barriers, local-memory spills).
"""
import collections
import gzip
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ni_slam_b200", "lib", "libnislam.so")
SIZES = ("480", "640", "720", "960", "1280", "1200")      # transform lengths of the 640x480 and 1280x960 configurations (polar 720x480)
GROUPS = [("packed f32x2", r"^(FADD2|FMUL2|FFMA2)"), ("scalar fp32", r"^(FADD|FMUL|FFMA)(\.|$)"), ("fp64", r"^D(ADD|MUL|FMA)"),
          ("TMA tensor (UTMALDG)", r"^UTMALDG"), ("bulk copy (UBLKCP)", r"^UBLKCP"), ("mbarrier (SYNCS)", r"^SYNCS"),
          ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("BAR", r"^BAR"), ("SHFL", r"^SHFL"),
          ("atomics (ATOM/RED)", r"^(ATOM|RED)"), ("local (LDL/STL)", r"^(LDL|STL)")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(d):
    d = re.sub(r"\(.*", "", d.replace("void ", "").replace("nis::", "").replace("(anonymous namespace)::", ""))
    return re.sub(r"\((int|bool)\)", "", d)


def main(tag):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur:
            funcs[cur].append(line)
    dm = demangle(list(funcs))
    outdir = os.path.join(ROOT, "profiles", "sass_%s" % tag)
    os.makedirs(outdir, exist_ok=True)
    rows = []
    for f, lines in funcs.items():
        name = short(dm[f])
        m = re.match(r"\w+<(\d+)", name)
        if m and m.group(1) not in SIZES:
            continue                                     # test-only transform lengths (64, 80, 96)
        ops = [mm.group(1) for mm in (re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l) for l in lines) if mm]
        cnt = [sum(1 for o in ops if re.match(rx, o)) for _, rx in GROUPS]
        fn = re.sub(r"[^A-Za-z0-9_]+", "_", name).strip("_")[:110] + ".sass.gz"
        with gzip.open(os.path.join(outdir, fn), "wt") as fh:
            fh.write("Function : %s\n%s\n" % (dm[f], "\n".join(lines)))
        rows.append((name, len(ops), cnt, fn))
    rows.sort()
    with open(os.path.join(ROOT, "profiles", "sass_%s_summary.md" % tag), "w") as fh:
        fh.write("# SASS of the production kernels (%s, `cuobjdump -sass ni_slam_b200/lib/libnislam.so`, sm_100a)\n\n" % tag)
        fh.write("Static instruction counts per kernel instantiation; full listings in `profiles/sass_%s/*.sass.gz` "
                 "(regenerate: `python profiles/make_sass.py %s`).\n\n" % (tag, tag))
        fh.write("| kernel | instr | " + " | ".join(g for g, _ in GROUPS) + " |\n|---|---|" + "---|" * len(GROUPS) + "\n")
        for name, n, cnt, fn in rows:
            fh.write("| `%s` | %d | " % (name, n) + " | ".join(str(c) if c else "" for c in cnt) + " |\n")
        tot = [sum(r[2][i] for r in rows) for i in range(len(GROUPS))]
        fh.write("\n%d kernels; totals: " % len(rows) + ", ".join("%s %d" % (g, t) for (g, _), t in zip(GROUPS, tot)) + "\n")
    print(len(rows), "kernels ->", outdir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
