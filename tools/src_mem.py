"""Per-instruction memory behaviour of one kernel from an ncu source-page CSV (tools/profile.sh): python tools/src_mem.py file.csv.gz 'kernel substring'"""
import csv, gzip, io, sys
txt = gzip.open(sys.argv[1], "rt").read()
blocks = txt.split('"Kernel Name",')[1:]
for b in blocks:
    name, rest = b.split("\n", 1)
    if sys.argv[2] not in name:
        continue
    rows = list(csv.reader(io.StringIO(rest)))
    hdr = rows[0]; H = {h: i for i, h in enumerate(hdr)}
    print(name[:150])
    tot = {"inst": 0, "tag": 0, "sec": 0, "ideal": 0, "shw": 0, "shi": 0}
    print("%-70s %9s %9s %9s %9s %9s %9s" % ("SASS", "executed", "L1 tags", "sectors", "ideal", "sh wavef", "sh ideal"))
    for r in rows[1:]:
        if len(r) < len(hdr): continue
        f = lambda k: float(r[H[k]] or 0)
        ins = f("Instructions Executed"); tot["inst"] += ins
        tag, sec, idl, shw, shi = f("L1 Tag Requests Global"), f("L2 Theoretical Sectors Global"), f("L2 Theoretical Sectors Global Ideal"), f("L1 Wavefronts Shared"), f("L1 Wavefronts Shared Ideal")
        tot["tag"] += tag; tot["sec"] += sec; tot["ideal"] += idl; tot["shw"] += shw; tot["shi"] += shi
        if tag or shw:
            print("%-70s %9.0f %9.0f %9.0f %9.0f %9.0f %9.0f" % (r[H["Source"]].strip()[:70], ins, tag, sec, idl, shw, shi))
    print("TOTAL", tot)
    break
