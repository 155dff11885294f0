"""Dynamic (executed) SASS instruction mix per kernel from an ncu report captured with --set full --import-source on.
    python tools/dyn_mix.py gpurun_out/prof.ncu-rep [kernel-regex]
Prints warp-instructions executed per opcode class and the stall samples attributed to each class."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; flt = sys.argv[2] if len(sys.argv) > 2 else "."
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + flt],
                     capture_output=True, text=True).stdout
CLASS = [("fp32x2", r"^(FADD2|FMUL2|FFMA2)"), ("fp32", r"^(FADD|FMUL|FFMA|FMNMX|FSEL|FSETP|FCHK|MUFU|F2F|FRND)"), ("fp64", r"^D(ADD|MUL|FMA|SETP)|^F2F\.F64|^I2F\.F64|^F2I\.F64"),
         ("cvt", r"^(I2F|F2I|I2I|I2FP|F2IP)"), ("lds", r"^LDS"), ("sts", r"^STS"), ("ldg", r"^(LDG|LD\.)"), ("stg", r"^(STG|ST\.)"), ("ldc", r"^(LDC|ULDC|LDCU)"),
         ("atom", r"^(ATOM|RED|ATOMS|ATOMG)"), ("shfl", r"^(SHFL|VOTE|REDUX|MATCH)"), ("bar", r"^(BAR|MEMBAR|WARPSYNC|BSSY|BSYNC|NANOSLEEP|SYNCS|ERRBAR)"), ("tma", r"^(UTMA|UBLKCP|UTMALDG)"),
         ("branch", r"^(BRA|EXIT|RET|CALL|JMP|BRX|BREAK)"), ("mov", r"^(MOV|IMAD\.MOV|UMOV|PRMT|SEL|R2UR|S2R|S2UR|CS2R|SHFL)"),
         ("int", r"^(IMAD|IADD|IADD3|LEA|LOP3|SHF|ISETP|IMNMX|ULEA|UIADD3|UIMAD|ULOP3|USHF|UISETP|VIADD|VIMNMX|IABS|POPC|FLO|BMSK|SGXT|UFLO|USEL|PLOP3|UPLOP3|P2R|R2P|LOP)")]
cur = None; acc = {}
rd = csv.reader(io.StringIO(out))
hdr = None
for row in rd:
    if not row: continue
    if row[0] == "Kernel Name": cur = re.sub(r"\(int\)|nis::|void ", "", row[1])[:90]; acc.setdefault(cur, [collections.Counter(), collections.Counter()]); hdr = None; continue
    if row[0] == "Address": hdr = {h: i for i, h in enumerate(row)}; continue
    if hdr is None or cur is None: continue
    src = row[hdr["Source"]].strip()
    src = re.sub(r"^@!?U?P\d+\s+", "", src)
    op = src.split()[0] if src else "?"
    n = float(row[hdr["Instructions Executed"]] or 0); st = float(row[hdr["# Samples"]] or 0)
    cls = next((c for c, rx in CLASS if re.match(rx, op)), "other:" + op.split(".")[0])
    acc[cur][0][cls] += n; acc[cur][1][cls] += st
for k, (ins, st) in acc.items():
    tot = sum(ins.values()); ts = sum(st.values()) or 1
    print("%s\n  total warp-instr %.3fM" % (k, tot / 1e6))
    for c, v in ins.most_common():
        print("   %-14s %9.0f  %5.1f%%   stall samples %5.1f%%" % (c, v, 100 * v / tot, 100 * st[c] / ts))
