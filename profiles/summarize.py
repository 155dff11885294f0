"""Turns the ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py gpurun_out/launches_r01c.csv gpurun_out/prof_r01c.ncu-rep r01

Writes profiles/launches_<tag>.csv (copy), profiles/ncu_<tag>_summary.json and profiles/ncu_<tag>_summary.md.
"""
import collections
import csv
import json
import re
import shutil
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def short(name):
    m = re.match(r"(?:void )?(?:nis::)?(\w+)<([^>]*?)(?:, (?:nis::)?(Pro\w+|Epi\w+|Mid\w+)(?:<\d>)?)?(?:, (?:nis::)?(Epi\w+))?>", name)
    if not m:
        return re.sub(r"\(.*", "", name)[:60]
    kern, targs = m.group(1), m.group(2)
    n = targs.split(",")[0].strip()
    ops = [x for x in (m.group(3), m.group(4)) if x]
    inv = ""
    if kern == "row_kernel":
        inv = " inv" if targs.split(",")[5].strip() == "1" else " fwd"
    return "%s<%s>%s %s" % (kern, n, inv, "+".join(ops))


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(launch_csv, rep, tag):
    shutil.copy(launch_csv, "profiles/launches_%s.csv" % tag)
    rows = list(csv.reader(open(launch_csv)))
    hdr, acc = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"]) * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1, "us": 1, "msecond": 1e3}.get(d["Metric Unit"], 1)
        k = short(d["Kernel Name"])
        acc[k][0] += 1
        acc[k][1] += v
    tot = sum(v[1] for v in acc.values())
    shares = {k: {"launches": v[0], "total_us": round(v[1], 1), "avg_us": round(v[1] / v[0], 2), "share": round(v[1] / tot, 4)}
              for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1])}
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    idx = {x: i for i, x in enumerate(h)}
    kernels = {}
    for r in rr[2:]:
        k = short(r[idx["Kernel Name"]])
        if k in kernels:
            continue
        e = {}
        for key in KEYS:
            if key in idx and r[idx[key]] != "":
                u = units[idx[key]]
                e[key] = to_bytes(r[idx[key]], u) if "byte" in u else float(r[idx[key]])
        e["dram_bytes_per_launch"] = e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0)
        kernels[k] = e
    out = {"tag": tag, "launch_list_shares": shares, "full_capture": kernels,
           "note": "ncu per-launch times are cold-cache and serialised: compare shares, not absolutes; dram bytes are per launch "
                   "at the captured batch size (see launch__grid_size)"}
    json.dump(out, open("profiles/ncu_%s_summary.json" % tag, "w"), indent=1)
    with open("profiles/ncu_%s_summary.md" % tag, "w") as f:
        f.write("# ncu summary %s\n\n## launch list (gpu__time_duration.sum, --clock-control none)\n\n" % tag)
        f.write("| kernel | launches | avg us | share |\n|---|---|---|---|\n")
        for k, v in shares.items():
            f.write("| %s | %d | %.2f | %.3f |\n" % (k, v["launches"], v["avg_us"], v["share"]))
        f.write("\n## full capture (--set full), one launch per kernel\n\n")
        f.write("| kernel | us | grid | regs | waves/SM | warps active % | issue active % | fma pipe % | Minst | smem bank conflicts | "
                "dram MB/launch | dram % | tensor % |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for k, e in kernels.items():
            g = lambda key, d=0.0: e.get(key, d)
            f.write("| %s | %.1f | %d | %d | %.2f | %.1f | %.1f | %.1f | %.2f | %d | %.1f | %.1f | %.1f |\n" % (
                k, g("gpu__time_duration.sum") / 1e3 if g("gpu__time_duration.sum") > 1e3 else g("gpu__time_duration.sum"),
                g("launch__grid_size"), g("launch__registers_per_thread"), g("launch__waves_per_multiprocessor"),
                g("sm__warps_active.avg.pct_of_peak_sustained_active"), g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), g("smsp__inst_executed.sum") / 1e6,
                g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), e["dram_bytes_per_launch"] / 1e6,
                g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")))
    print(open("profiles/ncu_%s_summary.md" % tag).read())


if __name__ == "__main__":
    main(*sys.argv[1:4])
