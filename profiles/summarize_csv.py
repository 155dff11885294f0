"""Turns the ncu CSV artefacts of tools/profile.sh (brought back in gpurun_out/) into the committed summaries under profiles/.

    python profiles/summarize_csv.py r02

Reads gpurun_out/{launches_<tag>.csv, launches_<tag>_scan.csv, raw_<tag>_track.csv, raw_<tag>_scan.csv}; writes copies of the launch
lists, profiles/ncu_<tag>_summary.json and profiles/ncu_<tag>_summary.md.  The .ncu-rep files themselves (3 MB per kernel with
--import-source on) are not kept; the raw-page CSVs hold every metric of the --set full capture.
"""
import collections
import csv
import gzip
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "mio_throttle", "lg_throttle", "math_pipe_throttle", "wait",
          "not_selected", "dispatch_stall", "branch_resolving", "no_instruction"]
TIME = {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = re.sub(r"void |nis::|\((int|bool)\)", "", name)
    return re.sub(r"\(.*", "", name)


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def read_csv(path):
    with open(path) as fh:
        return list(csv.reader(l for l in fh if not l.startswith("==")))


def launch_list(path, metrics=("gpu__time_duration.sum",)):
    """[(kernel, grid, {metric: value in us / bytes})] in launch order"""
    rows = read_csv(path)
    hdr = rows[0]
    out, byid = [], {}
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        if d["ID"] not in byid:
            byid[d["ID"]] = (short(d["Kernel Name"]), d["Grid Size"], {})
            out.append(byid[d["ID"]])
        v, u = fnum(d["Metric Value"]), d["Metric Unit"]
        byid[d["ID"]][2][d["Metric Name"]] = v * (TIME.get(u) or BYTES.get(u) or 1)
    return out


def shares(launches):
    acc = collections.OrderedDict()
    for k, _, m in launches:
        a = acc.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += m["gpu__time_duration.sum"]
    tot = sum(a[1] for a in acc.values())
    return tot, {k: {"launches": a[0], "total_us": round(a[1], 1), "avg_us": round(a[1] / a[0], 2), "share": round(a[1] / tot, 4)}
                 for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1])}


def full_capture(path):
    rows = read_csv(path)
    hdr, units, data = rows[0], rows[1], rows[2:]
    H = {h: i for i, h in enumerate(hdr)}
    acc = collections.OrderedDict()
    for d in data:
        def g(k):
            return fnum(d[H[k]])
        rec = {
            "us": g("gpu__time_duration.sum") * TIME[units[H["gpu__time_duration.sum"]]],
            "grid": g("launch__grid_size"), "block": g("launch__block_size"), "regs": g("launch__registers_per_thread"),
            "smem_per_block_KB": (g("launch__shared_mem_per_block_dynamic") + g("launch__shared_mem_per_block_static")) *
                                 BYTES.get(units[H["launch__shared_mem_per_block_dynamic"]], 1) / 1e3 if "launch__shared_mem_per_block_dynamic" in H else None,
            "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "l1tex_pct": g("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l2_pct": g("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "dram_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "tensor_pipe_pct": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in H else 0.0,
            "warp_inst_M": g("smsp__inst_executed.sum") / 1e6,
            "smem_bank_conflict_pct": 100.0 * g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / max(1.0, g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")),
            "dram_read_MB": g("dram__bytes_read.sum") * BYTES[units[H["dram__bytes_read.sum"]]] / 1e6,
            "dram_write_MB": g("dram__bytes_write.sum") * BYTES[units[H["dram__bytes_write.sum"]]] / 1e6,
        }
        for s in STALLS:
            rec["stall_" + s] = g("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s)
        acc.setdefault(short(d[H["Kernel Name"]]), []).append(rec)
    out = collections.OrderedDict()
    for k, recs in acc.items():
        m = {f: (sum(r[f] for r in recs) / len(recs) if recs[0][f] is not None else None) for f in recs[0]}
        m["captures"] = len(recs)
        out[k] = {f: (round(v, 3) if isinstance(v, float) else v) for f, v in m.items()}
    return out


def table(cap, fh):
    fh.write("| kernel | n | µs | grid×block | regs | warps act % | issue % | fma % | l1tex % | L2 % | DRAM % | Minst | smem conflicts % | DRAM rd+wr MB | "
             "stall cycles per issue: barrier / long_sb / short_sb / wait / not_selected / math_throttle / mio |\n|" + "---|" * 15 + "\n")
    for k, m in sorted(cap.items(), key=lambda kv: -kv[1]["us"] * kv[1]["captures"]):
        fh.write("| `%s` | %d | %.1f | %d×%d | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.2f | %.1f | %.1f + %.1f | %.2f / %.2f / %.2f / %.2f / %.2f / %.2f / %.2f |\n" % (
            k, m["captures"], m["us"], m["grid"], m["block"], m["regs"], m["warps_active_pct"], m["issue_active_pct"], m["fma_pipe_pct"],
            m["l1tex_pct"], m["l2_pct"], m["dram_pct"], m["warp_inst_M"], m["smem_bank_conflict_pct"], m["dram_read_MB"], m["dram_write_MB"],
            m["stall_barrier"], m["stall_long_scoreboard"], m["stall_short_scoreboard"], m["stall_wait"], m["stall_not_selected"],
            m["stall_math_pipe_throttle"], m["stall_mio_throttle"]))


def main(tag):
    go = os.path.join(ROOT, "gpurun_out")
    for f in ("launches_%s.csv" % tag, "launches_%s_scan.csv" % tag):
        shutil.copy(os.path.join(go, f), os.path.join(ROOT, "profiles", f))
    for f in ("raw_%s_track.csv" % tag, "raw_%s_scan.csv" % tag):
        with open(os.path.join(go, f), "rb") as src, gzip.open(os.path.join(ROOT, "profiles", "ncu_" + f + ".gz"), "wb") as dst:
            shutil.copyfileobj(src, dst)
    track = launch_list(os.path.join(go, "launches_%s.csv" % tag))
    t_tot, t_sh = shares(track)
    scan_all = launch_list(os.path.join(go, "launches_%s_scan.csv" % tag))
    # the second query = from the second select_count_kernel to the end of the list / the next scan_reduce_kernel
    starts = [i for i, (k, _, _) in enumerate(scan_all) if k == "select_count_kernel"]
    ends = [i for i, (k, _, _) in enumerate(scan_all) if k == "scan_reduce_kernel"]
    q = scan_all[starts[-1]:(ends[-1] + 1 if ends and ends[-1] > starts[-1] else len(scan_all))] if starts else scan_all
    s_tot, s_sh = shares(q)
    n_cand = 2048
    dram = sum(m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0) for _, _, m in q)
    cap_t = full_capture(os.path.join(go, "raw_%s_track.csv" % tag))
    cap_s = full_capture(os.path.join(go, "raw_%s_scan.csv" % tag))
    frames = 1000
    inst_per_solve = None
    js = {"tag": tag, "tracking": {"what": "one step = 1000 frames 640x480 (18 batches of 56 on 3 lanes), ncu launch list, cold-cache serialised",
                                   "launches": len(track), "total_us": round(t_tot, 1), "shares": t_sh, "full_capture": cap_t},
          "scan": {"what": "one query over %d keyframes (full store mode, rotated-query cache on), ncu launch list with DRAM bytes" % n_cand,
                   "launches": len(q), "total_us": round(s_tot, 1), "shares": s_sh, "dram_bytes_per_candidate": round(dram / n_cand),
                   "algorithmic_bytes_per_candidate": 2620160, "full_capture": cap_s}}
    colcol = [m for k, m in cap_t.items() if k.startswith("colcol_kernel")]
    if colcol:
        js["tracking"]["dominant_kernel_traffic_bytes_per_launch"] = round(sum((m["dram_read_MB"] + m["dram_write_MB"]) * m["captures"] for m in colcol) /
                                                                             sum(m["captures"] for m in colcol) * 1e6)
        js["tracking"]["dominant_kernel_images_per_launch"] = 56
    json.dump(js, open(os.path.join(ROOT, "profiles", "ncu_%s_summary.json" % tag), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", "ncu_%s_summary.md" % tag), "w") as fh:
        fh.write("# ncu summary %s (B200, `tools/profile.sh %s`, `--clock-control none`; raw pages: `profiles/ncu_raw_%s_*.csv.gz`)\n\n" % (tag, tag, tag))
        fh.write("Times under ncu are serialised and cold-cache: use the SHARES, not the absolutes (bench.py prints the live per-kernel times).\n\n")
        fh.write("## Tracking step (1000 frames, %d launches, %.1f ms under ncu)\n\n| kernel | launches | total µs | avg µs | share |\n|---|---|---|---|---|\n" % (len(track), t_tot / 1e3))
        for k, v in t_sh.items():
            fh.write("| `%s` | %d | %.1f | %.2f | %.1f %% |\n" % (k, v["launches"], v["total_us"], v["avg_us"], 100 * v["share"]))
        fh.write("\n### `--set full` capture of two consecutive batches (56 frames each; mean over the captures of each instantiation)\n\n")
        table(cap_t, fh)
        fh.write("\n## Loop-closure scan (one query over %d keyframes, %d launches, %.1f ms under ncu)\n\n" % (n_cand, len(q), s_tot / 1e3))
        fh.write("DRAM traffic of the whole query: %.1f MB = **%.2f MB per candidate** (algorithmic 2.62 MB: the candidate's F and P; the full "
                 "store mode also reads its cached Ht, Hp = 5.24 MB; two hypotheses per candidate).\n\n" % (dram / 1e6, dram / n_cand / 1e6))
        fh.write("| kernel | launches | total µs | avg µs | share |\n|---|---|---|---|---|\n")
        for k, v in s_sh.items():
            fh.write("| `%s` | %d | %.1f | %.2f | %.1f %% |\n" % (k, v["launches"], v["total_us"], v["avg_us"], 100 * v["share"]))
        fh.write("\n### `--set full` capture of three scan batches (56 candidates x 2 hypotheses each)\n\n")
        table(cap_s, fh)
    print("tracking", len(track), "launches", round(t_tot), "us; scan", len(q), "launches", round(s_tot), "us, dram/cand", round(dram / n_cand))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
