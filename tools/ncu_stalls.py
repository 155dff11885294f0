"""Per-kernel summary + stall breakdown of an ncu --set full report: python tools/ncu_stalls.py gpurun_out/prof.ncu-rep"""
import csv, io, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[0], rows[2:]
H = {h: i for i, h in enumerate(hdr)}
stalls = ['barrier', 'long_scoreboard', 'short_scoreboard', 'mio_throttle', 'lg_throttle', 'math_pipe_throttle', 'wait', 'not_selected', 'dispatch_stall', 'branch_resolving', 'no_instruction']
def f(x):
    try: return float(x.replace(',', ''))
    except Exception: return float('nan')
print("%-52s %6s %4s %5s %5s %5s %5s %5s %6s %7s %7s | " % ("kernel", "us", "regs", "warp%", "iss%", "l1%", "lts%", "dram%", "Minst", "bankcf", "smemwf") + " ".join(s[:6] for s in stalls))
for d in data:
    name = re.sub(r'void |nis::|\(int\)', '', d[H['Kernel Name']])[:52]
    vals = [f(d[H['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s]]) for s in stalls]
    g = lambda k: f(d[H[k]])
    print("%-52s %6.1f %4.0f %5.1f %5.1f %5.1f %5.1f %5.1f %6.2f %7.0f %7.0f | " % (
        name, g('gpu__time_duration.sum'), g('launch__registers_per_thread'), g('sm__warps_active.avg.pct_of_peak_sustained_active'),
        g('smsp__issue_active.avg.pct_of_peak_sustained_active'), g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'),
        g('lts__throughput.avg.pct_of_peak_sustained_elapsed'), g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), g('smsp__inst_executed.sum') / 1e6,
        g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'), g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')) + " ".join("%6.2f" % v for v in vals))
