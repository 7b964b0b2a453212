"""Summarises an .ncu-rep (raw page + CUDA-source hot lines) into text for profiles/. Usage: ncu_summary.py REP [N]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda k: next((vals[i] + " " + units[i] for i, h in enumerate(hdr) if h == k), "n/a")
print("kernel:", get("Kernel Name"), "grid", get("Grid Size"), "block", get("Block Size"))
for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"):
    print("%-75s %s" % (k, get(k)))
print("stall cycles per issued instruction:")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = float(vals[i].replace(",", ""))
        if v >= 0.05:
            print("   %-28s %.2f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur, data = None, []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) >= 9 and r[2] == "-":
        try:
            data.append((cur, int(r[0]), r[1].strip(), int(r[6] or 0), int(r[7] or 0), int(r[8] or 0)))
        except ValueError:
            pass
tot = sum(d[3] for d in data) or 1
toti = sum(d[4] for d in data) or 1
totti = sum(d[5] for d in data)
print("source lines by stall samples (total %d samples, %.1f threads/instruction overall):" % (tot, totti / toti))
byfile = collections.Counter()
for f, l, s, n, i, ti in data:
    byfile[f] += n
print("   by file: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in byfile.most_common(8)))
for f, l, s, n, i, ti in sorted(data, key=lambda x: -x[3])[:topn]:
    print("   %5.1f%%  inst %4.1f%%  thr/inst %4.1f  %s:%d  %s" % (100.0 * n / tot, 100.0 * i / toti, ti / max(i, 1), f, l, s[:90]))

print("source lines by executed warp-instructions:")
for f, l, s, n, i, ti in sorted(data, key=lambda x: -x[4])[:topn]:
    print("   inst %4.1f%%  thr/inst %4.1f  stall %4.1f%%  %s:%d  %s" % (100.0 * i / toti, ti / max(i, 1), 100.0 * n / tot, f, l, s[:90]))
print("warp-instructions by thread occupancy (threads per instruction, share of all warp-instructions):")
buckets = collections.Counter()
for f, l, s, n, i, ti in data:
    if i:
        buckets[min(int(ti / i) // 4 * 4, 28)] += i
for b in sorted(buckets):
    print("   %2d-%2d threads: %5.1f%%" % (b, b + 3 if b < 28 else 32, 100.0 * buckets[b] / toti))
