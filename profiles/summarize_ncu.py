"""ncu report -> short text summary (run here, no GPU needed): python profiles/summarize_ncu.py REP OUT 'title line' [steps]"""
import collections
import csv
import subprocess
import sys

rep, out_path, title = sys.argv[1], sys.argv[2], sys.argv[3]
unit_steps = float(sys.argv[4]) if len(sys.argv) > 4 else None
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, row = raw[0], raw[1], raw[2]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.avg.per_second"]
lines = [f"# {title}", f"# report: {rep.split('/')[-1]} (ncu --set full --clock-control none, 1 launch; per-launch times are cold-cache and serialised)", ""]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        lines.append(f"{k:72s} {row[i]:>18s} {units[i]}")
stalls = [(h, float(row[i])) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
lines += ["", "# warp stall reasons (warps per issue-active cycle, > 0.05)"]
lines += [f"{h:92s} {v:8.3f}" for h, v in sorted(stalls, key=lambda t: -t[1]) if v > 0.05][:12]
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h2 = src[1]
ia, ie = h2.index("Source"), h2.index("Instructions Executed")
ops, total = collections.Counter(), 0
for r in src[2:]:
    try:
        n = int(r[ie])
    except (ValueError, IndexError):
        continue
    t = r[ia].strip().split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += n
    total += n
lines += ["", f"# executed warp instructions by SASS opcode (total {total}" + (f", {total / unit_steps:.1f} per unit of work" if unit_steps else "") + ")"]
lines += [f"{op:10s} {n:14d} {100.0 * n / total:5.1f}%" + (f" {n / unit_steps:8.2f}/unit" if unit_steps else "") for op, n in ops.most_common(16)]
open(out_path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
