"""ncu report -> executed warp instructions per CUDA source line (needs -lineinfo and --import-source on):
python profiles/source_lines.py REP [top_n] [units]   (units = e.g. chain steps / 32 lanes, to print instructions per unit)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
units = float(sys.argv[3]) if len(sys.argv) > 3 else None
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout.splitlines()))
fname, ie, out, total = "?", None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        ie = r.index("Instructions Executed")
    elif r[0] and r[0].isdigit() and ie is not None and len(r) > ie:
        try:
            n = int(r[ie])
        except ValueError:
            continue
        out.append((n, fname, int(r[0]), r[1].strip()))
        total += n
print(f"# total executed warp instructions attributed to source lines: {total}" + (f" ({total / units:.1f} per unit)" if units else ""))
for n, f, l, s in sorted(out, reverse=True)[:top]:
    print(f"{100.0 * n / total:5.1f}%" + (f" {n / units:7.2f}/unit" if units else "") + f"  {f}:{l}: {s[:120]}")
