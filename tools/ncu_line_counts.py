"""Executed-instruction counts per CUDA source line of one captured launch (divergence diagnosis).
    python tools/ncu_line_counts.py prof.ncu-rep kernel-regex launch-skip line,line,...
"""
import csv
import subprocess
import sys

rep, pat, skip = sys.argv[1], sys.argv[2], sys.argv[3]
want = {int(x) for x in sys.argv[4].split(",")} if len(sys.argv) > 4 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
hdr, seen, total = None, set(), 0
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Line No":
        hdr = r
        ie, te = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    elif hdr and r and r[0].isdigit() and int(r[0]) not in seen:
        seen.add(int(r[0]))
        total += int(r[ie])
        if want is None or int(r[0]) in want:
            print(f"{r[0]:>5} {r[1].strip()[:56]:56s} warp-inst {int(r[ie]):>10} thread-inst {int(r[te]):>11} lanes {int(r[te]) / max(int(r[ie]), 1):5.1f}")
print("total warp-inst", total)
