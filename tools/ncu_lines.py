"""Stall samples of a kernel aggregated per CUDA source line (needs -lineinfo and --import-source on).
    python tools/ncu_lines.py prof.ncu-rep kernel-regex [top]
"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fn, hdr, acc = None, None, {}
def flush():
    if not acc:
        return
    tot = sum(v[0] for v in acc.values())
    print("=====", fn[:110], "samples", tot)
    for (f, ln), v in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(v[2].items(), key=lambda kv: -kv[1])[:3]
        print(f"  {v[0] / tot * 100:5.1f}%  {ln:>5}  {v[1][:70]:70s} {' '.join(f'{k}={n}' for k, n in st if n)}")
cur_file = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        if fn != r[1]:
            flush(); acc = {}
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        stall = [(i, h.replace("stall_", "")) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    elif hdr and r[0].isdigit():
        try:
            n = int(r[si])
        except ValueError:
            continue
        if n:
            a = acc.setdefault((cur_file, r[0]), [0, r[1].strip(), {}])
            a[0] += n
            for i, h in stall:
                try:
                    a[2][h] = a[2].get(h, 0) + int(r[i])
                except ValueError:
                    pass
flush()
