"""Summarises an .ncu-rep: per-kernel headline metrics + stall-sample breakdown by SASS opcode.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex]
"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_active.avg"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
seen = set()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if pat and not re.search(pat, name):
        continue
    key = name.split("(")[0]
    if key in seen:
        continue
    seen.add(key)
    print("=====", name[:100])
    for w in want:
        if w in hdr:
            print(f"  {w}: {r[hdr.index(w)]}")
    st = sorted(((float(r[hdr.index(s)] or 0), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for s in stall), reverse=True)
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + re.escape(key.split("<")[0].split("::")[-1])],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    try:
        hi = next(i for i, x in enumerate(srows) if x and x[0] == "Address")
    except StopIteration:
        continue
    sh = srows[hi]
    si, ni = sh.index("Source"), sh.index("# Samples")
    agg, tot, top = collections.Counter(), 0, []
    for idx, x in enumerate(srows[hi + 1:]):
        if x and x[0] == "Kernel Name":
            break
        try:
            n = int(x[ni])
        except (ValueError, IndexError):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", x[si])
        agg[m.group(2).split(".")[0] if m else "?"] += n
        tot += n
        top.append((n, idx, x[si][:60]))
    if tot:
        print("  samples by opcode:", ", ".join(f"{k}={v / tot * 100:.1f}%" for k, v in agg.most_common(10)))
        for n, idx, s in sorted(top, reverse=True)[:8]:
            print(f"    {n / tot * 100:5.2f}% @{idx:4d} {s}")
