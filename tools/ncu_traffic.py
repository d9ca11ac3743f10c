"""Writes profiles/traffic.json - the ncu counters bench.py cannot measure inside a timed run (DRAM bytes per launch,
lanes active per executed instruction) - from an `ncu --set full` capture of tools/profile_step.py.

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep <chains> <tag> [<closest-point .ncu-rep of the 1e6-query launch>]

bench.py labels these values as profile constants and names this file as their source.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    return hdr, rows[2:]


def metric(hdr, units, r, name):
    """value of a raw-page metric in bytes / plain units (ncu prints Mbyte / Kbyte / byte columns)"""
    i = hdr.index(name)
    v = float(r[i].replace(",", "") or 0)
    u = units[i].lower()
    return v * {"mbyte": 1e6, "kbyte": 1e3, "gbyte": 1e9, "byte": 1.0}.get(u, 1.0)


def first(hdr, rows, units, pattern):
    for r in rows:
        if pattern in r[hdr.index("Kernel Name")]:
            return r
    return None


def main():
    rep, chains, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    allrows = list(csv.reader(raw.splitlines()))
    hdr, units, rows = allrows[0], allrows[1], allrows[2:]
    out = {}
    for key, pat in (("rank_update", "k_posterior_fused<24, 13, 4, 3, 0>"), ("cholesky", "k_cholesky_packed"),
                     ("closest_point_in_step", "k_nearest")):
        r = first(hdr, rows, units, pat)
        if r is None:
            continue
        out[key] = {"kernel": r[hdr.index("Kernel Name")].split("(")[0],
                    "dram_bytes_per_launch": metric(hdr, units, r, "dram__bytes_read.sum") + metric(hdr, units, r, "dram__bytes_write.sum"),
                    "chains": chains, "source": f"profiles/{tag} (ncu --set full, C = {chains})"}
    if len(sys.argv) > 4:
        raw = subprocess.run(["ncu", "-i", sys.argv[4], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        allrows = list(csv.reader(raw.splitlines()))
        hdr, units, rows = allrows[0], allrows[1], allrows[2:]
        r = max((x for x in rows if "k_nearest" in x[hdr.index("Kernel Name")]), key=lambda x: metric(hdr, units, x, "gpu__time_duration.sum"))
        out["closest_point"] = {"kernel": r[hdr.index("Kernel Name")].split("(")[0],
                                "dram_bytes_per_launch": metric(hdr, units, r, "dram__bytes_read.sum") + metric(hdr, units, r, "dram__bytes_write.sum"),
                                "lanes_active_per_instruction": metric(hdr, units, r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                                "queries": 1000000, "source": f"profiles/{tag} (ncu --set full, 1e6 near-surface queries)"}
    path = os.path.join(ROOT, "profiles", "traffic.json")
    old = {}
    if os.path.exists(path):
        with open(path) as f:
            old = json.load(f)
    old.update(out)
    with open(path, "w") as f:
        json.dump(old, f, indent=1)
    print(json.dumps(old, indent=1))


if __name__ == "__main__":
    main()
