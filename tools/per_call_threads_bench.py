"""Compiles examples/per_call_threads.cpp and runs it on the bench workload (femur GPMM-100 twin): MH steps/s of the per-call C ABI
from 1 / 2 / 4 / 10 / 16 std::threads on shared handles (no GIL: what JVM threads would see)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from test_gpu_cpp_host import write_model_bin  # noqa: E402

m, tv, tc, ids, eids, tp = bench.workload()
mm = dict(m, target=tv, target_cells=tc)
d = tempfile.mkdtemp()
path = os.path.join(d, "model.bin")
write_model_bin(path, mm, bench.K_RANK, ids, eids, tp)
exe = os.path.join(d, "per_call_threads")
libdir = os.path.join(ROOT, "icp-proposal_b200")
subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "per_call_threads.cpp"),
                       "-o", exe, "-L", libdir, "-licpcuda", f"-Wl,-rpath,{libdir}"])
for t in (1, 2, 4, 10, 16):
    out = subprocess.run([exe, path, str(t), os.environ.get("STEPS", "150")], capture_output=True, text=True, timeout=600)
    print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr, flush=True)
