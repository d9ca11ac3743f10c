"""Phase times inside k_rank_update_i8 from clock64() stamps (experiment build only).
    ICPCUDA_LIB_TAG=timing ICPCUDA_NVCC_EXTRA=-DICP_I8_TIMING python icp-proposal_b200/build.py   # libicpcuda_timing.so
    ICPCUDA_LIB_TAG=timing python tools/i8_timing.py [--direction 0|1]   # 0 = model sampling (1 row / obs), 1 = target sampling
Per CTA (persistent, ~16 chains): cycles the roles spend blocked on each mbarrier and working; all in SM clocks.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_proposal_b200 import _lib, core  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--direction", type=int, default=0)
ap.add_argument("--chains", type=int, default=2368)
a = ap.parse_args()
m, tv, tc, ids, eids, tp = bench.workload()
ctx = core.Context(0)
model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
tgt = core.Target(ctx, tv, tc)
prop = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, a.direction, True, ids, tp, rank_update=_lib.RANK_UPDATE_INT8)
th = bench.init_thetas(m, a.chains)
prop.posterior(th, want_M=False)
th2 = th.copy(); th2[:, 10:] += 1e-3
prop.posterior(th2, want_M=False)
lib = _lib.load()
buf = np.zeros(256 * 16, np.int64)
lib.icp_debug_i8_timing.argtypes = [C.c_void_p, C.c_int]
rc = lib.icp_debug_i8_timing(buf.ctypes.data, buf.size)
assert rc == 0, rc
t = buf.reshape(256, 16)[:148]
names = ["CTA total", "converter: wait full_raw", "converter: wait empty_dig", "converter: convert + store", "converter: b reduce + barrier",
         "mma: wait full_dig", "mma: issue", "gather: wait empty_raw", "gather: issue", "epilogue: wait acc_done", "epilogue: barrier after the tasks", "chains", "converter: LDS + release raw", "converter: proxy fence + arrive", "epilogue: warp 0 own tasks (fewest)", "converter warp 0: everything else (loop control, b reduction, b output)"]
print(f"direction {a.direction}: median over 148 CTAs (p10 .. p90), SM clocks per CTA; per chain in brackets")
ch = max(1.0, float(np.median(t[:, 11])))
for i, nm in enumerate(names):
    v = t[:, i]
    print(f"  {nm:38s} {np.median(v):10.0f} ({np.percentile(v, 10):.0f} .. {np.percentile(v, 90):.0f})  [{np.median(v) / ch:8.0f}]")
