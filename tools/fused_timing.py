"""Phase times inside k_posterior_fused from clock64() stamps (experiment build only).
    ICPCUDA_LIB_TAG=timing ICPCUDA_NVCC_EXTRA=-DICP_FUSED_TIMING python icp-proposal_b200/build.py   # libicpcuda_timing.so
    ICPCUDA_LIB_TAG=timing python tools/fused_timing.py [--direction 0|1]   # 0 = model sampling (1 row / obs), 1 = target sampling
Per CTA (= chain): consumer warp 0's time blocked on FULL, producers' time blocked on EMPTY and on their own cp.async
group, end of the rank-update phase, start and end of the factorisation; all in SM clocks.
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
prop = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, a.direction, True, ids, tp)
th = bench.init_thetas(m, a.chains)
prop.posterior(th, want_M=False)
prop.clear_cache() if hasattr(prop, "clear_cache") else None
th2 = th.copy(); th2[:, 10:] += 1e-3
prop.posterior(th2, want_M=False)
lib = _lib.load()
n = min(a.chains, 8192)
ST = 12
buf = np.zeros(ST * n, np.int64)
lib.icp_debug_fused_timing.argtypes = [C.c_void_p, C.c_int]
rc = lib.icp_debug_fused_timing(buf.ctypes.data, ST * n)
assert rc == 0, rc
t = buf.reshape(n, ST)
if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
    np.save(os.path.join(ROOT, "gpurun_out", f"fused_timing_dir{a.direction}.npy"), t)
smid = t[:, 7] & 0xFF
tot = t[:, 7] >> 8
start = t[:, 0] - t[:, 0].min()
names = ["consumer blocked on FULL", "consumer build end", "producer blocked on EMPTY", "producer blocked on cp.async",
         "producer build end", "factorisation start", "CTA end"]
cols = [t[:, 1], t[:, 2], t[:, 3], t[:, 4], t[:, 5], t[:, 6], tot]
print(f"direction {a.direction}: {n} CTAs, median clocks (p10 .. p90)")
for nm, v in zip(names, cols):
    print(f"  {nm:32s} {np.median(v):9.0f}  ({np.percentile(v, 10):.0f} .. {np.percentile(v, 90):.0f})")
if os.environ.get("ICPCUDA_FUSE") == "1":
    print(f"  factorisation + store             {np.median(tot - t[:, 6]):9.0f}  (factor {np.median(t[:, 8]):.0f}, back substitution "
          f"{np.median(t[:, 9]):.0f}, store L / mu {np.median(tot - t[:, 6] - t[:, 8] - t[:, 9]):.0f})")
else:
    print(f"  k_cholesky_packed: load {np.median(t[:, 8]):.0f}, factor {np.median(t[:, 9]):.0f}, back substitution {np.median(t[:, 10]):.0f}, "
          f"store L / mu {np.median(t[:, 11]):.0f}")
first = start < np.percentile(start, 12)
print(f"  first-wave CTAs: total {np.median(tot[first]):.0f}; later waves: {np.median(tot[~first]):.0f}")
print(f"  span of the launch {start.max() + tot[start.argmax()]:.0f} clocks; CTAs per SM {np.bincount(smid).max()}")
