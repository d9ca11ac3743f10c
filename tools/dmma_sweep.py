"""DMMA (mma.sync.m8n8k4.f64) throughput vs resident warps per SM and independent accumulators per warp."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp_proposal_b200 import _lib, core  # noqa: E402

ctx = core.Context(0)
print("warps/cta ctas/sm nacc -> TFLOP/s")
for wpc, cps in ((4, 1), (4, 2), (8, 1), (4, 3), (4, 4), (8, 2), (16, 1), (8, 4)):
    row = []
    for nacc in (1, 2, 4, 8):
        v = C.c_double(0)
        _lib.check(ctx.lib.icp_debug_dmma_sweep(ctx.h, wpc, cps, nacc, C.byref(v)), ctx.h)
        row.append(f"{v.value:6.2f}")
    print(f"{wpc:2d} x {cps} = {wpc * cps:2d} warps/SM : nacc 1/2/4/8 -> {' '.join(row)}")
