"""Closest-point traversal experiments: query order (random vs Morton-sorted vs mesh order)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_proposal_b200 import _lib, core, synth  # noqa: E402


def morton_order(q):
    lo, hi = q.min(0), q.max(0)
    g = ((q - lo) / (hi - lo) * 1023).astype(np.uint64)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    return np.argsort((spread(g[:, 0]) << 2) | (spread(g[:, 1]) << 1) | spread(g[:, 2]), kind="stable")


m, tv, tc, ids, eids, tp = bench.workload()
ctx = core.Context(0)
tgt = core.Target(ctx, tv, tc)
dev = torch.device("cuda", 0)
nq = 1_000_000
for name, q in (("near", synth.near_surface_queries(tv, tc, nq, seed=11)), ("far", synth.far_field_queries(tv, nq, seed=12))):
    for order in ("random", "morton"):
        qq = q[morton_order(q)] if order == "morton" else q
        qd = torch.from_numpy(np.ascontiguousarray(qq)).to(dev)
        tri = torch.empty(nq, dtype=torch.int32, device=dev); cp = torch.empty((nq, 3), dtype=torch.float64, device=dev)
        d2 = torch.empty(nq, dtype=torch.float64, device=dev)
        ms = C.c_double(0)
        _lib.check(tgt.lib.icp_debug_time_closest_point(tgt.h, nq, qd.data_ptr(), tri.data_ptr(), cp.data_ptr(), d2.data_ptr(), 10, C.byref(ms)), ctx.h)
        print(f"{name:5s} {order:7s} {ms.value:.3f} ms  {nq / ms.value / 1e6:.2f} Gq/s  {nq * 1000 / ms.value / 1e6:.0f} GB/s algorithmic")
