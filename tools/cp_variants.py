"""Closest-point traversal variants side by side on the bench workload's target: the per-thread binary walk (ICPCUDA_WIDE=0)
and the warp-cooperative L-ary walk (4 and 8 lanes per query). Checks that every variant returns the same answers.
    python tools/cp_variants.py [nq]
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from icp_proposal_b200 import _lib, core, synth  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m, tv, tc, ids, eids, tp = bench.workload()
ctx = core.Context(0)
dev = torch.device("cuda", 0)
out = {"l2_read_gbs": ctx.l2_bandwidth(32 << 20)}
sets = {"near_surface": synth.near_surface_queries(tv, tc, nq, seed=11), "far_field": synth.far_field_queries(tv, nq, seed=12)}
ref = {}
for variant in ("0", "4", "8"):
    os.environ["ICPCUDA_WIDE"] = variant
    tgt = core.Target(ctx, tv, tc)
    for name, q in sets.items():
        qd = torch.from_numpy(q).to(dev)
        tri = torch.empty(nq, dtype=torch.int32, device=dev); cp = torch.empty((nq, 3), dtype=torch.float64, device=dev)
        d2 = torch.empty(nq, dtype=torch.float64, device=dev)
        ms = C.c_double(0)
        _lib.check(ctx.lib.icp_debug_time_closest_point(tgt.h, nq, qd.data_ptr(), tri.data_ptr(), cp.data_ptr(), d2.data_ptr(), 10, C.byref(ms)), ctx.h)
        torch.cuda.synchronize()
        res = (tri.cpu().numpy(), d2.cpu().numpy(), cp.cpu().numpy())
        if variant == "0":
            ref[name] = res
        same = all(np.array_equal(a, b) for a, b in zip(res, ref[name]))
        out[f"wide{variant}_{name}"] = {"ms": ms.value, "Gq_per_s": nq / ms.value / 1e6, "GBps_1000B": nq * 1000.0 / (ms.value * 1e-3) / 1e9,
                                        "identical_to_binary": bool(same)}
    tgt.close()
print(json.dumps(out, indent=1))
