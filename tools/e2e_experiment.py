"""Repeated host-buffer icp_chain_run calls: wall-clock variance of the end-to-end path."""
import ctypes as Cc
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_proposal_b200 import _lib, core  # noqa: E402

C, steps = 1184, int(sys.argv[1]) if len(sys.argv) > 1 else 10
m, tv, tc, ids, eids, tp = bench.workload()
ctx = core.Context(0)
model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"]); tgt = core.Target(ctx, tv, tc)
pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, 1, True, ids, tp); pm = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, 0, True, ids, tp)
comps = [dict(kind=0, weight=0.45, proposal=pt), dict(kind=0, weight=0.45, proposal=pm), dict(kind=1, weight=0.1, sd=0.1)]
ev = core.Evaluator(model, tgt, 1, 0, True, 0.0, 2.0, 0.0, eids, tp)
chain = core.Chain(model, tgt, comps, ev, max_chains=C)
L = 111
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
h_th0 = pin((C, L), torch.float64); h_th0[:] = bench.init_thetas(m, C)
io = _lib.ChainIO(); io.seed = 1024
h_comp, h_acc = pin((steps, C), torch.int32), pin((steps, C), torch.uint8)
h_val, h_thl = pin((steps, C, 3), torch.float64), pin((steps, C, L), torch.float64)
h_fin, h_nacc = pin((C, L), torch.float64), pin((C,), torch.int64)
io.log_component, io.log_accepted, io.log_values, io.log_theta = h_comp.ctypes.data, h_acc.ctypes.data, h_val.ctypes.data, h_thl.ctypes.data
io.theta_final, io.n_accepted = h_fin.ctypes.data, h_nacc.ctypes.data
for rep in range(8):
    t0 = time.perf_counter()
    _lib.check(chain.lib.icp_chain_run(chain.h, C, steps, _lib.dptr(h_th0), Cc.byref(io)), ctx.h)
    dt = (time.perf_counter() - t0) * 1e3
    dev_ms, _ = chain.last_run_stats()
    print(f"rep {rep}: wall {dt:.2f} ms ({dt / steps:.3f} ms/step), device {dev_ms:.2f} ms, accept {h_acc.mean():.3f}")
