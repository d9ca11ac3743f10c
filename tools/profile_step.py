"""Runs a few MH steps of the bench workload (nothing else), for `ncu` captures.
    ncu ... python tools/profile_step.py --steps 3 --chains 1184
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_proposal_b200 import _lib, core, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--chains", type=int, default=1184)
ap.add_argument("--queries", type=int, default=0, help="also run the closest-point primitive on this many near-surface queries")
ap.add_argument("--rank-update", default="int8", choices=["int8", "fp64"])
a = ap.parse_args()
m, tv, tc, ids, eids, tp = bench.workload()
ctx = core.Context(0)
model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
tgt = core.Target(ctx, tv, tc)
ru = _lib.RANK_UPDATE_INT8 if a.rank_update == "int8" else _lib.RANK_UPDATE_FP64
pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp, rank_update=ru)
pm = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp, rank_update=ru)
comps = [dict(kind=0, weight=0.45, proposal=pt), dict(kind=0, weight=0.45, proposal=pm), dict(kind=1, weight=0.1, sd=0.1)]
ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 0, True, 0.0, 2.0, 0.0, eids, tp)
chain = core.Chain(model, tgt, comps, ev, max_chains=a.chains)
th0 = bench.init_thetas(m, a.chains)
out = chain.run(th0, a.steps, seed=1024, log_theta=False)
print("accept rate", out["accepted"].mean(), "device ms", chain.last_run_stats())
if a.queries:
    q = synth.near_surface_queries(tv, tc, a.queries, seed=11)
    tgt.closest_point_surface(q)
