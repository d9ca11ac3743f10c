"""Per-call (drop-in) API throughput alone: bench.per_call_latency on the bench workload, shared handles from 1 / 10 threads."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_proposal_b200 import _lib, core  # noqa: E402

m, tv, tc, ids, eids, tp = bench.workload()
ctx = core.Context(0)
model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
tgt = core.Target(ctx, tv, tc)
for ru in (_lib.RANK_UPDATE_INT8, _lib.RANK_UPDATE_FP64):
    pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp, rank_update=ru)
    pm = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp, rank_update=ru)
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, _lib.MODEL_TO_TARGET, True, 0.0, 2.0, 0.0, eids, tp)
    th0 = bench.init_thetas(m, 128, 0)
    out = bench.per_call_latency(m, model, pt, pm, ev, th0, iters=int(os.environ.get("ITERS", "100")))
    print(json.dumps({"rank_update": "int8" if ru == _lib.RANK_UPDATE_INT8 else "fp64", **out}))
