"""Per-kernel device times of one MH step with a single resident chain (BASELINE configs[0] shape): where the launch- and
latency-bound 0.19 ms of a C = 1 step goes. Eager pass with CUDA events around every stage (icp_chain_profile)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_proposal_b200 import _lib, core  # noqa: E402

m, tv, tc, ids, eids, tp = bench.workload()
ctx = core.Context(0)
model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
tgt = core.Target(ctx, tv, tc)
out = {}
for mode, ru in (("fp64", _lib.RANK_UPDATE_FP64), ("int8", _lib.RANK_UPDATE_INT8)):
    pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp, rank_update=ru)
    pm = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp, rank_update=ru)
    comps = [dict(kind=_lib.PROP_ICP, weight=0.45, proposal=pt), dict(kind=_lib.PROP_ICP, weight=0.45, proposal=pm),
             dict(kind=_lib.PROP_RANDOM_SHAPE, weight=0.1, sd=0.1)]
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, _lib.MODEL_TO_TARGET, True, 0.0, 2.0, 0.0, eids, tp)
    for C in (1, 8, 148):
        chain = core.Chain(model, tgt, comps, ev, max_chains=C)
        th0 = bench.init_thetas(m, C)
        prof = chain.profile(th0, 8, seed=1)
        chain.run(th0, 20)
        import time
        t0 = time.perf_counter(); chain.run(th0, 200); dt = time.perf_counter() - t0
        out[f"{mode}_C{C}"] = {"ms_per_step_by_stage": {k: round(v["ms"] / 8, 4) for k, v in prof.items() if v["launches"]},
                               "eager_sum_ms": round(sum(v["ms"] for v in prof.values()) / 8, 4), "host_run_ms_per_step": round(dt / 200 * 1e3, 4)}
        chain.close()
print(json.dumps(out, indent=1))
