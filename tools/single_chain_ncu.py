"""One resident chain, a few MH steps (eager): run under `ncu --metrics gpu__time_duration.sum` for the per-kernel durations at C = 1."""
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
ru = _lib.RANK_UPDATE_INT8 if os.environ.get("RU", "int8") == "int8" else _lib.RANK_UPDATE_FP64
pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp, rank_update=ru)
pm = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp, rank_update=ru)
comps = [dict(kind=_lib.PROP_ICP, weight=0.45, proposal=pt), dict(kind=_lib.PROP_ICP, weight=0.45, proposal=pm),
         dict(kind=_lib.PROP_RANDOM_SHAPE, weight=0.1, sd=0.1)]
ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, _lib.MODEL_TO_TARGET, True, 0.0, 2.0, 0.0, eids, tp)
C = int(os.environ.get("CHAINS", "1"))
chain = core.Chain(model, tgt, comps, ev, max_chains=C)
th0 = bench.init_thetas(m, C)
chain.profile(th0, 6, seed=1)     # eager pass, kernel by kernel
if os.environ.get("LOOKAHEAD"):      # a few look-ahead rounds as well (graph launches: ncu still lists every kernel)
    import torch
    chain.set_lookahead(int(os.environ["LOOKAHEAD"]))
    t0 = torch.from_numpy(th0).to(torch.device("cuda", 0))
    chain.run_device(C, 30, t0.data_ptr(), seed=1)
