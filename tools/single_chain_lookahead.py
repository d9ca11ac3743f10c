"""One chain alone on the GPU (BASELINE configs[0] shape): steps/s of the fused runner against the width of the rejection
look-ahead (icp_chain_set_lookahead), device time of icp_chain_run_device, after burn-in."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_proposal_b200 import _lib, core  # noqa: E402

m, tv, tc, ids, eids, tp = bench.workload()
ctx = core.Context(0)
model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
tgt = core.Target(ctx, tv, tc)
dev = torch.device("cuda", 0)
out = {}
n_chains = int(os.environ.get("CHAINS", "1"))
steps = int(os.environ.get("STEPS", "3000"))
for ru, nm in ((_lib.RANK_UPDATE_INT8, "int8"), (_lib.RANK_UPDATE_FP64, "fp64")):
    pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp, rank_update=ru)
    pm = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp, rank_update=ru)
    comps = [dict(kind=_lib.PROP_ICP, weight=0.45, proposal=pt), dict(kind=_lib.PROP_ICP, weight=0.45, proposal=pm),
             dict(kind=_lib.PROP_RANDOM_SHAPE, weight=0.1, sd=0.1)]
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, _lib.MODEL_TO_TARGET, True, 0.0, 2.0, 0.0, eids, tp)
    th0 = torch.from_numpy(bench.init_thetas(m, max(n_chains, 2))[1:1 + n_chains].copy()).to(dev)   # a random-init chain
    for W in (0, 2, 4, 8, 16, 32):
        chain = core.Chain(model, tgt, comps, ev, max_chains=n_chains)
        chain.set_lookahead(W)
        nacc = torch.zeros(n_chains, dtype=torch.int64, device=dev)
        chain.run_device(n_chains, 500, th0.data_ptr(), seed=1024)            # burn-in (acceptance is high at first)
        chain.run_device(n_chains, steps, None, seed=1024, n_accepted=nacc.data_ptr())
        ms, _ = chain.last_run_stats()
        rounds = chain.last_run_rounds()
        out[f"{nm}_W{W}"] = {"steps_per_s": steps / (ms * 1e-3), "ms_per_step": ms / steps, "rounds": rounds, "steps_per_round": steps / rounds,
                             "ms_per_round": ms / rounds}
        print(nm, "W", W, json.dumps(out[f"{nm}_W{W}"]), "accepted", nacc.cpu().numpy().tolist(), flush=True)
        chain.close()
print(json.dumps(out))
