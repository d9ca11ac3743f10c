"""Long-run check of the adaptive rejection look-ahead: the chain log of N steps with the automatic width (the active width changes
many times along the run) must equal the step-by-step runner's, bit for bit (config-1 workload, device-resident logs)."""
import os
import sys

import numpy as np
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
pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp, rank_update=_lib.RANK_UPDATE_INT8)
pm = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp, rank_update=_lib.RANK_UPDATE_INT8)
comps = [dict(kind=_lib.PROP_ICP, weight=0.45, proposal=pt), dict(kind=_lib.PROP_ICP, weight=0.45, proposal=pm),
         dict(kind=_lib.PROP_RANDOM_SHAPE, weight=0.1, sd=0.1)]
ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, _lib.MODEL_TO_TARGET, True, 0.0, 2.0, 0.0, eids, tp)
L = bench.K_RANK + 10
steps = int(os.environ.get("STEPS", "5000"))
for C in (1, 5, 40):
    th0 = torch.from_numpy(bench.init_thetas(m, C + 1)[1:].copy()).to(dev)
    logs = {}
    for width in (0, -1):
        chain = core.Chain(model, tgt, comps, ev, max_chains=C)
        chain.set_lookahead(width)
        comp = torch.zeros((steps, C), dtype=torch.int32, device=dev); acc = torch.zeros((steps, C), dtype=torch.uint8, device=dev)
        val = torch.zeros((steps, C, 3), dtype=torch.float64, device=dev); th = torch.zeros((steps, C, L), dtype=torch.float64, device=dev)
        fin = torch.zeros((C, L), dtype=torch.float64, device=dev); nacc = torch.zeros(C, dtype=torch.int64, device=dev)
        half = steps // 2     # two calls: the second resumes the resident state
        for k, (n, first) in enumerate(((half, True), (steps - half, False))):
            o = 0 if first else half
            chain.run_device(C, n, th0.data_ptr() if first else None, seed=4242, log_component=comp[o:].data_ptr(), log_accepted=acc[o:].data_ptr(),
                             log_values=val[o:].data_ptr(), log_theta=th[o:].data_ptr(), theta_final=fin.data_ptr(), n_accepted=nacc.data_ptr())
        ms, _ = chain.last_run_stats()
        logs[width] = [x.cpu().numpy() for x in (comp, acc, val, th, fin, nacc)]
        print(f"C {C} width {width}: last call {ms / (steps - half) :.4f} ms per step, {chain.last_run_rounds()} rounds for {steps - half} steps, accepted {nacc.cpu().numpy().tolist()[:5]}", flush=True)
        chain.close()
    same = all(np.array_equal(a, b) for a, b in zip(logs[0], logs[-1]))
    print(f"C {C}: logs of {steps} steps identical: {same}", flush=True)
    assert same
print("ok")
