"""Target-sampling ICP with many target points (config-3-like: n >= N): per-stage device times of the MH step with and
without merging the observations that share their closest model vertex (ICPCUDA_NO_GROUPING=1 switches it off).
    python tools/grouping_experiment.py --points 2000 --chains 1184
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=2000)
ap.add_argument("--chains", type=int, default=1184)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--child", action="store_true")
a = ap.parse_args()

if not a.child:
    for flag in ("0", "1"):
        env = dict(os.environ, ICPCUDA_NO_GROUPING=flag)
        out = subprocess.run([sys.executable, __file__, "--child", "--points", str(a.points), "--chains", str(a.chains),
                              "--steps", str(a.steps)], env=env, capture_output=True, text=True)
        print("grouping", "off" if flag == "1" else "on ", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:])
    sys.exit(0)

import numpy as np  # noqa: E402

import bench  # noqa: E402
from icp_proposal_b200 import _lib, core  # noqa: E402

m, tv, tc, ids, eids, tp = bench.workload()
rng = np.random.default_rng(5)
pick = rng.integers(0, len(tv), a.points)
tp = tv[pick] + rng.normal(0, 0.05, (a.points, 3))
ctx = core.Context(0)
model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
tgt = core.Target(ctx, tv, tc)
pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp)
comps = [dict(kind=0, weight=0.9, proposal=pt), dict(kind=1, weight=0.1, sd=0.1)]
ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 0, True, 0.0, 2.0, 0.0, eids, tp[: len(eids)])
chain = core.Chain(model, tgt, comps, ev, max_chains=a.chains)
th0 = bench.init_thetas(m, a.chains)
chain.run(th0, 2, seed=3, log_theta=False)
prof = chain.profile(th0, a.steps, seed=3)
print(json.dumps({k: round(v["ms"] / a.steps, 4) for k, v in prof.items() if v["launches"]}))
