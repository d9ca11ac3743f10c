"""Checks the fused runner's accept decisions against the per-call entry points (independent code path for the transition densities):
for every step of a chain with caller-supplied randomness, a = v(theta') - v(theta) - (lf - lb) with lf / lb the log-sum-exp over the
mixture of icp_log_transition / random-walk densities, and the decision u_acc < exp(a) - compared with the chain log."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp_proposal_b200 import _lib, core, synth  # noqa: E402

K = int(os.environ.get("RANK", "31"))
m = synth.femur_twin(rank=K)
tv, tc, _ = synth.synthetic_target(m)
ctx = core.Context(0)
model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
tgt = core.Target(ctx, tv, tc)
ids, eids, tp = np.arange(2 * K), np.arange(4 * K), tv[:: max(1, len(tv) // (2 * K))][:2 * K]
ru = _lib.RANK_UPDATE_INT8 if os.environ.get("RU") == "int8" else _lib.RANK_UPDATE_FP64
fac = _lib.FACTOR_SVD if os.environ.get("FACTOR") == "svd" else _lib.FACTOR_CHOLESKY
p0 = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp, factor=fac, rank_update=ru)
p1 = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp, factor=fac, rank_update=ru)
# ICP_ONLY=1: no random-walk component (its density dominates the mixture's at these step sizes and hides the ICP terms)
icp_only = os.environ.get("ICP_ONLY", "1") == "1"
w = np.array([0.5, 0.5, 0.0]) if icp_only else np.array([0.45, 0.45, 0.1]); sd = 0.1
comps = [dict(kind=_lib.PROP_ICP, weight=w[0], proposal=p0), dict(kind=_lib.PROP_ICP, weight=w[1], proposal=p1)]
if not icp_only:
    comps.append(dict(kind=_lib.PROP_RANDOM_SHAPE, weight=w[2], sd=sd))
ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, _lib.MODEL_TO_TARGET, True, 0.0, 2.0, 0.0, eids, tp)
chain = core.Chain(model, tgt, comps, ev, max_chains=1)
chain.set_lookahead(int(os.environ.get("LOOKAHEAD", "0")))
rng = np.random.default_rng(3)
n = int(os.environ.get("STEPS", "300"))
th0 = model.theta(rng.normal(0, 0.5, K))[None]
u_comp, u_acc, z = rng.random((n, 1)), rng.random((n, 1)), rng.normal(size=(n, 1, K))
got = chain.run(th0, n, u_comp=u_comp, z=z, u_acc=u_acc)


def lse(ls):
    ls = np.array(ls); mx = ls.max()
    return -np.inf if not np.isfinite(mx) else mx + np.log(np.sum(w[:len(ls)] * np.exp(ls - mx)))


cur = th0.copy(); vcur = ev.log_value(cur)[0, 0]
bad = 0; worst = 0.0; per_comp = {0: [0, 0], 1: [0, 0], 2: [0, 0]}
for s in range(n):
    ci = int(got["component"][s, 0])
    # the proposal the chain made: the logged state if accepted; otherwise regenerate it through the per-call API
    if ci == 0: prop = p0.propose(cur, z[s])
    elif ci == 1: prop = p1.propose(cur, z[s])
    else:
        prop = cur.copy(); prop[0, 10:] += sd * z[s, 0]
    vprop = ev.log_value(prop)[0, 0]
    d2 = float(np.sum((prop[0, 10:] - cur[0, 10:]) ** 2))
    rw = -0.5 * (K * np.log(2 * np.pi) + K * np.log(sd * sd) + d2 / (sd * sd))
    f0, f1, b0, b1 = p0.log_transition(cur, prop)[0], p1.log_transition(cur, prop)[0], p0.log_transition(prop, cur)[0], p1.log_transition(prop, cur)[0]
    lf = lse([f0, f1] + ([] if icp_only else [rw]))
    lb = lse([b0, b1] + ([] if icp_only else [rw]))
    if s < 6:
        print(f"step {s} comp {ci}: fwd {f0:.3f} {f1:.3f} bwd {b0:.3f} {b1:.3f} rw {rw:.3f} a {vprop - vcur - (lf - lb):.3f}")
    a = vprop - vcur - (lf - lb)
    ok = (a > 0) or (u_acc[s, 0] < np.exp(a))
    per_comp[ci][1] += 1
    if ok != bool(got["accepted"][s, 0]):
        bad += 1; per_comp[ci][0] += 1
        print(f"step {s}: component {ci}: per-call decision {ok} (a = {a:.4f}, u = {u_acc[s, 0]:.4f}), chain {bool(got['accepted'][s, 0])}")
    if got["accepted"][s, 0]:
        np.testing.assert_allclose(got["theta"][s, 0], prop[0], rtol=0, atol=1e-9)
        cur = got["theta"][s:s + 1, 0].copy(); vcur = got["values"][s, 0, 0]
print("steps", n, "decisions that differ", bad, "by component [differ, total]", per_comp)
