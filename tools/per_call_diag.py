"""Per-call API timing diagnostics: time of each entry point at C chains per call, step by step."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_proposal_b200 import _lib, core  # noqa: E402
from concurrent.futures import ThreadPoolExecutor

m, tv, tc, ids, eids, tp = bench.workload()
ctx = core.Context(0)
model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
tgt = core.Target(ctx, tv, tc)
ru = _lib.RANK_UPDATE_INT8 if os.environ.get("RU", "int8") == "int8" else _lib.RANK_UPDATE_FP64
pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp, rank_update=ru)
pm = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp, rank_update=ru)
ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, _lib.MODEL_TO_TARGET, True, 0.0, 2.0, 0.0, eids, tp)
th0 = bench.init_thetas(m, 128, 0)
K = bench.K_RANK

def run(Cb, n, tag):
    rng = np.random.default_rng(900 + Cb)
    th = th0[:Cb].copy()
    rows = []
    for k in range(n):
        z = rng.normal(size=(Cb, K))
        t = [time.perf_counter()]
        prop = (pt if rng.random() < 0.5 else pm).propose(th, z); t.append(time.perf_counter())
        for p in (pt, pm):
            p.log_transition(th, prop); t.append(time.perf_counter())
            p.log_transition(prop, th); t.append(time.perf_counter())
        ev.log_value(prop); t.append(time.perf_counter())
        th = prop
        rows.append(np.diff(t) * 1e3)
    rows = np.array(rows)
    print(tag, "C", Cb, "ms per call [propose, pt.fwd, pt.bwd, pm.fwd, pm.bwd, eval]: first", np.round(rows[0], 2), "median", np.round(np.median(rows, 0), 3),
          "max", np.round(rows.max(0), 2), "argmax", rows.argmax(0), "total/step", round(float(rows.sum(1).mean()), 3), flush=True)

run(1, 30, "cold")
run(10, 30, "before-threads")
def chain(t, n):
    rng = np.random.default_rng(500 + t)
    th = th0[t].copy()[None]
    for _ in range(n):
        z = rng.normal(size=(1, K))
        prop = (pt if rng.random() < 0.5 else pm).propose(th, z)
        for p in (pt, pm):
            p.log_transition(th, prop); p.log_transition(prop, th)
        ev.log_value(prop)
        th = prop
for nt in (2, 4, 10, 16):
    with ThreadPoolExecutor(max_workers=nt) as ex:
        list(ex.map(lambda t: chain(t, 5), range(nt)))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=nt) as ex:
        list(ex.map(lambda t: chain(t, 100), range(nt)))
    dt = time.perf_counter() - t0
    print("threads", nt, "steps/s", round(nt * 100 / dt), flush=True)
run(1, 30, "after-threads")
run(10, 60, "after-threads")
run(100, 30, "after-threads")
