"""Throughput of the other BASELINE.json configs, batched on one GPU: config 3 (n_icp = n_eval = N, symmetric
evaluator), config 4 (config-1 mixture, Hausdorff evaluator), config 5 (face-sized open surface, collective evaluator, pose
proposals) - MCMC samples/s through icp_chain_run (device time from icp_chain_last_run_stats) and the per-stage times."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_proposal_b200 import _lib, core, synth  # noqa: E402

which = os.environ.get("CONFIGS", "3,4,5").split(",")
ctx = core.Context(0)
out = {}


def run(name, model, tgt, comps, ev, th0, steps):
    C = len(th0)
    chain = core.Chain(model, tgt, comps, ev, max_chains=C)
    chain.run(th0, 3, seed=5, log_theta=False)
    if C <= 148:      # latency-bound batches: a longer run (the adaptive look-ahead times its widths during the first ~100 steps)
        steps = int(os.environ.get("SMALL_STEPS", "400"))
        chain.run(th0, 150, seed=6, log_theta=False)
    r = chain.run(th0, steps, seed=6, log_theta=False)
    ms, launches = chain.last_run_stats()
    prof = chain.profile(th0, 4, seed=7)
    out[name] = {"chains": C, "steps": steps, "ms_per_step": ms / steps, "samples_per_s": C * steps / (ms * 1e-3),
                 "accept_rate": float(r["n_accepted"].mean() / steps), "launches": launches, "rounds": chain.last_run_rounds(),
                 "stage_ms_per_step": {k: round(v["ms"] / 4, 4) for k, v in prof.items() if v["launches"]}}
    print(name, json.dumps(out[name]), flush=True)
    chain.close()


if "3" in which or "4" in which:
    m, tv, tc, ids, eids, tp = bench.workload()
    K, N = bench.K_RANK, len(m["ref"])
    model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
    tgt = core.Target(ctx, tv, tc)
    if "3" in which:
        C = int(os.environ.get("C3", "592"))
        all_ids = np.arange(N, dtype=np.int32)
        ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, _lib.SYMMETRIC, True, 0.0, 2.0, 0.0, all_ids, tv)
        th0 = np.stack([model.theta(np.random.default_rng(s).normal(0, 0.3, K)) for s in range(C)])
        for ru, nm in ((_lib.RANK_UPDATE_FP64, "fp64"), (_lib.RANK_UPDATE_INT8, "int8")):
            gp = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, all_ids, tv, rank_update=ru)
            comps = [dict(kind=_lib.PROP_ICP, weight=0.9, proposal=gp), dict(kind=_lib.PROP_RANDOM_SHAPE, weight=0.1, sd=0.1)]
            run("config3_n_icp_N_symmetric_" + nm, model, tgt, comps, ev, th0, 10)
    if "4" in which:
        C = int(os.environ.get("C4", "2368"))
        for ru, nm in ((_lib.RANK_UPDATE_INT8, "int8"),):
            pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp, rank_update=ru)
            pm = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp, rank_update=ru)
            comps = [dict(kind=_lib.PROP_ICP, weight=0.45, proposal=pt), dict(kind=_lib.PROP_ICP, weight=0.45, proposal=pm),
                     dict(kind=_lib.PROP_RANDOM_SHAPE, weight=0.1, sd=0.1)]
            ev = core.Evaluator(model, tgt, _lib.EVAL_HAUSDORFF, 0, True, 100.0)
            run("config4_hausdorff_" + nm, model, tgt, comps, ev, bench.init_thetas(m, C), 10)
if "5" in which:
    m = synth.face_twin(rank=100, side=169)
    tv, tc, _ = synth.partial_target(m, seed=7, alpha_sd=0.5)
    K, N = 100, len(m["ref"])
    model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
    tgt = core.Target(ctx, tv, tc)
    rng = np.random.default_rng(5)
    ids = np.sort(rng.choice(N, 500, replace=False)).astype(np.int32)
    tp = tv[np.sort(rng.choice(len(tv), 500, replace=False))]
    gp = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp)
    gt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp)
    pose = [dict(kind=2, weight=0.05, sd=0.01, axis=a) for a in range(3)] + [dict(kind=3, weight=0.05, sd=0.1, axis=a) for a in range(3)]
    comps = [dict(kind=0, weight=0.35, proposal=gt), dict(kind=0, weight=0.35, proposal=gp), dict(kind=1, weight=0.1, sd=0.05)] + pose
    ev = core.Evaluator(model, tgt, _lib.EVAL_COLLECTIVE, _lib.SYMMETRIC, True, 0.1, 0.3, 1.0, ids, tp)
    C = int(os.environ.get("C5", "148"))
    th0 = np.zeros((C, K + 10)); th0[:, 0] = 1.0; th0[:, 7:10] = m["ref"].mean(0)
    th0[:, 10:] = np.random.default_rng(9).normal(0, 0.3, (C, K))
    run("config5_face_sized_collective_pose", model, tgt, comps, ev, th0, 10)
    # closest-point queries on the face-sized target (1256 B / query algorithmic)
    import ctypes as Cc
    import torch
    dev = torch.device("cuda", 0)
    nq = 1_000_000
    for name, q in (("near", synth.near_surface_queries(tv, tc, nq, seed=11)), ("far", synth.far_field_queries(tv, nq, seed=12))):
        qd = torch.from_numpy(np.ascontiguousarray(q)).to(dev)
        tri = torch.empty(nq, dtype=torch.int32, device=dev); cp = torch.empty((nq, 3), dtype=torch.float64, device=dev)
        d2 = torch.empty(nq, dtype=torch.float64, device=dev)
        ms = Cc.c_double(0)
        _lib.check(tgt.lib.icp_debug_time_closest_point(tgt.h, nq, qd.data_ptr(), tri.data_ptr(), cp.data_ptr(), d2.data_ptr(), 10, Cc.byref(ms)), ctx.h)
        out["closest_point_face_" + name] = {"queries_per_s": nq / (ms.value * 1e-3), "ms_per_launch": ms.value, "algorithmic_GBps": nq * 1256 / ms.value / 1e6}
        print(name, json.dumps(out["closest_point_face_" + name]), flush=True)
print(json.dumps(out))
