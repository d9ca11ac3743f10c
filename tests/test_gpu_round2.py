"""GPU parity tests added in round 2: the reference's SVD covariance factor (value parity of `propose` for z != 0), the
reference-structure chain with random normals, registration measures on an open mesh, the Dice coefficient, per-chain
status of the fused runner, best-sample tracking, thread safety of shared handles and the step-graph cache."""
import threading

import numpy as np
import pytest

from conftest import random_theta
from icp_proposal_b200 import _lib, core, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # BASELINE.json north_star


def _dev(ctx, m):
    return core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"]), core.Target(ctx, m["target"], m["target_cells"])


def _orc(m):
    return orc.Model(m["ref"], m["cells"], m["basis"], m["variance"]), orc.Mesh(m["target"], m["target_cells"])


def _femur(femur, which):
    return dict(ref=femur["ref"], cells=femur["cells"], target=femur["target"], target_cells=femur["target_cells"], **femur[which])


@pytest.mark.parametrize("direction", [_lib.MODEL_SAMPLING, _lib.TARGET_SAMPLING])
@pytest.mark.parametrize("fixture", ["twin31", "femur50", "femur100"])
def test_propose_svd_factor_matches_reference_structure(ctx, request, femur, direction, fixture):
    """NonRigidIcpProposal.scala:53-68 with the posterior sampled through the SVD of D M^-1 D (ICP_FACTOR_SVD): the device
    value of theta' for a caller-supplied z against the oracle in the REFERENCE's structure (rotated basis on all N points
    + full-mesh regression), z != 0."""
    m = request.getfixturevalue("twin31") if fixture == "twin31" else _femur(femur, "gpmm_50" if fixture == "femur50" else "gpmm_100")
    K = len(m["variance"])
    model, tgt = _dev(ctx, m)
    om, ot = _orc(m)
    rng = np.random.default_rng(17)
    C = 3 if K <= 51 else 2
    th = random_theta(m, rng, C, pose=(fixture == "twin31"))
    ids = np.arange(2 * K)
    tp = m["target"][:: max(1, len(m["target"]) // (2 * K))][: 2 * K] + rng.normal(0, 0.1, (2 * K, 3))
    gp = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, direction, True, ids, tp, factor=_lib.FACTOR_SVD)
    op = orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, direction, True, ids, tp)
    z = rng.normal(size=(C, K))
    prop = gp.propose(th, z)
    lt = gp.log_transition(th, prop)
    for c in range(C):
        want = op.propose(th[c], z[c])                       # reference structure, same z
        np.testing.assert_allclose(prop[c], want, rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(lt[c], op.log_transition(th[c], want), rtol=RTOL)
    # the factor is a square root of the posterior covariance and differs from the Cholesky one
    mu, M, _ = gp.posterior(th[:1])
    prop0 = gp.propose(th[:1], np.zeros((1, K)))[0, 10:]
    W = np.stack([gp.propose(th[:1], np.eye(K)[i][None])[0, 10:] - prop0 for i in range(K)], axis=1) / 0.1
    # (W is read off through propose, i.e. as S W with the 1e-5 regulariser's S = I + O(2e-7): absolute slack on that scale)
    Minv = np.linalg.inv(M[0])
    np.testing.assert_allclose(W @ W.T, Minv, rtol=1e-5, atol=1e-5 * np.abs(Minv).max())
    gc = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, direction, True, ids, tp)
    assert np.abs(gc.propose(th, z) - prop).max() > 1e-3
    gc.close(); gp.close(); model.close(); tgt.close()


def test_config1_reference_structure_chain_random_normals(ctx, femur):
    """Config 1 on the reference's femur GPMM-100: the fused runner with ICP_FACTOR_SVD against the oracle chain in the
    reference's own structure (closed_form=False) for 32 steps with random z (SamplingRegistration.scala:60-85)."""
    m = _femur(femur, "gpmm_100")
    K = 101
    model, tgt = _dev(ctx, m)
    om, ot = _orc(m)
    ids, eids = np.arange(2 * K), np.arange(4 * K)
    tp = m["target"][:: len(m["target"]) // (2 * K)][: 2 * K]
    mk = lambda d: core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, d, True, ids, tp, factor=_lib.FACTOR_SVD)
    mo = lambda d: orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, d, True, ids, tp)
    comps = [dict(kind=0, weight=0.45, proposal=mk(1)), dict(kind=0, weight=0.45, proposal=mk(0)), dict(kind=1, weight=0.1, sd=0.1)]
    comps_o = [dict(kind=0, weight=0.45, icp=mo(1)), dict(kind=0, weight=0.45, icp=mo(0)), dict(kind=1, weight=0.1, sd=0.1)]
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 0, True, 0.0, 2.0, 0.0, eids, tp)
    chain = core.Chain(model, tgt, comps, ev, max_chains=2)
    rng = np.random.default_rng(4201)
    n = 32
    th0 = model.theta(rng.normal(0, 0.3, K))[None]
    u_comp, u_acc, z = rng.random((n, 1)), rng.random((n, 1)), rng.normal(size=(n, 1, K))
    got = chain.run(th0, n, u_comp=u_comp, z=z, u_acc=u_acc)
    want = orc.chain_run(om, ot, comps_o, True, orc.EVAL_INDEPENDENT, 0, (0.0, 2.0), eids, tp, th0[0], n, u_comp[:, 0], z[:, 0],
                         u_acc[:, 0], closed_form=False)
    assert np.array_equal(got["component"][:, 0], want["comp"])
    assert np.array_equal(got["accepted"][:, 0], want["accepted"])
    np.testing.assert_allclose(got["values"][:, 0], want["logv"], rtol=RTOL)
    np.testing.assert_allclose(got["theta"][:, 0], want["theta"], rtol=0, atol=1e-5)
    assert want["n_accepted"] >= 3 and (want["comp"][want["accepted"]] < 2).any()     # ICP proposals were accepted
    assert got["status"][0] == 0
    # BestSampleLogger: theta0 and every post-step state compete
    vals = np.concatenate([[np.nan], got["values"][:, 0, 0]])
    best = np.nanargmax(vals)
    np.testing.assert_allclose(got["value_best"][0], np.nanmax(vals))
    if best > 0:
        np.testing.assert_array_equal(got["theta_best"][0], got["theta"][best - 1, 0])
    chain.close(); ev.close(); model.close(); tgt.close()


def test_registration_metrics_and_dice_open_mesh(ctx, open_twin, twin31):
    """RegistrationComparison.scala:24-49 with the boundary filter active (open target), and MeshMetrics.diceCoefficient."""
    m = open_twin
    model, tgt = _dev(ctx, m)
    om, ot = _orc(m)
    assert tgt.boundary_flags().any()
    th = random_theta(m, np.random.default_rng(12), 3, pose=True)
    got = core.registration_metrics(model, tgt, th)
    for c in range(len(th)):
        want = orc.registration_metrics(om, ot, th[c])
        np.testing.assert_allclose(got[c], want, rtol=1e-9)
        assert want[2] != want[0]                    # the filter dropped something
    model.close(); tgt.close()
    # Dice on the closed femur twin (an inside test only means something on a closed surface)
    m = twin31
    model, tgt = _dev(ctx, m)
    om, ot = _orc(m)
    rng = np.random.default_rng(3)
    th = random_theta(m, rng, 3, pose=True)
    th[2, 1] += 15.0                                 # a clearly displaced mesh: lower overlap
    u = rng.random((4000, 3))
    got = core.dice_coefficient(model, tgt, th, unit_samples=u)
    want = np.array([orc.dice_coefficient(om, ot, t, u) for t in th])
    # a sample that is equidistant to two vertices within rounding may flip: allow 2 of 4000
    assert np.all(np.abs(got - want) <= 2.0 * 2 / 4000)
    assert got[2] < got[0] - 0.05 and 0.8 < got[0] <= 1.0
    # device-generated samples: deterministic in the seed, close to the host-sample estimate
    a = core.dice_coefficient(model, tgt, th, n_samples=10000, seed=7)
    b = core.dice_coefficient(model, tgt, th, n_samples=10000, seed=7)
    assert np.array_equal(a, b) and np.all(np.abs(a - got) < 0.05)
    model.close(); tgt.close()


def test_chain_status_empty_set_and_nan(ctx, open_twin):
    """Where the reference throws (CollectiveAverage...Evaluator.scala:51 on an empty filtered list; a NaN log-value), the fused
    runner finishes, writes its outputs and reports per-chain status words plus a non-OK return code."""
    m = open_twin
    K = len(m["variance"])
    model, tgt = _dev(ctx, m)
    ids = np.arange(0, len(m["ref"]), 9)
    tp = m["target"][::11]
    gp = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, 0, True, ids, tp)
    comps = [dict(kind=0, weight=0.5, proposal=gp), dict(kind=1, weight=0.5, sd=0.05)]
    # model -> target with a single model point whose closest target vertex lies on the boundary: every state is empty
    flags = tgt.boundary_flags()
    th0 = random_theta(m, np.random.default_rng(4), 2)
    X = model.reconstruct(th0[:1])[0]
    _, _, cp, _ = tgt.closest_point_surface(X)
    vid, _ = tgt.closest_vertex(cp)
    on_b = np.nonzero(flags[vid])[0]
    assert len(on_b) > 0
    ev = core.Evaluator(model, tgt, _lib.EVAL_COLLECTIVE, 0, True, 0.1, 0.3, 1.0, on_b[:1], tp)
    chain = core.Chain(model, tgt, comps, ev, max_chains=2)
    with pytest.raises(_lib.IcpCudaError) as e:
        chain.run(th0[:1], 5, seed=3)
    assert e.value.code == _lib.ERR_EMPTY_SET and "chain 0" in str(e.value)
    out = chain.run(th0[:1], 5, seed=3, raise_on_status=False)
    assert out["status_code"] == _lib.ERR_EMPTY_SET and out["status"][0] & _lib.CHAIN_EMPTY_SET
    assert not out["accepted"].any() and np.isnan(out["values"][:, 0, 0]).all()     # NaN is propagated, every step rejected
    np.testing.assert_array_equal(out["theta_final"][0], th0[0])
    chain.close(); ev.close()
    # a healthy evaluator: status 0; then a NaN start (non-finite coefficient) flags that chain only
    ev = core.Evaluator(model, tgt, _lib.EVAL_COLLECTIVE, 2, True, 0.1, 0.3, 1.0, ids, tp)
    chain = core.Chain(model, tgt, comps, ev, max_chains=2)
    ok = chain.run(th0, 6, seed=3)
    assert np.all(ok["status"] == 0) and ok["status_code"] == 0
    bad = th0.copy(); bad[1, 12] = np.nan
    out = chain.run(bad, 6, seed=3, raise_on_status=False)
    assert out["status"][0] == 0 and out["status"][1] != 0 and out["status_code"] != 0
    for k in ("accepted", "values", "theta"):
        assert np.array_equal(out[k][:, 0], ok[k][:, 0])        # the healthy chain is unaffected
    chain.close(); ev.close(); gp.close(); model.close(); tgt.close()


def test_periodic_metrics_of_best_sample(ctx, twin31):
    """SamplingRegistration.scala:75-82: boundary-aware registration measures of the best sample every interval steps."""
    m = twin31
    K = 31
    model, tgt = _dev(ctx, m)
    ids, eids = np.arange(62), np.arange(124)
    tp = m["target"][::26][:62]
    mk = lambda d: core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, d, True, ids, tp)
    comps = [dict(kind=0, weight=0.45, proposal=mk(1)), dict(kind=0, weight=0.45, proposal=mk(0)), dict(kind=1, weight=0.1, sd=0.1)]
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 0, True, 0.0, 2.0, 0.0, eids, tp)
    chain = core.Chain(model, tgt, comps, ev, max_chains=4)
    th0 = random_theta(m, np.random.default_rng(8), 3, alpha_sd=0.5)
    n, every = 24, 8
    out = chain.run(th0, n, seed=5, metrics_interval=every)
    assert out["metrics"].shape == (n // every, 3, 4)
    ref = chain.run(th0, n, seed=5)
    for k in ("accepted", "values", "theta"):
        assert np.array_equal(out[k], ref[k])                    # measuring does not perturb the chain
    for r in range(n // every):
        upto = (r + 1) * every
        for c in range(3):
            vals = np.concatenate([[orc_value0(model, ev, th0[c])], out["values"][:upto, c, 0]])
            b = int(np.argmax(vals))
            theta_b = th0[c] if b == 0 else out["theta"][b - 1, c]
            want = core.registration_metrics(model, tgt, theta_b[None])[0]
            np.testing.assert_allclose(out["metrics"][r, c], want, rtol=1e-12)
    # improvement: the last row is not worse than the first on average
    assert out["metrics"][-1, :, 0].mean() <= out["metrics"][0, :, 0].mean() + 1e-9
    chain.close(); ev.close(); model.close(); tgt.close()


def orc_value0(model, ev, theta):
    return ev.log_value(theta[None])[0, 0]


def test_shared_handles_from_two_threads(ctx, twin31):
    """The reference calls one proposal / evaluator object from up to 10 JVM threads (RunMHRandomInitComparison.scala:59-86):
    concurrent calls on the same handles must return what the serial calls return."""
    m = twin31
    K = 31
    model, tgt = _dev(ctx, m)
    ids = np.arange(62)
    tp = m["target"][::26][:62]
    gp = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, 0, True, ids, tp)
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 2, True, 0.0, 2.0, 0.0, np.arange(124), tp)
    rng = np.random.default_rng(6)
    n_threads, reps = 4, 12
    th = [random_theta(m, rng, 2 + t, pose=True) for t in range(n_threads)]
    z = [rng.normal(size=(len(x), K)) for x in th]
    serial = [(gp.propose(th[t], z[t]), gp.log_transition(th[t], gp.propose(th[t], z[t])), ev.log_value(th[t])) for t in range(n_threads)]
    errors = []

    def worker(t):
        try:
            for _ in range(reps):
                p = gp.propose(th[t], z[t])
                lt = gp.log_transition(th[t], p)
                v = ev.log_value(th[t])
                if not (np.array_equal(p, serial[t][0]) and np.array_equal(lt, serial[t][1]) and np.array_equal(v, serial[t][2])):
                    errors.append(f"thread {t}: result differs from the serial call")
        except Exception as e:   # noqa: BLE001
            errors.append(f"thread {t}: {e!r}")

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for x in ts:
        x.start()
    for x in ts:
        x.join()
    assert not errors, errors
    ev.close(); gp.close(); model.close(); tgt.close()


@pytest.mark.parametrize("n_chains,width", [(1, 8), (3, 4), (5, 32), (2, 2), (1, -1), (4, -1), (20, -1)])
def test_rejection_lookahead_is_bit_identical(ctx, twin31, n_chains, width):
    """icp_chain_set_lookahead: `width` lanes per chain evaluate the proposals of the next `width` steps from the same current
    state in one batched round (SamplingRegistration.scala:60-85 is sequential; a rejected step leaves the state where it was and
    the randomness of a step depends on (seed, chain, step) only). The chain log - components, accept decisions, log-values,
    states - the accepted counts, the best sample and the final state must be bit-identical to the step-by-step runner's, with
    the Philox streams and with caller-supplied random numbers, on a mixture with ICP, random-walk and pose components."""
    m = twin31
    K = 31
    model, tgt = _dev(ctx, m)
    ids = np.arange(62)
    tp = m["target"][::26][:62]
    pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp)
    pm = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp, rank_update=_lib.RANK_UPDATE_INT8)
    comps = [dict(kind=_lib.PROP_ICP, weight=0.3, proposal=pt), dict(kind=_lib.PROP_ICP, weight=0.3, proposal=pm),
             dict(kind=_lib.PROP_RANDOM_SHAPE, weight=0.3, sd=0.15), dict(kind=2, weight=0.05, sd=0.01, axis=1), dict(kind=3, weight=0.05, sd=0.2, axis=0)]
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, _lib.MODEL_TO_TARGET, True, 0.0, 1.0, 0.0, np.arange(124), tp)
    rng = np.random.default_rng(60 + n_chains)
    th0 = random_theta(m, rng, n_chains, alpha_sd=0.4)
    n = 45 if width > 0 else 150     # the automatic width times every active width first (8 rounds each), then adapts
    keys = ("component", "accepted", "values", "theta", "theta_final", "n_accepted", "theta_best", "value_best", "status")
    for host_rng in (False, True):
        kw = dict(seed=77)
        if host_rng:
            kw = dict(u_comp=rng.random((n, n_chains)), z=rng.normal(size=(n, n_chains, K)), u_acc=rng.random((n, n_chains)))
        plain = core.Chain(model, tgt, comps, ev, max_chains=n_chains)
        plain.set_lookahead(0)
        want = plain.run(th0, n, **kw)
        assert plain.last_run_rounds() == n
        ahead = core.Chain(model, tgt, comps, ev, max_chains=n_chains)
        ahead.set_lookahead(width)
        got = ahead.run(th0, n, **kw)
        for k in keys:
            assert np.array_equal(got[k], want[k], equal_nan=True), f"{k} differs (host_rng={host_rng})"
        rate = want["n_accepted"].mean() / n
        assert 0.05 < rate < 0.98                      # both branches of the resolution are exercised
        assert ahead.last_run_rounds() < n or width < 0   # rounds took more than one step (the automatic width may settle at one lane)
        plain.close(); ahead.close()
    # resumed device-resident run: the look-ahead state carries over, the step counter and the Philox stream continue
    import torch
    dev = torch.device("cuda", ctx.device if hasattr(ctx, "device") else 0)
    L = K + 10
    outs = []
    for width_ in (0, width):
        ch = core.Chain(model, tgt, comps, ev, max_chains=n_chains)
        ch.set_lookahead(width_)
        t0 = torch.from_numpy(th0).to(dev)
        logs = []
        for part, first in ((20, True), (25, False)):
            acc = torch.zeros((part, n_chains), dtype=torch.uint8, device=dev); tl = torch.zeros((part, n_chains, L), dtype=torch.float64, device=dev)
            ch.run_device(n_chains, part, t0.data_ptr() if first else None, seed=77, log_accepted=acc.data_ptr(), log_theta=tl.data_ptr())
            logs += [acc.cpu().numpy(), tl.cpu().numpy()]
        outs.append(logs)
        ch.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    ev.close(); pt.close(); pm.close(); model.close(); tgt.close()


def test_hausdorff_early_break_is_exact(ctx, twin31):
    """HausdorffDistanceEvaluator.scala:31-35 uses only the largest distance of the two directions. The device stops every
    closest-point query that cannot raise the chain's running maximum (k_nearest, HDMAX); the value must stay the exact one:
    compared with the Hausdorff distance of icp_registration_metrics (every query run to the end) on 333 chains with pose,
    on a cold evaluator (no seeds: nothing can be pruned before the first queries finish) and again on the warm one."""
    m = twin31
    model, tgt = _dev(ctx, m)
    rng = np.random.default_rng(46)
    th = random_theta(m, rng, 333, alpha_sd=0.6, pose=True)
    hd = core.registration_metrics(model, tgt, th)[:, 1]
    want = np.log(100.0) - 100.0 * hd
    ev = core.Evaluator(model, tgt, _lib.EVAL_HAUSDORFF, 0, False, 100.0)
    for _ in range(3):
        v = ev.log_value(th)
        np.testing.assert_allclose(v[:, 2], want, rtol=1e-12, atol=0)
    # moved states on the warm evaluator (stale seeds), and a single chain (every warp works on the same maximum)
    th2 = th + np.concatenate([np.zeros((333, 10)), rng.normal(0, 0.05, (333, 31))], axis=1)
    np.testing.assert_allclose(ev.log_value(th2)[:, 2], np.log(100.0) - 100.0 * core.registration_metrics(model, tgt, th2)[:, 1], rtol=1e-12)
    np.testing.assert_allclose(ev.log_value(th2[7:8])[:, 2], np.log(100.0) - 100.0 * core.registration_metrics(model, tgt, th2[7:8])[:, 1], rtol=1e-12)
    ev.close(); model.close(); tgt.close()


def test_ten_threads_on_shared_handles_and_shared_cache(ctx, twin31):
    """RunMHRandomInitComparison.scala:59-86 as it is written: ONE proposal mixture and ONE evaluator shared by ten fitting
    threads. Concurrent calls on a handle run on the handle's call slots (own stream, scratch and replayed graph each) and share
    its posterior cache: all ten walks start from the same state, so the first calls collide on one cache key (one call computes
    the posterior, the others wait for it). Every walk must equal the same walk made alone on fresh handles."""
    m = twin31
    K = 31
    model, tgt = _dev(ctx, m)
    ids = np.arange(62)
    tp = m["target"][::26][:62]
    n_threads, steps = 10, 8
    rng = np.random.default_rng(26)
    th0 = random_theta(m, rng, 1)
    z = rng.normal(size=(n_threads, steps, 1, K))

    def handles():
        return (core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp),
                core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp, rank_update=_lib.RANK_UPDATE_INT8),
                core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, _lib.MODEL_TO_TARGET, True, 0.0, 2.0, 0.0, np.arange(124), tp))

    def walk(t, hs):
        pm, pt, ev = hs
        out, th = [], th0
        for k in range(steps):
            new = (pm if (k + t) % 2 == 0 else pt).propose(th, z[t, k])
            out += [new, pm.log_transition(th, new), pm.log_transition(new, th), pt.log_transition(th, new), pt.log_transition(new, th),
                    ev.log_value(new)]
            th = new
        return out

    shared = handles()
    results, errors = [None] * n_threads, []

    def worker(t):
        try:
            results[t] = walk(t, shared)
        except Exception as e:   # noqa: BLE001
            errors.append(f"thread {t}: {e!r}")

    for _round in range(2):     # the second round replays the graphs the first one captured
        ts = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
        for x in ts:
            x.start()
        for x in ts:
            x.join()
        assert not errors, errors
        for t in range(n_threads):
            alone = handles()
            want = walk(t, alone)
            for h in alone:
                h.close()
            for a, b in zip(results[t], want):
                assert np.array_equal(a, b, equal_nan=True), f"round {_round}, thread {t}: differs from the walk made alone"
    for h in shared:
        h.close()
    model.close(); tgt.close()


def test_per_thread_handles_run_concurrently(ctx, twin31):
    """RunMHRandomInitComparison.scala:59-86: ten fitting threads, each with its OWN proposals and evaluators over the shared
    model and target. Self-contained handles (model-sampling and target-sampling ICP proposals, model->target evaluators) run
    on their own streams under their own locks, the symmetric evaluator (it refits the model's BVH) takes the context lock;
    every thread must get exactly what the same calls return serially."""
    m = twin31
    K = 31
    model, tgt = _dev(ctx, m)
    ids = np.arange(62)
    tp = m["target"][::26][:62]
    n_threads, steps = 10, 6
    rng = np.random.default_rng(16)
    th0 = [random_theta(m, rng, 1, pose=(t % 2 == 0)) for t in range(n_threads)]
    z = rng.normal(size=(n_threads, steps, 1, K))

    def handles():
        return (core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, 0, True, ids, tp), core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, 1, True, ids, tp),
                core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 0, True, 0.0, 2.0, 0.0, np.arange(124), tp),
                core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 2, True, 0.0, 2.0, 0.0, np.arange(124), tp))

    def walk(t, hs):
        pm, pt, em, es = hs
        out, th = [], th0[t]
        for k in range(steps):
            p = (pm if k % 2 == 0 else pt)
            new = p.propose(th, z[t, k])
            out += [new, pm.log_transition(th, new), pm.log_transition(new, th), pt.log_transition(th, new), pt.log_transition(new, th),
                    em.log_value(new), es.log_value(new)]
            th = new
        return out

    serial = []
    for t in range(n_threads):
        hs = handles()
        serial.append(walk(t, hs))
        for h in hs:
            h.close()
    all_handles = [handles() for _ in range(n_threads)]
    results, errors = [None] * n_threads, []

    def worker(t):
        try:
            results[t] = walk(t, all_handles[t])
        except Exception as e:   # noqa: BLE001
            errors.append(f"thread {t}: {e!r}")

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for x in ts:
        x.start()
    for x in ts:
        x.join()
    assert not errors, errors
    for t in range(n_threads):
        assert len(results[t]) == len(serial[t])
        for a, b in zip(results[t], serial[t]):
            assert np.array_equal(a, b, equal_nan=True), f"thread {t}: concurrent result differs from the serial one"
    for hs in all_handles:
        for h in hs:
            h.close()
    model.close(); tgt.close()


def test_step_graph_survives_reallocation_of_model_scratch(ctx, twin31):
    """A cached step graph bakes in the model's shared BVH scratch. Run chain A at C = 8 (symmetric evaluator: the graph
    refits the model triangle BVH), make another call reallocate that scratch with a larger batch, then re-run A with the
    same seed and buffers: the log must be identical (the graph is re-captured, not replayed with dangling pointers)."""
    m = twin31
    K = 31
    model, tgt = _dev(ctx, m)
    ids, eids = np.arange(62), np.arange(124)
    tp = m["target"][::26][:62]
    mk = lambda d: core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, d, True, ids, tp)
    comps = [dict(kind=0, weight=0.45, proposal=mk(1)), dict(kind=0, weight=0.45, proposal=mk(0)), dict(kind=1, weight=0.1, sd=0.1)]
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 2, True, 0.0, 2.0, 0.0, eids, tp)
    chain = core.Chain(model, tgt, comps, ev, max_chains=8)
    rng = np.random.default_rng(2)
    th0 = random_theta(m, rng, 8, alpha_sd=0.4)
    a1 = chain.run(th0, 12, seed=11)
    big = random_theta(m, rng, 64, pose=True)
    q = synth.near_surface_queries(m["target"], m["target_cells"], 256)
    model.closest_point_surface(big, q)          # C = 64 on the same model: grows tri_bvh.nodes / nodebox
    model.closest_vertex(big, q)                 # and the vertex BVH scratch
    a2 = chain.run(th0, 12, seed=11)
    for k in ("component", "accepted", "values", "theta"):
        assert np.array_equal(a1[k], a2[k]), k
    chain.close(); ev.close(); model.close(); tgt.close()


def test_handles_are_reference_counted(ctx, twin31):
    """SURVEY 8b ownership: destroying a model / target / proposal that a live handle still references is an error, not a
    dangling pointer."""
    m = twin31
    model, tgt = _dev(ctx, m)
    ids = np.arange(62)
    tp = m["target"][::26][:62]
    gp = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, 0, True, ids, tp)
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 0, True, 0.0, 2.0, 0.0, np.arange(124), tp)
    chain = core.Chain(model, tgt, [dict(kind=0, weight=1.0, proposal=gp)], ev, max_chains=2)
    lib = ctx.lib
    assert lib.icp_model_destroy(model.h) == _lib.ERR_INVALID_ARGUMENT and "still referenced" in _lib.last_error(ctx.h)
    assert lib.icp_target_destroy(tgt.h) == _lib.ERR_INVALID_ARGUMENT
    assert lib.icp_proposal_destroy(gp.h) == _lib.ERR_INVALID_ARGUMENT
    assert lib.icp_evaluator_destroy(ev.h) == _lib.ERR_INVALID_ARGUMENT
    th0 = random_theta(m, np.random.default_rng(0), 2)
    assert chain.run(th0, 3)["theta"].shape == (3, 2, 41)          # everything is still alive
    for h, fn in ((chain, lib.icp_chain_destroy), (gp, lib.icp_proposal_destroy), (ev, lib.icp_evaluator_destroy),
                  (model, lib.icp_model_destroy), (tgt, lib.icp_target_destroy)):
        assert fn(h.h) == _lib.OK
        h.h = None


def test_std_icp_iteration_with_rigid_transform(ctx, femur):
    """IcpBasedSurfaceFitting.scala:55-92 with a non-identity currentTrans (:61): correspondences on the transformed instance,
    posterior of the untransformed model on world-frame targets (:81)."""
    m = _femur(femur, "gpmm_50")
    K = 51
    model, tgt = _dev(ctx, m)
    om, ot = _orc(m)
    rng = np.random.default_rng(15)
    th = random_theta(m, rng, 2, pose=True)
    ids = rng.integers(0, 1622, 400)
    tp = synth.near_surface_queries(m["target"], m["target_cells"], 400, sd=0.0)
    for direction in (0, 1):
        out = core.std_icp_iteration_theta(model, tgt, direction, ids, tp, 1e-15, 1.0, th)
        for c in range(2):
            want = orc.std_icp_iteration_theta(om, ot, direction, ids, tp, 1e-15, 1.0, th[c])
            np.testing.assert_allclose(out[c], want, rtol=RTOL, atol=1e-7)
        # identity pose reduces to the original entry point
        th_id = th.copy(); th_id[:, 1:10] = 0.0
        np.testing.assert_array_equal(core.std_icp_iteration_theta(model, tgt, direction, ids, tp, 1e-15, 1.0, th_id),
                                      core.std_icp_iteration(model, tgt, direction, ids, tp, 1e-15, 1.0, th_id[:, 10:]))
    model.close(); tgt.close()


@pytest.mark.parametrize("direction", [_lib.MODEL_SAMPLING, _lib.TARGET_SAMPLING])
@pytest.mark.parametrize("fixture", ["twin31", "femur100", "open_twin"])
def test_int8_tensor_core_rank_update(ctx, request, femur, direction, fixture):
    """ICP_RANK_UPDATE_INT8: the posterior's rank update on tcgen05 INT8 tensor cores (split-integer emulation of the FP64
    product) against the oracle at the contract tolerance (1e-5 relative on posterior means and transition densities), and
    against the FP64 path of the library."""
    m = _femur(femur, "gpmm_100") if fixture == "femur100" else request.getfixturevalue(fixture)
    K = len(m["variance"])
    model, tgt = _dev(ctx, m)
    om, ot = _orc(m)
    rng = np.random.default_rng(23)
    C = 5
    th = random_theta(m, rng, C, pose=(fixture != "femur100"))
    if fixture == "open_twin":
        ids, tp = np.arange(0, len(m["ref"]), 5), m["target"][::4]         # boundary filtering drops observations
    else:
        ids = np.arange(2 * K)
        tp = m["target"][:: max(1, len(m["target"]) // (2 * K))][: 2 * K] + rng.normal(0, 0.1, (2 * K, 3))
    gi = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, direction, True, ids, tp, rank_update=_lib.RANK_UPDATE_INT8)
    gf = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, direction, True, ids, tp)
    op = orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, direction, True, ids, tp)
    mu, M, n = gi.posterior(th)
    mu64, M64, n64 = gf.posterior(th)
    assert np.array_equal(n, n64)
    scale = np.abs(M64).max(axis=(1, 2), keepdims=True)
    assert np.abs(M - M64).max() / scale.max() < 2e-8 and np.abs(M - M64).max() > 0.0      # a different arithmetic, 2e-9 typical
    z = rng.normal(size=(C, K))
    prop = gi.propose(th, z)
    lt = gi.log_transition(th, prop)
    for c in range(C):
        po = op.posterior(th[c])
        assert n[c] == po["n"]
        np.testing.assert_allclose(M[c], po["M"], rtol=0, atol=2e-8 * np.abs(po["M"]).max())
        np.testing.assert_allclose(mu[c], po["mu"], rtol=RTOL, atol=RTOL * np.abs(po["mu"]).max())
        assert np.abs(mu[c] - po["mu"]).max() <= 1e-6 * np.abs(po["mu"]).max()        # measured ~1e-8
        np.testing.assert_allclose(prop[c], op.propose(th[c], z[c], closed_form=True), rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(lt[c], op.log_transition(th[c], prop[c], closed_form=True), rtol=RTOL)
    gi.close(); gf.close(); model.close(); tgt.close()


def test_int8_rank_update_chain_matches_oracle(ctx, femur):
    """Config 1 (femur GPMM-100) with both ICP proposals on the INT8 tensor-core path: the fused runner against the oracle chain."""
    m = _femur(femur, "gpmm_100")
    K = 101
    model, tgt = _dev(ctx, m)
    om, ot = _orc(m)
    ids, eids = np.arange(2 * K), np.arange(4 * K)
    tp = m["target"][:: len(m["target"]) // (2 * K)][: 2 * K]
    mk = lambda d: core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, d, True, ids, tp, rank_update=_lib.RANK_UPDATE_INT8)
    mo = lambda d: orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, d, True, ids, tp)
    comps = [dict(kind=0, weight=0.45, proposal=mk(1)), dict(kind=0, weight=0.45, proposal=mk(0)), dict(kind=1, weight=0.1, sd=0.1)]
    comps_o = [dict(kind=0, weight=0.45, icp=mo(1)), dict(kind=0, weight=0.45, icp=mo(0)), dict(kind=1, weight=0.1, sd=0.1)]
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 0, True, 0.0, 2.0, 0.0, eids, tp)
    chain = core.Chain(model, tgt, comps, ev, max_chains=4)
    rng = np.random.default_rng(77)
    n, C = 40, 3
    th0 = np.stack([model.theta()] + [model.theta(rng.normal(0, 0.3, K)) for _ in range(C - 1)])
    u_comp, u_acc, z = rng.random((n, C)), rng.random((n, C)), rng.normal(size=(n, C, K))
    got = chain.run(th0, n, u_comp=u_comp, z=z, u_acc=u_acc)
    for c in range(C):
        want = orc.chain_run(om, ot, comps_o, True, orc.EVAL_INDEPENDENT, 0, (0.0, 2.0), eids, tp, th0[c], n, u_comp[:, c], z[:, c],
                             u_acc[:, c], closed_form=True)
        assert np.array_equal(got["component"][:, c], want["comp"])
        assert np.array_equal(got["accepted"][:, c], want["accepted"])
        np.testing.assert_allclose(got["values"][:, c], want["logv"], rtol=RTOL)
        np.testing.assert_allclose(got["theta"][:, c], want["theta"], rtol=0, atol=1e-5)
    assert np.all(got["status"] == 0)
    chain.close(); ev.close(); model.close(); tgt.close()


@pytest.mark.parametrize("rank,n_obs", [(20, 1), (20, 33), (110, 95), (105, 202)])
def test_int8_rank_update_shapes_and_persistence(ctx, rank, n_obs):
    """The persistent INT8 kernel at its edges: column groups that do not fill a converter pair (Kp = 24), the widest tile
    (Kp = 112), one observation, observation counts off the 16 / 32 stage grid, and more chains than SMs (every CTA walks
    several chains: ring slots, barrier parities and the early first step of the next chain carry over). Checked against the
    FP64 tensor-pipe path of the same library on every chain."""
    m = synth.femur_twin(rank=rank, n=600)
    tv, tc, _ = synth.synthetic_target(m)
    model, tgt = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"]), core.Target(ctx, tv, tc)
    rng = np.random.default_rng(rank + n_obs)
    C = 333
    th = random_theta(m, rng, C, pose=True)
    ids = np.sort(rng.choice(len(m["ref"]), n_obs, replace=False))
    tp = tv[np.sort(rng.choice(len(tv), n_obs, replace=False))] + rng.normal(0, 0.2, (n_obs, 3))
    for direction in (_lib.MODEL_SAMPLING, _lib.TARGET_SAMPLING):
        gi = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, direction, True, ids, tp, rank_update=_lib.RANK_UPDATE_INT8)
        gf = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, direction, True, ids, tp)
        mu, M, n = gi.posterior(th)
        mu64, M64, n64 = gf.posterior(th)
        assert np.array_equal(n, n64)
        scale = np.abs(M64).max(axis=(1, 2))
        err = np.abs(M - M64).max(axis=(1, 2)) / scale
        assert err.max() < 2e-8 and err.max() > 0.0, (direction, err.max())
        assert (np.abs(mu - mu64).max(axis=1) <= 1e-6 * np.maximum(np.abs(mu64).max(axis=1), 1e-3)).all()
        gi.close(); gf.close()
    model.close(); tgt.close()
