"""GPU parity on the shapes of BASELINE.json's configs: the device chain / fitting against the
oracle on the same inputs. Sizes are the named ones where the oracle finishes in seconds, reduced step counts."""
import numpy as np
import pytest

from conftest import random_theta
from icp_proposal_b200 import _lib, core, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _femur(femur, which):
    return dict(ref=femur["ref"], cells=femur["cells"], target=femur["target"], target_cells=femur["target_cells"], **femur[which])


def _both(ctx, m):
    model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
    tgt = core.Target(ctx, m["target"], m["target_cells"])
    om = orc.Model(m["ref"], m["cells"], m["basis"], m["variance"])
    ot = orc.Mesh(m["target"], m["target_cells"])
    return model, tgt, om, ot


def _compare_chain(got, want, c, rtol=1e-7, atol=1e-7):
    assert np.array_equal(got["component"][:, c], want["comp"])
    assert np.array_equal(got["accepted"][:, c], want["accepted"])
    np.testing.assert_allclose(got["values"][:, c], want["logv"], rtol=rtol)
    np.testing.assert_allclose(got["theta"][:, c], want["theta"], rtol=0, atol=atol)


def test_config1_femur_gpmm100_icp_proposal_chain(ctx, femur):
    """Config 1: femur GPMM-100 (the reference's own model and target), n_icp = 2K, n_eval = 4K, mixture
    0.9 * (0.5 target-ICP + 0.5 model-ICP) + 0.1 * RW(0.1), prior x Gaussian-point(sd 2) model->target."""
    m = _femur(femur, "gpmm_100")
    K = 101
    model, tgt, om, ot = _both(ctx, m)
    ids, eids = np.arange(2 * K), np.arange(4 * K)
    tp = m["target"][:: len(m["target"]) // (2 * K)][: 2 * K]
    mk = lambda d: core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, d, True, ids, tp)
    mo = lambda d: orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, d, True, ids, tp)
    comps = [dict(kind=0, weight=0.45, proposal=mk(1)), dict(kind=0, weight=0.45, proposal=mk(0)), dict(kind=1, weight=0.1, sd=0.1)]
    comps_o = [dict(kind=0, weight=0.45, icp=mo(1)), dict(kind=0, weight=0.45, icp=mo(0)), dict(kind=1, weight=0.1, sd=0.1)]
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 0, True, 0.0, 2.0, 0.0, eids, tp)
    chain = core.Chain(model, tgt, comps, ev, max_chains=4)
    rng = np.random.default_rng(1024)
    n, C = 60, 2
    th0 = np.stack([model.theta(), model.theta(rng.normal(0, 0.3, K))])     # chain 0 starts at the mean (SamplingRegistration.scala:40-43)
    u_comp, u_acc, z = rng.random((n, C)), rng.random((n, C)), rng.normal(size=(n, C, K))
    got = chain.run(th0, n, u_comp=u_comp, z=z, u_acc=u_acc)
    for c in range(C):
        want = orc.chain_run(om, ot, comps_o, True, orc.EVAL_INDEPENDENT, 0, (0.0, 2.0), eids, tp, th0[c], n, u_comp[:, c], z[:, c],
                             u_acc[:, c], closed_form=True)
        _compare_chain(got, want, c)
        assert want["n_accepted"] > 5
    # and against the chain in the REFERENCE's structure (SVD-rotated basis, full-mesh regressions) for the first steps:
    # z = 0 makes the proposals independent of the covariance factor
    n2 = 6
    z0 = np.zeros((n2, 1, K))
    got0 = chain.run(th0[1:2], n2, u_comp=u_comp[:n2, 1:2], z=z0, u_acc=u_acc[:n2, 1:2])
    want0 = orc.chain_run(om, ot, comps_o, True, orc.EVAL_INDEPENDENT, 0, (0.0, 2.0), eids, tp, th0[1], n2, u_comp[:n2, 1], z0[:, 0],
                          u_acc[:n2, 1], closed_form=False)
    _compare_chain(got0, want0, 0, rtol=1e-5, atol=1e-5)
    chain.close(); ev.close(); model.close(); tgt.close()


def test_config2_femur_gpmm50_standard_icp(ctx, femur):
    """Config 2: deterministic ICP, K = 51, n = N samples per direction, sigma^2 = 1e-15, step 1.0 (IcpRegistration.scala:40-43)."""
    m = _femur(femur, "gpmm_50")
    K, N = 51, 1622
    model, tgt, om, ot = _both(ctx, m)
    rng = np.random.default_rng(1024)
    ids = rng.integers(0, N, N)                                              # nearest vertices of uniform surface samples
    tp = synth.near_surface_queries(m["target"], m["target_cells"], N, seed=1024, sd=0.0)
    alpha_d = np.zeros((1, K)); alpha_o = np.zeros(K)
    directions = [0, 1, 0, 0, 1, 1, 0, 1]
    for d in directions:
        alpha_d = core.std_icp_iteration(model, tgt, d, ids, tp, 1e-15, 1.0, alpha_d)
        alpha_o = orc.std_icp_iteration(om, ot, d, ids, tp, 1e-15, 1.0, alpha_o)
        np.testing.assert_allclose(alpha_d[0], alpha_o, rtol=1e-5, atol=1e-6)
    dist = np.sqrt(tgt.closest_point_surface(model.reconstruct(model.theta(alpha_d[0], center=(0, 0, 0)))[0])[3])
    d0 = np.sqrt(tgt.closest_point_surface(m["ref"])[3])
    assert dist.mean() < 1.0 and dist.mean() < 0.2 * d0.mean()    # 8 iterations with the 51-component model
    model.close(); tgt.close()


def test_config3_random_init_chains_symmetric_full_mesh(ctx, femur):
    """Config 3: 5 random-init chains batched, n_icp = n_eval = N, ModelSampling, SymmetricEvaluation sd 2
    (RunMHRandomInitComparison.scala:54-61)."""
    m = _femur(femur, "gpmm_100")
    K, N = 101, 1622
    model, tgt, om, ot = _both(ctx, m)
    ids = np.arange(N)
    tp = m["target"]
    gp = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, 0, True, ids, tp)
    op = orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, 0, True, ids, tp)
    comps = [dict(kind=0, weight=0.9, proposal=gp), dict(kind=1, weight=0.1, sd=0.1)]
    comps_o = [dict(kind=0, weight=0.9, icp=op), dict(kind=1, weight=0.1, sd=0.1)]
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, _lib.SYMMETRIC, True, 0.0, 2.0, 0.0, ids, tp)
    chain = core.Chain(model, tgt, comps, ev, max_chains=5)
    rng = np.random.default_rng(3)
    n, C = 12, 5
    th0 = np.stack([model.theta(np.random.default_rng(s).normal(0, 0.3, K)) for s in range(C)])
    u_comp, u_acc, z = rng.random((n, C)), rng.random((n, C)), rng.normal(size=(n, C, K))
    got = chain.run(th0, n, u_comp=u_comp, z=z, u_acc=u_acc)
    for c in (0, 3):
        want = orc.chain_run(om, ot, comps_o, True, orc.EVAL_INDEPENDENT, 2, (0.0, 2.0), ids, tp, th0[c], n, u_comp[:, c], z[:, c],
                             u_acc[:, c], closed_form=True)
        _compare_chain(got, want, c)
    chain.close(); ev.close(); gp.close(); model.close(); tgt.close()


@pytest.mark.parametrize("which", ["gpmm_100", "gpmm_200"])
def test_config4_hausdorff_chains_on_perturbed_targets(ctx, femur, which):
    """Config 4: random inits (alpha0 = 0 for index 0, else N(0, 0.1 I)) x synthetic perturbed targets, Hausdorff
    evaluator Exponential(100), config-1 mixture (StdIcpVsChainICPrandomInitComparisonAll.scala:100-160). BASELINE.json
    names GPMM-100; the shipped main opens the 200-component file (rank 201, :88), which takes the large-rank kernels
    (rank update with 12 consumer warps, block-packed factorisation at one chain per SM)."""
    base = _femur(femur, which)
    K = len(base["variance"])
    for t_index in range(2 if which == "gpmm_100" else 1):
        target = synth.model_instance(base, np.random.default_rng(100 + t_index).normal(0, 1.0, K))
        target = target + np.random.default_rng(7 + t_index).normal(0, 0.5, target.shape)      # 0.5 mm vertex noise
        m = dict(base, target=target, target_cells=base["cells"])
        model, tgt, om, ot = _both(ctx, m)
        ids = np.arange(2 * K)
        tp = target[:: len(target) // (2 * K)][: 2 * K]
        mk = lambda d: core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, d, True, ids, tp)
        mo = lambda d: orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, d, True, ids, tp)
        comps = [dict(kind=0, weight=0.45, proposal=mk(1)), dict(kind=0, weight=0.45, proposal=mk(0)), dict(kind=1, weight=0.1, sd=0.1)]
        comps_o = [dict(kind=0, weight=0.45, icp=mo(1)), dict(kind=0, weight=0.45, icp=mo(0)), dict(kind=1, weight=0.1, sd=0.1)]
        ev = core.Evaluator(model, tgt, _lib.EVAL_HAUSDORFF, 0, True, 100.0)
        chain = core.Chain(model, tgt, comps, ev, max_chains=4)
        rng = np.random.default_rng(11 + t_index)
        n, C = 25, 3
        th0 = np.stack([model.theta()] + [model.theta(rng.normal(0, np.sqrt(0.1), K)) for _ in range(C - 1)])
        u_comp, u_acc, z = rng.random((n, C)), rng.random((n, C)), rng.normal(size=(n, C, K))
        got = chain.run(th0, n, u_comp=u_comp, z=z, u_acc=u_acc)
        for c in range(C):
            want = orc.chain_run(om, ot, comps_o, True, orc.EVAL_HAUSDORFF, 0, (100.0,), ids, tp, th0[c], n, u_comp[:, c], z[:, c],
                                 u_acc[:, c], closed_form=True)
            _compare_chain(got, want, c)
        chain.close(); ev.close(); model.close(); tgt.close()


def test_config5_face_sized_partial_target_with_pose(ctx):
    """Config 5: face-sized open surface (169 x 169 grid: N = 28 561, T = 56 448, boundary present), analytic GPMM K = 100,
    partial target, ModelSampling boundary-aware ICP + pose proposals, collective avg + max evaluator (sd 0.3, mean 0.1,
    rate 1; BfmFittingPartial.scala:66-80)."""
    m = synth.face_twin(rank=100, side=169)
    tv, tc, _ = synth.partial_target(m, seed=7, alpha_sd=0.5)
    m["target"], m["target_cells"] = tv, tc
    K, N = 100, len(m["ref"])
    assert N == 28561 and len(m["cells"]) == 56448
    model, tgt, om, ot = _both(ctx, m)
    assert model.boundary_flags().sum() == 4 * 168 and tgt.boundary_flags().sum() > 4 * 168 - 200   # outer border + the holes
    rng = np.random.default_rng(5)
    ids = np.sort(rng.choice(N, 500, replace=False))            # decimated to 500 points "for speed up" (BfmFittingPartial.scala:46-48)
    tp = tv[np.sort(rng.choice(len(tv), 500, replace=False))]
    gp = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, 0, True, ids, tp)
    gt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, 1, True, ids, tp)
    op = orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, 0, True, ids, tp)
    opt = orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, 1, True, ids, tp)
    # primitives at this size first
    th = random_theta(m, rng, 2, pose=True)
    for g, o in ((gp, op), (gt, opt)):
        mu, M, n = g.posterior(th)
        for c in range(2):
            po = o.posterior(th[c])
            assert n[c] == po["n"] and 0 < n[c] < 500              # boundary hits were dropped
            np.testing.assert_allclose(M[c], po["M"], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(mu[c], po["mu"], rtol=1e-6, atol=1e-9)
    pose = [dict(kind=2, weight=0.05, sd=0.01, axis=a) for a in range(3)] + [dict(kind=3, weight=0.05, sd=0.1, axis=a) for a in range(3)]
    comps = [dict(kind=0, weight=0.35, proposal=gt), dict(kind=0, weight=0.35, proposal=gp), dict(kind=1, weight=0.1, sd=0.05)] + pose
    comps_o = [dict(kind=0, weight=0.35, icp=opt), dict(kind=0, weight=0.35, icp=op), dict(kind=1, weight=0.1, sd=0.05)] + pose
    ev = core.Evaluator(model, tgt, _lib.EVAL_COLLECTIVE, _lib.SYMMETRIC, True, 0.1, 0.3, 1.0, ids, tp)
    chain = core.Chain(model, tgt, comps, ev, max_chains=2)
    n, C = 14, 2
    th0 = random_theta(m, rng, C, alpha_sd=0.3)
    u_comp, u_acc, z = rng.random((n, C)), rng.random((n, C)), rng.normal(size=(n, C, K))
    got = chain.run(th0, n, u_comp=u_comp, z=z, u_acc=u_acc)
    for c in range(C):
        want = orc.chain_run(om, ot, comps_o, True, orc.EVAL_COLLECTIVE, 2, (0.1, 0.3, 1.0), ids, tp, th0[c], n, u_comp[:, c], z[:, c],
                             u_acc[:, c], closed_form=True)
        _compare_chain(got, want, c, rtol=1e-6, atol=1e-6)
    assert len(np.unique(got["component"])) >= 3
    chain.close(); ev.close(); model.close(); tgt.close()
