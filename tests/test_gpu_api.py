"""GPU tests of the host-side mirror of the reference API (icp-proposal_b200/api.py): the classes a Scalismo
chain would call (propose / logTransitionProbability / logValue), the fused SamplingRegistration and the logger."""
import json
import math

import numpy as np
import pytest

from icp_proposal_b200 import api
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(ctx, twin31):
    m = twin31
    model = api.StatisticalMeshModel(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
    target = api.TriangleMesh3D(ctx, m["target"], m["target_cells"])
    om = orc.Model(m["ref"], m["cells"], m["basis"], m["variance"])
    ot = orc.Mesh(m["target"], m["target_cells"])
    yield m, model, target, om, ot
    model.close(); target.close()


def test_per_call_api_matches_oracle(setup):
    m, model, target, om, ot = setup
    K = model.rank
    rng = np.random.default_rng(0)
    ids = np.arange(2 * K); tp = m["target"][::26][:2 * K]
    prop = api.NonRigidIcpProposal(model, target, 0.1, 10.0, 5.0, 2 * K, api.ModelSampling, rand=np.random.default_rng(5),
                                   model_point_ids=ids, target_points=tp)
    oprop = orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, 0, True, ids, tp)
    theta = model.initial_parameters().copy(shapeParameters=api.ShapeParameters(rng.normal(0, 0.3, K)))
    new = prop.propose(theta)
    assert new.generatedBy == "ShapeIcpProposal" and np.array_equal(new.allParameters[:10], theta.allParameters[:10])
    z = np.random.default_rng(5).standard_normal(K)       # the stream the proposal consumed
    np.testing.assert_allclose(new.allParameters, oprop.propose(theta.allParameters, z, closed_form=True), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(prop.logTransitionProbability(theta, new), oprop.log_transition(theta.allParameters, new.allParameters), rtol=1e-5)
    np.testing.assert_allclose(prop.logTransitionProbability(new, theta), oprop.log_transition(new.allParameters, theta.allParameters), rtol=1e-5)
    evs = api.ProductEvaluators.proximityAndIndependent(model, target, api.ModelToTargetEvaluation, 2.0, 4 * K, target_points=tp)
    assert list(evs) == ["product", "prior", "distance"]
    d = orc.eval_independent(om, ot, 0, 0.0, 2.0, np.arange(4 * K), tp, new.allParameters)
    p = orc.eval_prior(K, new.allParameters)
    np.testing.assert_allclose([evs[k].logValue(new) for k in evs], [p + d, p, d], rtol=1e-9)
    hd = api.ProductEvaluators.proximityAndHausdorff(model, target, 100.0)
    np.testing.assert_allclose(hd["distance_haussdorff"].logValue(new), orc.eval_hausdorff(om, ot, 100.0, new.allParameters), rtol=1e-9)


def test_host_metropolis_hastings_loop(setup):
    """The reference's own driver structure: Scalismo-style MH over the per-call (drop-in) classes."""
    m, model, target, om, ot = setup
    K = model.rank
    rng = np.random.default_rng(1)
    icp = api.MixedProposalDistributions.mixedProposalICP(model, target, 2 * K, rand=rng)
    rnd = api.MixedProposalDistributions.mixedRandomShapeProposal(model, rand=rng)
    gen = api.MixtureProposal.fromProposalsWithTransition((0.9, icp), (0.1, rnd), rand=rng)
    evs = api.ProductEvaluators.proximityAndIndependent(model, target, api.ModelToTargetEvaluation, 2.0, 4 * K)
    chain = api.MetropolisHastings(gen, evs["product"], rand=rng)
    logger = api.JSONAcceptRejectLogger(None, evs)
    theta = model.initial_parameters()
    v0 = evs["product"].logValue(theta)
    it = chain.iterator(theta, logger)
    for _ in range(25):
        theta = next(it)
    assert logger.totalSamples == 25 and 0 < logger.numOfAccepted <= 25
    assert evs["product"].logValue(theta) > v0
    names = {l.name for l in logger.logStatus}
    assert names <= {"IcpProposal-TargetSampling-0.1Step", "IcpProposal-ModelSampling-0.1Step", "RandomShape-0.1"}
    acc = [l for l in logger.logStatus if l.status]
    assert all(len(l.coeff) == K and len(l.rigid) == 9 for l in acc)


def test_sampling_registration_fused(setup, tmp_path):
    m, model, target, om, ot = setup
    K = model.rank
    icp = api.MixedProposalDistributions.mixedProposalICP(model, target, 2 * K)
    rnd = api.MixedProposalDistributions.mixedRandomShapeProposal(model)
    gen = api.MixtureProposal.fromProposalsWithTransition((0.9, icp), (0.1, rnd))
    evs = api.ProductEvaluators.proximityAndIndependent(model, target, api.ModelToTargetEvaluation, 2.0, 4 * K)
    reg = api.SamplingRegistration(model, target)
    path = str(tmp_path / "icpProposalRegistration.json")
    best = reg.runfitting(evs, gen, 1500, jsonName=path)
    log = json.load(open(path))
    assert len(log) == 1500 and set(log[0]) == {"index", "name", "logvalue", "status", "rigid", "coeff", "datetime"}
    assert set(log[0]["logvalue"]) == {"product", "prior", "distance"}
    # the registration moves the model towards the target (slowly: step 0.1, 5-10 mm observation noise, 62 points)
    # and the best sample's product value equals the best accepted log entry
    avg0, _ = api.RegistrationComparison.evaluateReconstruction2GroundTruth("init", model, model.initial_parameters(), target)
    avg1, hd1 = api.RegistrationComparison.evaluateReconstruction2GroundTruth("best", model, best, target)
    assert avg1 < 0.9 * avg0
    assert evs["product"].logValue(best) > evs["product"].logValue(model.initial_parameters()) + 5.0
    best_logged = max(l["logvalue"]["product"] for l in log if l["status"])
    np.testing.assert_allclose(evs["product"].logValue(best), best_logged, rtol=1e-9)
    # batched random-init chains (RunMHRandomInitComparison shape)
    th0 = np.tile(model.initial_parameters().allParameters, (5, 1))
    th0[1:, 10:] = np.random.default_rng(0).normal(0, math.sqrt(0.1), (4, K))
    bests = reg.runfitting(evs, gen, 200, n_chains=5, initial_batch=th0)
    assert len(bests) == 5
    for i, b in enumerate(bests):
        start = api.ModelFittingParameters.from_vector(th0[i])
        assert evs["product"].logValue(b) >= evs["product"].logValue(start)


def test_deterministic_icp(setup):
    m, model, target, om, ot = setup
    fit = api.IcpBasedSurfaceFitting(model, target, 400, 1.0, api.ModelSampling)
    mesh = fit.runfitting(15, iterationSeq=(1e-15,))
    d = np.sqrt(target.closest_point_surface(mesh)[3])
    d0 = np.sqrt(target.closest_point_surface(m["ref"])[3])
    assert d.mean() < 0.25 * d0.mean()


@pytest.mark.parametrize("sum_normals", [True, False])
def test_posterior_variability_matches_oracle(setup, sum_normals):
    """icp_posterior_variability against the loop-for-loop restatement of
    apps/util/PosteriorVariability.scala on the shapes LogHelper.logSamples2shapes would hand it."""
    from oracle import np_oracle as npo
    m, model, target, om, ot = setup
    K = model.rank
    rng = np.random.default_rng(21)
    S = 37
    thetas = np.zeros((S, 10 + K)); thetas[:, 0] = 1.0
    thetas[:, 10:] = rng.normal(0, 0.4, (S, K))
    thetas[:, 1:4] = rng.normal(0, 0.5, (S, 3)); thetas[:, 4:7] = rng.normal(0, 0.01, (S, 3)); thetas[:, 7:10] = m["ref"].mean(0)
    ref = thetas[5]
    meshes = [om.transformed_mesh(t) for t in thetas]
    want = npo.posterior_variability(meshes, m["cells"], ref_verts=om.transformed_mesh(ref), sum_normals=sum_normals)
    got = api.PosteriorVariability.statistics(model, thetas, ref=api.ModelFittingParameters.from_vector(ref), sumNormals=sum_normals)
    np.testing.assert_allclose(got["mean"], want[0], rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(got["cov"], want[1], rtol=1e-5, atol=1e-10)       # tolerance of BASELINE.json; observed ~1e-12
    np.testing.assert_allclose(got["total_variance"], want[2], rtol=1e-9)
    np.testing.assert_allclose(got["normal_variance"], want[3], rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(api.PosteriorVariability.computeDistanceMapFromMeshesTotal(model, thetas), got["total_variance"])
    np.testing.assert_array_equal(api.PosteriorVariability.computeDistanceMapFromMeshesNormal(
        model, thetas, api.ModelFittingParameters.from_vector(ref), sumNormals=sum_normals), got["normal_variance"])
    # through a chain log: accepted entries only, walked back from rejected ones (LogHelper.samplesFromLog)
    log = [api.jsonLogFormat(i, "p", {"product": 0.0}, i % 3 != 1, t[1:10].tolist() if i % 3 != 1 else [], t[10:].tolist() if i % 3 != 1 else [], "")
           for i, t in enumerate(thetas)]
    picked = api.LogHelper.samplesFromLog(log, takeEveryN=2, total=30, burnIn=3)
    idx = npo.samples_from_log([l.status for l in log], 2, 30, 3)
    assert [j for _, j in picked] == idx
    shapes = api.LogHelper.logSamples2shapes(model, [l for l, _ in picked])
    np.testing.assert_allclose(shapes, np.stack([meshes[j] for j in idx]), rtol=1e-12, atol=1e-10)
    tot = api.PosteriorVariability.computeDistanceMapFromMeshesTotal(model, [l for l, _ in picked])
    np.testing.assert_allclose(tot, npo.posterior_variability([meshes[j] for j in idx], m["cells"])[2], rtol=1e-9)
    # one sample: NaN like the reference; empty input is an error
    assert np.isnan(api.PosteriorVariability.computeDistanceMapFromMeshesTotal(model, thetas[:1])).all()
    with pytest.raises(Exception):
        api.PosteriorVariability.computeDistanceMapFromMeshesTotal(model, thetas[:0])


def test_gpmm_construction_matches_oracle(ctx, twin31):
    """the femur kernel of apps/femur/CreateGPModel.scala:70-83 and its Nystrom low-rank approximation
    (kernel matrix and extension on the device) against the pair-by-pair oracle; the resulting model reconstructs."""
    from oracle import np_oracle as npo
    ref = twin31["ref"]
    kernel = api.femurKernel(ref)
    assert len(kernel.terms) == 3 and kernel.terms[0][0] == 10.0 and kernel.terms[1][:2] == (5.0, 40.0) and kernel.terms[2][:2] == (3.0, 10.0)
    rng = np.random.default_rng(1024)
    rank = 31
    nys = ref[np.sort(rng.choice(len(ref), 2 * rank, replace=False))]           # numOfSamplePoints = 2 i (:84)
    from icp_proposal_b200 import core
    kmm = core.gpmm_kernel_matrix(ctx, nys, nys, kernel.terms)
    np.testing.assert_allclose(kmm, npo.gauss_mixture_kernel(nys, nys, kernel.terms), rtol=1e-12, atol=1e-300)
    sub = ref[::7]
    np.testing.assert_allclose(core.gpmm_kernel_matrix(ctx, sub, nys, kernel.terms), npo.gauss_mixture_kernel(sub, nys, kernel.terms),
                               rtol=1e-12, atol=1e-300)
    # Nystrom extension on the device against the oracle's, from the same eigenpairs
    w, v = np.linalg.eigh(npo.gauss_mixture_kernel(nys, nys, kernel.terms))
    order = np.argsort(w)[::-1][:rank]
    w, v = w[order], v[:, order]
    got_b, got_v = core.gpmm_nystrom_extend(ctx, ref, nys, kernel.terms, v, w)
    want_b, want_v = npo.nystrom_extend(npo.gauss_mixture_kernel(ref[:300], nys, kernel.terms), v, w)
    np.testing.assert_allclose(got_v, want_v, rtol=1e-14)
    np.testing.assert_allclose(got_b[:900], want_b, rtol=1e-9, atol=1e-12 * np.abs(want_b).max())
    # the device eigensolver (one-sided Jacobi) against LAPACK: eigenvalues, and the invariant subspace through its projector
    we, ve = core.gpmm_eigen_psd(ctx, kmm, rank)
    np.testing.assert_allclose(we, w, rtol=1e-10)
    np.testing.assert_allclose(ve.T @ ve, np.eye(rank), atol=1e-11)
    np.testing.assert_allclose(kmm @ ve, ve * we, atol=1e-9 * w[0])
    assert (ve[np.abs(ve).argmax(0), np.arange(rank)] > 0).all()
    # the whole construction (eigenvectors of near-degenerate pairs may rotate with the last bit of the kernel matrix, the
    # covariance they span may not): Q Q^T on the Nystrom points is the rank-truncated kernel matrix
    basis, var = api.LowRankGaussianProcess.approximateGPNystrom(ctx, kernel, ref, nys, rank)
    np.testing.assert_allclose(var, want_v, rtol=1e-9)
    # ... compared at a rank where the spectrum has a gap (the ellipsoid's near-symmetries pair eigenvalues up; a pair cut
    # in half by the truncation has no unique half)
    wa, va = np.linalg.eigh(npo.gauss_mixture_kernel(nys, nys, kernel.terms))
    wa, va = wa[::-1], va[:, ::-1]
    rg = max(r for r in range(8, rank + 1) if wa[r - 1] - wa[r] > 0.05 * wa[r - 1])
    bg, vg = api.LowRankGaussianProcess.approximateGPNystrom(ctx, kernel, ref, nys, rg)
    sel = np.sort(np.random.default_rng(1024).choice(len(ref), 2 * rank, replace=False))
    rows = (3 * sel[:, None] + np.arange(3)).ravel()
    q = (bg * np.sqrt(vg))[rows]
    np.testing.assert_allclose(q @ q.T, (va[:, :rg] * wa[:rg]) @ va[:, :rg].T, rtol=1e-5, atol=1e-8 * wa[0])
    # the constructed model is a model: it loads and reconstructs its own instance
    model = api.StatisticalMeshModel(ctx, ref, twin31["cells"], basis, var)
    alpha = rng.normal(0, 0.5, rank)
    got = model.reconstruct(model.theta(alpha))[0]
    np.testing.assert_allclose(got, ref + ((basis * np.sqrt(var)) @ alpha).reshape(-1, 3), rtol=0, atol=1e-9)
    model.close()
    with pytest.raises(Exception):
        core.gpmm_nystrom_extend(ctx, ref, nys, kernel.terms, v, -w)          # non-positive eigenvalues are an error


@pytest.mark.gpu
def test_face_kernel_matches_oracle(ctx):
    """SURVEY 8f rank 4, second kernel family: the symmetrised multiscale B-spline kernel of apps/bfm/FaceKernel.scala on the
    device (icp_gpmm_face_kernel_matrix / icp_gpmm_face_nystrom_extend) against the pair-by-pair oracle, with and without
    face-mask region weights, and the Nystrom model it yields."""
    from oracle import np_oracle as npo
    from icp_proposal_b200 import core
    rng = np.random.default_rng(77)
    levels, scales = api.FaceKernel().levels, api.FaceKernel().scales
    assert levels == [-6, -5, -4, -3, -2] and scales == [128.0, 64.0, 32.0, 10.0, 4.0]            # FaceKernel.scala:60-65
    x, y = rng.uniform(-90, 90, (40, 3)), rng.uniform(-90, 90, (23, 3))
    y[:5] = x[:5] * np.array([-1.0, 1.0, 1.0])                                                        # mirrored partners
    wx, wy, wyb = rng.uniform(0, 1, (5, 40)), rng.uniform(0, 1, (5, 23)), rng.uniform(0, 1, (5, 23))
    for sym, plain, a, b, c in ((0.7, 0.3, None, None, None), (0.7, 0.3, wx, wy, wyb), (0.0, 1.0, wx, wy, None)):
        got = core.gpmm_face_kernel_matrix(ctx, x, y, levels, scales, sym, plain, a, b, c)
        want = npo.face_kernel(x, y, levels, scales, sym, plain, a, b, c)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13 * np.abs(want).max())
    # Nystrom model of a face-sized point cloud through the mirror (region weight = a smooth function of the position)
    pts = rng.uniform(-80, 80, (300, 3))
    nys = pts[::6]
    region = lambda level, p: 0.5 + 0.5 * np.exp(-(p[:, 1] / (40.0 * 2.0 ** (level + 6))) ** 2)      # symmetric in x
    fk = api.FaceKernel(region)
    rank = 20
    kmm = fk.matrix(ctx, nys, nys)
    np.testing.assert_allclose(kmm, kmm.T, rtol=1e-11, atol=1e-12 * np.abs(kmm).max())
    w, v = np.linalg.eigh(kmm)
    order = np.argsort(w)[::-1][:rank]
    w, v = w[order], v[:, order]
    got_b, got_v = core.gpmm_face_nystrom_extend(ctx, pts, nys, levels, scales, v, w, 0.7, 0.3, fk.weights(pts), fk.weights(nys), fk.weights(nys, True))
    kpn = npo.face_kernel(pts[:60], nys, levels, scales, 0.7, 0.3, fk.weights(pts[:60]), fk.weights(nys), fk.weights(nys, True))
    want_b, want_v = npo.nystrom_extend(kpn, v, w)
    np.testing.assert_allclose(got_v, want_v, rtol=1e-14)
    np.testing.assert_allclose(got_b[:180], want_b, rtol=1e-9, atol=1e-12 * np.abs(want_b).max())
    basis, var = api.LowRankGaussianProcess.approximateGPNystrom(ctx, fk, pts, nys, rank)
    np.testing.assert_allclose(var, want_v, rtol=1e-9)
    assert basis.shape == (900, rank) and np.isfinite(basis).all()
