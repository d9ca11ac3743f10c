"""CPU suite for the host side: the C-ABI library loads and exports every symbol include/icpcuda.h declares,
fails loudly without a GPU, and the O(K) host logic of the mirror (mixtures, random-walk / pose densities,
chain log format, sharding + gloo gather) matches the oracle / the reference's formats."""
import ctypes
import json
import math
import os
import re
import socket

import numpy as np
import pytest

from conftest import ROOT
from icp_proposal_b200 import _lib, api, sharding
from oracle import oracle as orc


def _declared_functions():
    # the drop-in boundary (icpcuda.h) and the introspection / micro-benchmark entry points (icpcuda_debug.h)
    src = open(os.path.join(ROOT, "include", "icpcuda.h")).read() + open(os.path.join(ROOT, "include", "icpcuda_debug.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(icp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_functions()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/icpcuda.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(declared)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from icp_proposal_b200 import core
    with pytest.raises(_lib.IcpCudaError) as e:
        core.Context(0)
    assert e.value.code in (_lib.ERR_CUDA, _lib.ERR_INVALID_ARGUMENT)
    lib = _lib.load()
    # compute entry points refuse null handles instead of computing anything on the host
    assert lib.icp_reconstruct(None, 1, None, None) == _lib.ERR_INVALID_ARGUMENT
    assert lib.icp_chain_run(None, 1, 1, None, None) == _lib.ERR_INVALID_ARGUMENT


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "icp-proposal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "icp_oracle", "libicporacle", "oracle/", "np_oracle"):
                    assert needle not in txt, f"{f} reaches into the oracle ({needle})"


def _theta(K, rng):
    th = np.zeros(K + 10); th[0] = 1; th[7:10] = rng.normal(size=3); th[10:] = rng.normal(0, 0.3, K)
    return th


def test_model_fitting_parameters_equality():
    rng = np.random.default_rng(0)
    a = api.ModelFittingParameters.from_vector(_theta(5, rng), "A")
    b = a.copy(generatedBy="B")
    assert a == b and hash(a) == hash(b)          # generatedBy is not part of equality (Appendix B6)
    c = a.copy(shapeParameters=api.ShapeParameters(a.shapeParameters.parameters + 1e-17))
    assert (a == c) == (a.allParameters.tobytes() == c.allParameters.tobytes())
    assert len(a.allParameters) == 15 and a.allParameters[0] == 1.0


def test_random_walk_and_pose_transitions_match_oracle():
    rng = np.random.default_rng(1)
    K = 11

    class M:  # minimal model stand-in for the host-only proposals
        pass
    M.K = K
    frm = api.ModelFittingParameters.from_vector(_theta(K, rng))
    rw = api.RandomShapeUpdateProposal(M, 0.1, rand=rng)
    to = rw.propose(frm)
    np.testing.assert_allclose(rw.logTransitionProbability(frm, to), orc.random_walk_log_transition(K, 0.1, frm.allParameters, to.allParameters), rtol=1e-13)
    rot = api.GaussianAxisRotationProposal(0.01, api.PitchAxis, rand=rng)
    tr = api.GaussianAxisTranslationProposal(0.1, 2, rand=rng)
    t2, t3 = rot.propose(frm), tr.propose(frm)
    assert t2.allParameters[5] != frm.allParameters[5] and np.array_equal(np.delete(t2.allParameters, 5), np.delete(frm.allParameters, 5))
    np.testing.assert_allclose(rot.logTransitionProbability(frm, t2), orc.pose_log_transition(K, 0, 1, 0.01, frm.allParameters, t2.allParameters), rtol=1e-13)
    np.testing.assert_allclose(tr.logTransitionProbability(frm, t3), orc.pose_log_transition(K, 1, 2, 0.1, frm.allParameters, t3.allParameters), rtol=1e-13)
    # guards: a proposal of another family is impossible under this one
    assert rw.logTransitionProbability(frm, t2) == -math.inf
    assert rot.logTransitionProbability(frm, to) == -math.inf and rot.logTransitionProbability(frm, t3) == -math.inf
    assert orc.pose_log_transition(K, 0, 1, 0.01, frm.allParameters, t3.allParameters) == -math.inf


def test_mixture_flatten_and_log_sum_exp():
    rng = np.random.default_rng(2)

    class M:
        pass
    M.K = 7
    rw1, rw2 = api.RandomShapeUpdateProposal(M, 0.1, "a", rng), api.RandomShapeUpdateProposal(M, 0.01, "b", rng)
    inner = api.MixtureProposal([(0.5, rw1), (0.5, rw2)], rng)
    pose = api.MixedProposalDistributions.mixedRandomPoseProposal(rand=rng)
    outer = api.MixtureProposal([(0.9, inner), (0.1, pose)], rng)
    flat = outer.flatten()
    assert [c["name"] for c in flat][:2] == ["a", "b"] and len(flat) == 8
    np.testing.assert_allclose(sum(c["weight"] for c in flat), 1.0, rtol=1e-15)
    np.testing.assert_allclose([c["weight"] for c in flat[:2]], [0.45, 0.45])
    np.testing.assert_allclose([c["weight"] for c in flat[2:]], [0.1 / 6] * 6)
    frm = api.ModelFittingParameters.from_vector(_theta(7, rng))
    to = rw1.propose(frm)
    # nested log-sum-exp equals the flat mixture over leaf components
    ls = [0.45 * math.exp(rw1.logTransitionProbability(frm, to)), 0.45 * math.exp(rw2.logTransitionProbability(frm, to))]
    np.testing.assert_allclose(outer.logTransitionProbability(frm, to), math.log(sum(ls)), rtol=1e-12)
    assert outer.logTransitionRatio(frm, to) == pytest.approx(0.0, abs=1e-12)   # symmetric random walks
    picks = [outer.propose(frm).generatedBy for _ in range(400)]
    assert 0.8 < sum(p in ("a", "b") for p in picks) / 400 < 0.97


def test_densities():
    from scipy import stats
    np.testing.assert_allclose(api.Gaussian(0.1, 0.3).logPdf(0.25), stats.norm(0.1, 0.3).logpdf(0.25), rtol=1e-13)
    np.testing.assert_allclose(api.Exponential(100.0).logPdf(0.02), stats.expon(scale=1 / 100.0).logpdf(0.02), rtol=1e-13)


def test_json_logger_format_roundtrip(tmp_path):
    class Ev:
        def __init__(self, v):
            self.v = v

        def logValue(self, s):
            return self.v + float(s.allParameters[10])
    path = str(tmp_path / "log.json")
    lg = api.JSONAcceptRejectLogger(path, {"product": Ev(1.0), "prior": Ev(2.0), "distance": Ev(3.0)})
    rng = np.random.default_rng(3)
    cur = api.ModelFittingParameters.from_vector(_theta(4, rng), "init")
    new = api.ModelFittingParameters.from_vector(_theta(4, rng), "IcpProposal-ModelSampling-0.1Step")
    lg.accept(cur, new, None, Ev(0))
    lg.reject(new, cur.copy(generatedBy="RandomShape-0.1"), None, Ev(0))
    lg.writeLog()
    raw = json.load(open(path))
    assert list(raw[0].keys()) == ["index", "name", "logvalue", "status", "rigid", "coeff", "datetime"]   # jsonLogFormat, :35
    assert raw[0]["status"] is True and len(raw[0]["rigid"]) == 9 and len(raw[0]["coeff"]) == 4
    assert raw[1]["status"] is False and raw[1]["rigid"] == [] and raw[1]["coeff"] == []                # :101-105
    assert raw[1]["logvalue"]["product"] == pytest.approx(1.0 + new.allParameters[10])                  # the CURRENT state's values
    assert raw[1]["name"] == "RandomShape-0.1" and raw[1]["index"] == 1
    back = lg.sampleToModelParameters(lg.loadLog()[0])
    assert np.array_equal(back.allParameters[1:], new.allParameters[1:])
    assert lg.getBestFittingParsFromJSON() == back
    with pytest.raises(IOError):
        api.JSONAcceptRejectLogger(str(tmp_path / "missing_dir" / "x.json"))
    # device chain log -> same records
    lg2 = api.JSONAcceptRejectLogger(None)
    th = np.stack([new.allParameters, cur.allParameters])
    lg2.append_device_log(["icp", "rw"], ["product", "prior", "distance"], [0, 1], [1, 0], np.arange(6.0).reshape(2, 3), th)
    assert lg2.logStatus[0].coeff == new.allParameters[10:].tolist() and lg2.logStatus[1].coeff == []
    assert lg2.numOfAccepted == 1 and lg2.numOfRejected == 1 and lg2.logStatus[1].logvalue == {"product": 3.0, "prior": 4.0, "distance": 5.0}


def test_native_streaming_json_log_roundtrip(tmp_path):
    """icp_jsonlog_*: the C-boundary writer appends (valid JSON after every append, O(new records) work), writes the reference's
    record layout (JSONAcceptRejectLogger.scala:35,93-106) and its loader reads both its own files and spray-json-style ones."""
    from icp_proposal_b200 import core
    K, C, n = 4, 2, 7
    rng = np.random.default_rng(11)
    run = dict(component=rng.integers(0, 2, (n, C)).astype(np.int32), accepted=rng.random((n, C)) < 0.6,
               values=rng.normal(size=(n, C, 3)), theta=rng.normal(size=(n, C, K + 10)))
    run["accepted"][0, 1] = True
    run["values"][2, 1, 2] = np.nan; run["values"][3, 1, 0] = -np.inf            # spray-json writes null for these
    path = str(tmp_path / "chain.json")
    lg = core.JsonLog(path, K, ["IcpProposal-ModelSampling-0.1Step", "RandomShape-0.1"], ("product", "prior", "collective_distance"))
    assert json.load(open(path)) == []
    first = {k: v[:3] for k, v in run.items()}
    lg.append(first, chain=1)
    size1 = os.path.getsize(path)
    assert len(json.load(open(path))) == 3                                         # valid JSON between appends
    lg.append({k: v[3:] for k, v in run.items()}, chain=1)
    lg.close()
    raw = json.load(open(path))
    assert len(raw) == n and os.path.getsize(path) > size1
    assert list(raw[0].keys()) == ["index", "name", "logvalue", "status", "rigid", "coeff", "datetime"]
    for s_, r in enumerate(raw):
        ok = bool(run["accepted"][s_, 1])
        assert r["index"] == s_ and r["status"] is ok
        assert r["name"] == ["IcpProposal-ModelSampling-0.1Step", "RandomShape-0.1"][run["component"][s_, 1]]
        assert r["rigid"] == (run["theta"][s_, 1, 1:10].tolist() if ok else []) and r["coeff"] == (run["theta"][s_, 1, 10:].tolist() if ok else [])
        for k, key in enumerate(("product", "prior", "collective_distance")):
            v = run["values"][s_, 1, k]
            assert r["logvalue"][key] == (v if np.isfinite(v) else None)
    back = core.jsonlog_load(path, K)
    assert back["value_keys"] == ["product", "prior", "collective_distance"] and np.array_equal(back["status"], run["accepted"][:, 1])
    np.testing.assert_array_equal(np.nan_to_num(back["values"], nan=7.0, neginf=7.0), np.where(np.isfinite(run["values"][:, 1]), run["values"][:, 1], 7.0))
    acc = run["accepted"][:, 1]
    np.testing.assert_array_equal(back["theta"][acc, 1:], run["theta"][acc, 1, 1:])
    assert np.all(back["theta"][acc, 0] == 1.0) and np.isnan(back["theta"][~acc]).all()
    # a file as the Python mirror (spray-json layout) writes it loads the same way
    pl = api.JSONAcceptRejectLogger(str(tmp_path / "py.json"))
    pl.append_device_log(["a", "b"], ["product", "prior", "distance"], run["component"][:, 0], run["accepted"][:, 0], run["values"][:, 0], run["theta"][:, 0])
    pl.writeLog()
    b2 = core.jsonlog_load(str(tmp_path / "py.json"), K)
    assert np.array_equal(b2["status"], run["accepted"][:, 0]) and b2["names"][0] == ["a", "b"][run["component"][0, 0]]
    # LogHelper.samplesFromLog indices (apps/util/LogHelper.scala:27-37)
    st = np.array([1, 0, 0, 1, 0, 1, 0, 0, 0, 1], bool)
    assert core.chainlog_sample_indices(st, take_every_n=2, total=100, burn_in=1).tolist() == [0, 3, 5, 5, 9]
    assert core.chainlog_sample_indices(st, take_every_n=3, total=7, burn_in=0).tolist() == [0, 3, 5]
    with pytest.raises(_lib.IcpCudaError):
        core.chainlog_sample_indices(np.array([0, 0, 1], bool), take_every_n=1)    # no accepted record at or before index 0


def test_log_helper_samples_from_log_matches_oracle():
    """LogHelper.samplesFromLog (apps/util/LogHelper.scala:27-37): the host mirror against the oracle's index walk."""
    from oracle import np_oracle as npo
    rng = np.random.default_rng(11)
    status = [True] + [bool(b) for b in rng.random(199) < 0.4]
    log = [api.jsonLogFormat(i, "p", {"product": 0.0}, st, [0.0] * 9 if st else [], [float(i)] if st else [], "") for i, st in enumerate(status)]
    for every, total, burn in ((50, 100, 0), (7, 150, 20), (1, 30, 5), (13, 10000, 199)):
        got = api.LogHelper.samplesFromLog(log, takeEveryN=every, total=total, burnIn=burn)
        assert [j for _, j in got] == npo.samples_from_log(status, every, total, burn)
        assert all(l.status and l.index == j for l, j in got)
    th = api.LogHelper.logSamples2thetas([l for l, _ in api.LogHelper.samplesFromLog(log, 7, 150, 20)])
    assert th.shape[1] == 11 and (th[:, 0] == 1.0).all()
    with pytest.raises(IndexError):
        api.LogHelper.samplesFromLog([api.jsonLogFormat(0, "p", {}, False, [], [], "")], 1, 1, 0)


def test_kernel_algebra_builds_the_femur_kernel_terms():
    """The Scalismo-style kernel algebra of the GPMM-construction mirror (apps/femur/CreateGPModel.scala:70-83) reduces to
    the (scale, sigma, A) terms the device evaluates; checked against the oracle's pair-by-pair kernel written out by hand."""
    from oracle import np_oracle as npo
    rng = np.random.default_rng(2)
    pts = rng.normal(0, 40, (30, 3)) * np.array([1.0, 0.8, 5.0])
    k = api.femurKernel(pts)
    d = api.getAxisOfMainVariance(pts)
    base = d @ np.diag([10.0, 1.0, 1.0]) @ d.T
    assert [(s, sg) for s, sg, _ in k.terms] == [(10.0, 90.0), (5.0, 40.0), (3.0, 10.0)]
    np.testing.assert_allclose(k.terms[0][2], base, rtol=1e-14)
    assert k.terms[1][2] is None and k.terms[2][2] is None
    np.testing.assert_allclose(base, base.T, atol=1e-12)
    assert abs(abs(d[2, 0]) - 1.0) < 0.05                      # the long axis carries the 10x variance
    x, y = pts[:4], pts[4:9]
    want = np.zeros((12, 15))
    for i in range(4):
        for j in range(5):
            r2 = ((x[i] - y[j]) ** 2).sum()
            want[3 * i:3 * i + 3, 3 * j:3 * j + 3] = base * 10.0 * np.exp(-r2 / 90.0 ** 2) + np.eye(3) * (5.0 * np.exp(-r2 / 40.0 ** 2) + 3.0 * np.exp(-r2 / 10.0 ** 2))
    np.testing.assert_allclose(npo.gauss_mixture_kernel(x, y, k.terms), want, rtol=1e-13, atol=1e-300)
    k2 = (api.GaussianKernel3D(20) * 2.0 + api.DiagonalKernel3D(api.GaussianKernel3D(5), 3)) * 0.5
    assert [(s, sg) for s, sg, _ in k2.terms] == [(1.0, 20.0), (0.5, 5.0)]
    with pytest.raises(ValueError):
        api.DiagonalKernel3D(api.GaussianKernel3D(5), 2)


def test_shard_ranges_cover_all_chains():
    for n, w in ((100, 8), (5, 2), (7, 4), (3, 8), (1184 * 8, 8)):
        spans = [sharding.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    K, n, steps = 6, 10, 4
    theta = rng.normal(size=(n, K + 10)); logv = rng.normal(size=(steps, n, 3))
    lo, hi = sharding.shard_range(n, rank, world)
    local = sharding.chain_statistics(theta[lo:hi], logv[:, lo:hi])
    got = sharding.combine_statistics(sharding.gather_statistics(local))
    want = sharding.combine_statistics(sharding.chain_statistics(theta, logv)[None])
    ok = got["n"] == want["n"] and np.allclose(got["mean"], want["mean"], rtol=1e-13) and \
        np.allclose(got["var"], want["var"], rtol=1e-10) and np.isclose(got["mean_logp"], want["mean_logp"], rtol=1e-13)
    # posterior variability maps: every rank reduces its own samples, the gathered centred moments merge to the unsharded maps
    from oracle import np_oracle as npo
    S, nv = 11, 12
    verts = rng.normal(0, 50, (nv, 3)) + 200.0
    tris = np.array([[0, 1, 2], [2, 3, 4], [4, 5, 6], [6, 7, 8], [8, 9, 10], [10, 11, 0], [1, 3, 5], [7, 9, 11]])
    meshes = [verts + rng.normal(0, 0.3, verts.shape) for _ in range(S)]
    slo, shi = sharding.shard_range(S, rank, world)
    m_loc, c_loc, _, _ = npo.posterior_variability(meshes[slo:shi], tris, ref_verts=verts, sum_normals=False)
    parts = sharding.gather_statistics(sharding.variability_partials(shi - slo, m_loc, c_loc))
    nrm = npo.vertex_normals(verts, tris)
    merged = sharding.merge_variability(parts, normals=nrm)
    m_all, c_all, t_all, a_all = npo.posterior_variability(meshes, tris, ref_verts=verts, sum_normals=False)
    ok = ok and merged["n"] == S and np.allclose(merged["mean"], m_all, rtol=1e-14) and np.allclose(merged["cov"], c_all, rtol=1e-9, atol=1e-14) \
        and np.allclose(merged["total_variance"], t_all, rtol=1e-10) and np.allclose(merged["normal_variance"], a_all, rtol=1e-9)
    q.put((rank, bool(ok), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_statistics_gather_gloo_world2():
    """N > 1 path on CPU: two ranks shard the chains, all_gather their statistics, both get the unsharded answer."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2:] == (0, 5) and res[1][2:] == (5, 10)


def test_bench_init_thetas_are_sharding_independent():
    import bench
    m = {"ref": np.zeros((4, 3))}
    full = bench.init_thetas(m, 6, 0)
    parts = np.concatenate([bench.init_thetas(m, 2, 0), bench.init_thetas(m, 4, 2)])
    assert np.array_equal(full, parts) and not full[0, 10:].any() and full[1, 10:].any()
