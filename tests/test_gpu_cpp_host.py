"""The C++ host-side mirror (icp-proposal_b200/host/icp_host.hpp) and the example that mirrors
apps/femur/IcpProposalRegistration.scala: compiled with g++ against libicpcuda.so, run on the GPU, and compared with the
same chain driven through the Python binding (same Philox seed -> identical device chain)."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from icp_proposal_b200 import _lib, core

pytestmark = pytest.mark.gpu


def write_model_bin(path, m, K, ids, eids, tp):
    """The binary the C++ examples read: sizes, then the double arrays, then the int32 arrays."""
    with open(path, "wb") as f:
        np.array([len(m["ref"]), len(m["cells"]), K, len(m["target"]), len(m["target_cells"]), len(ids), len(tp), len(eids)], np.int32).tofile(f)
        for a in (m["ref"], m["basis"], m["variance"], m["target"], tp):
            np.ascontiguousarray(a, np.float64).tofile(f)
        for a in (m["cells"], m["target_cells"], ids, eids):
            np.ascontiguousarray(a, np.int32).tofile(f)


def test_cpp_threads_on_shared_handles(ctx, twin31, tmp_path):
    """examples/per_call_threads.cpp: ten std::threads (no GIL - what JVM threads would see) drive their own Metropolis-Hastings
    walks through the per-call C ABI on ONE shared pair of proposals and ONE shared evaluator
    (RunMHRandomInitComparison.scala:59-86); every walk must end exactly where it ends when it runs alone."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    m = twin31
    K = 31
    ids = np.arange(2 * K, dtype=np.int32)
    eids = np.arange(4 * K, dtype=np.int32)
    tp = np.ascontiguousarray(m["target"][::26][:2 * K])
    path = tmp_path / "model.bin"
    write_model_bin(path, m, K, ids, eids, tp)
    exe = tmp_path / "per_call_threads"
    libdir = os.path.join(ROOT, "icp-proposal_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "per_call_threads.cpp"), "-o", str(exe), "-L", libdir, "-licpcuda", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe), str(path), "10", "25"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["threads"] == 10 and res["failed_threads"] == 0 and res["walks_that_differ_from_the_serial_run"] == 0
    assert 0 < res["accepted_total"] < 250


def test_cpp_host_mirror_example(ctx, twin31, tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    m = twin31
    K = 31
    ids = np.arange(2 * K, dtype=np.int32)
    eids = np.arange(4 * K, dtype=np.int32)
    tp = np.ascontiguousarray(m["target"][::26][:2 * K])
    path = tmp_path / "model.bin"
    write_model_bin(path, m, K, ids, eids, tp)
    exe = tmp_path / "icp_proposal_registration"
    libdir = os.path.join(ROOT, "icp-proposal_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(libdir, "host"),
                           os.path.join(ROOT, "examples", "icp_proposal_registration.cpp"), "-o", str(exe), "-L", libdir, "-licpcuda",
                           f"-Wl,-rpath,{libdir}"])
    log = tmp_path / "log.json"
    n = 200
    out = subprocess.run([str(exe), str(path), str(n), str(log)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["K"] == K and res["fused_steps"] == n
    # per-call Metropolis-Hastings over the drop-in classes moves uphill
    assert 0 < res["host_mh_accepted"] <= 20 and res["product_after_host_mh"] > res["product_initial"]
    # best sample of the fused run: its logged product equals the value recomputed through the per-call evaluators
    np.testing.assert_allclose(res["fused_best_product"], res["fused_best_product_recomputed"], rtol=1e-9)
    # the same chain through the Python binding (same seed): identical device chain
    model = core.Model(ctx, m["ref"], m["cells"], m["basis"], m["variance"])
    tgt = core.Target(ctx, m["target"], m["target_cells"])
    pt = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.TARGET_SAMPLING, True, ids, tp)
    pm = core.IcpProposal(model, tgt, 0.1, 10.0, 5.0, _lib.MODEL_SAMPLING, True, ids, tp)
    comps = [dict(kind=0, weight=0.45, proposal=pt), dict(kind=0, weight=0.45, proposal=pm), dict(kind=1, weight=0.1, sd=0.1)]
    ev = core.Evaluator(model, tgt, _lib.EVAL_INDEPENDENT, 0, True, 0.0, 2.0, 0.0, eids, tp)
    chain = core.Chain(model, tgt, comps, ev, max_chains=1)
    ref = chain.run(model.theta(), n, seed=1024)
    assert int(ref["n_accepted"][0]) == res["fused_accepted"]
    best = np.where(ref["accepted"][:, 0], ref["values"][:, 0, 0], -np.inf).max()
    np.testing.assert_allclose(res["fused_best_product"], best, rtol=1e-12)
    # the JSON chain log has the reference's record layout
    entries = json.load(open(log))
    assert len(entries) == n and set(entries[0]) == {"index", "name", "logvalue", "status", "rigid", "coeff", "datetime"}
    acc = [e for e in entries if e["status"]]
    rej = [e for e in entries if not e["status"]]
    assert len(acc) == res["fused_accepted"] and all(len(e["coeff"]) == K and len(e["rigid"]) == 9 for e in acc)
    assert all(e["coeff"] == [] and e["rigid"] == [] for e in rej)
    assert {e["name"] for e in entries} <= {"IcpProposal-TargetSampling-0.1Step", "IcpProposal-ModelSampling-0.1Step", "RandomShape-0.1"}
    # posterior variability of every 10th logged state after a burn-in of 20 (LogHelper.samplesFromLog walks rejected
    # entries back to the last accepted one): C++ mirror == Python binding on the same chain
    from oracle import np_oracle as npo
    idx = npo.samples_from_log([bool(a) for a in ref["accepted"][:, 0]], 10, n, 20)
    assert res["variability_samples"] == len(idx) and len(idx) == 18
    pv = core.posterior_variability(model, ref["theta"][idx, 0], True)
    np.testing.assert_allclose(res["mean_total_variance"], pv["total_variance"].mean(), rtol=1e-9)
    np.testing.assert_allclose(res["mean_normal_variance"], pv["normal_variance"].mean(), rtol=1e-9)
    assert 0 < res["mean_normal_variance"] <= res["mean_total_variance"]
    # GPMM construction through the C++ mirror == through the Python mirror (same device calls)
    from icp_proposal_b200 import api
    kern = api.DiagonalKernel3D(api.GaussianKernel3D(40.0), 3) * 5.0 + api.DiagonalKernel3D(api.GaussianKernel3D(10.0), 3) * 3.0
    basis, var = api.LowRankGaussianProcess.approximateGPNystrom(ctx, kern, m["ref"], m["ref"][::20], 8)
    np.testing.assert_allclose(res["gp_variance_sum"], var.sum(), rtol=1e-10)
    np.testing.assert_allclose(res["gp_basis_abs_sum"], np.abs(basis).sum(), rtol=1e-9)
    chain.close(); ev.close(); model.close(); tgt.close()
