"""Regenerates the golden fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the GPU box only ever reads the committed files).

Inputs  : the reference's own data fixtures (data/femur/*.h5, *.stl, *.json), converted to .npz so
          that they can travel: femur reference mesh, landmark-aligned target
          (apps/femur/LoadTestData.scala:32-50) and the 50-/100-/200-component GPMMs (rank 51 / 101 / 201).
Outputs : oracle values on those inputs (oracle/icp_oracle.c, cross-checked against
          oracle/np_oracle.py) for fixed parameter vectors: femur_golden.json.
PARITY UNPINNED: the reference ships no golden vectors for this path; these pin the
oracle itself across machines and compilers.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from icp_proposal_b200 import fixtures_io as fx  # noqa: E402
from oracle import oracle as orc  # noqa: E402

REF = "/root/reference/data/femur"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    m100 = fx.load_gpmm_h5(f"{REF}/femur_gp_model_100-components.h5")
    m50 = fx.load_gpmm_h5(f"{REF}/femur_gp_model_50-components.h5")
    m200 = fx.load_gpmm_h5(f"{REF}/femur_gp_model_200-components.h5")   # rank 201: the model StdIcpVsChainICP...All.scala:88 opens
    tv, tc = fx.read_binary_stl(f"{REF}/femur_target.stl")
    r, t = fx.rigid_landmark_alignment(fx.read_landmarks_json(f"{REF}/femur_target.json"),
                                       fx.read_landmarks_json(f"{REF}/femur_reference.json"))
    tv_aligned = tv @ r.T + t
    np.savez_compressed(f"{OUT}/femur_meshes.npz", ref=m100["ref"].astype(np.float32), cells=m100["cells"],
                        target_raw=tv.astype(np.float32), target_cells=tc, target_aligned=tv_aligned,
                        align_R=r, align_t=t)
    for name, m in (("femur_gpmm_50", m50), ("femur_gpmm_100", m100), ("femur_gpmm_200", m200)):
        np.savez_compressed(f"{OUT}/{name}.npz", basis=m["basis"].astype(np.float32),
                            variance=m["variance"].astype(np.float32))
    golden = {}
    for name, m in (("gpmm_50", m50), ("gpmm_100", m100), ("gpmm_200", m200)):
        K = len(m["variance"])
        om = orc.Model(m["ref"], m["cells"], m["basis"], m["variance"])
        ot = orc.Mesh(tv_aligned, tc)
        rng = np.random.default_rng(2020)
        theta = np.zeros(K + 10); theta[0] = 1.0; theta[7:10] = m["ref"].mean(0)
        theta[10:] = rng.normal(0, 0.3, K)
        ids = np.arange(2 * K)
        tp = tv_aligned[:: max(1, len(tv_aligned) // (2 * K))][: 2 * K]
        eids = np.arange(4 * K)
        g = {"theta": theta.tolist(), "n_icp": int(2 * K), "n_eval": int(4 * K)}
        for direction, key in ((0, "model_sampling"), (1, "target_sampling")):
            p = orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, direction, True, ids, tp)
            post = p.posterior(theta)
            to = p.propose(theta, np.zeros(K))
            g[key] = {"n_obs": int(post["n"]), "mu": post["mu"].tolist(), "M_diag": np.diag(post["M"]).tolist(),
                      "M_fro": float(np.linalg.norm(post["M"])), "propose_z0": to[10:].tolist(),
                      "log_transition_to_z0": float(p.log_transition(theta, to))}
        g["independent_m2t"] = orc.eval_independent(om, ot, 0, 0.0, 2.0, eids, tp, theta)
        g["independent_sym"] = orc.eval_independent(om, ot, 2, 0.0, 2.0, eids, tp, theta)
        g["hausdorff"] = orc.eval_hausdorff(om, ot, 100.0, theta)
        g["collective_sym"] = orc.eval_collective(om, ot, 2, 0.1, 0.3, 1.0, eids, tp, theta)[0]
        g["prior"] = orc.eval_prior(K, theta)
        xyz = om.transformed_mesh(theta)
        g["mesh_checksum"] = [float(xyz.sum()), float(np.abs(xyz).sum())]
        golden[name] = g
    # posterior variability maps of five fixed samples of the 50-component model, and the femur kernel of
    # apps/femur/CreateGPModel.scala:70-83 on the first reference points (oracle/np_oracle.py)
    from oracle import np_oracle as npo
    K50 = len(m50["variance"])
    om50 = orc.Model(m50["ref"], m50["cells"], m50["basis"], m50["variance"])
    rng = np.random.default_rng(77)
    thetas = np.zeros((5, K50 + 10)); thetas[:, 0] = 1.0; thetas[:, 7:10] = m50["ref"].mean(0)
    thetas[:, 10:] = rng.normal(0, 0.4, (5, K50))
    meshes = [om50.transformed_mesh(t) for t in thetas]
    mean, cov, total, along = npo.posterior_variability(meshes, m50["cells"], sum_normals=True)
    golden["variability_gpmm_50"] = {"thetas": thetas.tolist(), "mean_checksum": [float(mean.sum()), float(np.abs(mean).sum())],
                                     "total_first8": total[:8].tolist(), "normal_first8": along[:8].tolist(),
                                     "total_sum": float(total.sum()), "normal_sum": float(along.sum())}
    pts = m50["ref"][:12]
    c = m50["ref"] - m50["ref"].mean(0)
    u = np.linalg.svd(c.T @ c / len(c))[0]
    base = u @ np.diag([10.0, 1.0, 1.0]) @ u.T
    terms = [(10.0, 90.0, base), (5.0, 40.0, None), (3.0, 10.0, None)]
    kk = npo.gauss_mixture_kernel(pts, pts, terms)
    w = np.linalg.eigvalsh(kk)[::-1]
    golden["femur_kernel"] = {"base_matrix": base.tolist(), "k_checksum": [float(kk.sum()), float(np.abs(kk).sum())],
                              "k_row0_first9": kk[0, :9].tolist(), "eigenvalues_first6": w[:6].tolist()}
    with open(f"{OUT}/femur_golden.json", "w") as f:
        json.dump(golden, f, indent=1)
    print({k: os.path.getsize(f"{OUT}/{k}") for k in sorted(os.listdir(OUT))})


if __name__ == "__main__":
    main()
