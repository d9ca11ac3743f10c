"""Split-precision (Ozaki) emulation of the ICP posterior's rank update M = I + A^T A on INT8 tensor cores: how many 7-bit
slices of A are needed for the 1e-5 contract on the posterior mean, on the reference's femur GPMM-100 (CPU study, numpy).

A (3n x K) = whitened basis rows F_i Q_i. Every column of A is scaled by a power of two into (-1, 1) and cut into s
slices of 7 bits (int8 operands); the products S_k^T S_l with k + l < s are exact in int32 (606 rows x 64^2 < 2^31) and are
combined in FP64. Reports the error of M and of mu = M^-1 b against the FP64 result.
    python tools/ozaki_study.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc  # noqa: E402  (a study tool, not the product path)


def slices(A, s, bits=7):
    e = np.ceil(np.log2(np.abs(A).max(0) + 1e-300)) + 1          # |A[:, j]| * 2^-e_j < 0.5
    At = A * 2.0 ** (-e)[None, :]
    out, r = [], At.copy()
    for k in range(s):
        q = np.rint(r * 2.0 ** (bits * (k + 1)))                 # |q| <= 64 for k = 0, <= 64 afterwards (|r| <= 2^-(7k+1))
        assert np.abs(q).max() <= 64
        out.append(q.astype(np.int64))
        r = r - q * 2.0 ** (-bits * (k + 1))
    return out, e


def ozaki_gram(A, s, bits=7):
    S, e = slices(A, s, bits)
    K = A.shape[1]
    G = np.zeros((K, K))
    n_gemm = 0
    for k in range(s):
        for l in range(s - k):
            P = S[k].T @ S[l]                                    # exact integers
            assert np.abs(P).max() < 2 ** 31
            G += P.astype(np.float64) * 2.0 ** (-bits * (k + l + 2))
            n_gemm += 1
    return G * 2.0 ** (e[:, None] + e[None, :]), n_gemm


def main():
    g = os.path.join(ROOT, "tests", "golden")
    meshes = np.load(os.path.join(g, "femur_meshes.npz"))
    gp = np.load(os.path.join(g, "femur_gpmm_100.npz"))
    ref, cells = meshes["ref"].astype(float), meshes["cells"].astype(np.int32)
    tgt, tcells = meshes["target_aligned"].astype(float), meshes["target_cells"].astype(np.int32)
    basis, var = gp["basis"].astype(float), gp["variance"].astype(float)
    K = len(var)
    om, ot = orc.Model(ref, cells, basis, var), orc.Mesh(tgt, tcells)
    Q = basis * np.sqrt(var)[None, :]
    rng = np.random.default_rng(0)
    rows = []
    for direction in (0, 1):
        for trial in range(3):
            th = np.zeros(K + 10); th[0] = 1; th[7:10] = ref.mean(0); th[10:] = rng.normal(0, 0.3 if trial else 0.0, K)
            ids = np.arange(2 * K); tp = tgt[:: len(tgt) // (2 * K)][: 2 * K]
            p = orc.IcpProposal(om, ot, 0.1, 10.0, 5.0, direction, True, ids, tp)
            po = p.posterior(th, with_obs=True)
            A, y = [], []
            for vid, yy, cov in zip(po["ids"], po["y"], po["cov"]):
                w, V = np.linalg.eigh(np.linalg.inv(cov))
                F = (V * np.sqrt(w)).T                           # Sigma^-1 = F^T F
                A.append(F @ Q[3 * vid: 3 * vid + 3]); y.append(F @ yy)
            A, y = np.concatenate(A), np.concatenate(y)
            M = np.eye(K) + A.T @ A
            b = A.T @ y
            mu = np.linalg.solve(M, b)
            assert np.allclose(M, po["M"], rtol=1e-9) and np.allclose(mu, po["mu"], rtol=1e-6, atol=1e-9)
            for s in (3, 4, 5, 6, 7, 8):
                G, n_gemm = ozaki_gram(A, s)
                Mo = np.eye(K) + G
                muo = np.linalg.solve(Mo, b)
                rows.append(dict(direction=direction, trial=trial, slices=s, int8_gemms=n_gemm, cond_M=float(np.linalg.cond(M)),
                                 M_rel_err=float(np.abs(Mo - M).max() / np.abs(M).max()),
                                 mu_rel_err=float(np.abs(muo - mu).max() / np.abs(mu).max())))
    agg = {}
    for r in rows:
        a = agg.setdefault(r["slices"], dict(slices=r["slices"], int8_gemms=r["int8_gemms"], M_rel_err=0.0, mu_rel_err=0.0))
        a["M_rel_err"] = max(a["M_rel_err"], r["M_rel_err"]); a["mu_rel_err"] = max(a["mu_rel_err"], r["mu_rel_err"])
    print(json.dumps(dict(model="femur GPMM-100 (K = 101), n = 202 observations, both projection directions, 3 states each",
                          cond_M_max=max(r["cond_M"] for r in rows), worst_case_by_slices=list(agg.values())), indent=1))


if __name__ == "__main__":
    main()
