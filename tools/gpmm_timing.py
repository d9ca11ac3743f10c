"""Times the device GPMM construction (kernel matrix -> Jacobi eigensolver -> Nystrom extension) on the femur recipe of
apps/femur/CreateGPModel.scala (2 K Nystrom points, K + 1 basis functions) and compares with LAPACK on the host."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp_proposal_b200 import api, core, synth  # noqa: E402

ctx = core.Context(0)
verts, tris = synth.fibonacci_ellipsoid_mesh(1622)
kernel = api.femurKernel(verts)
for comps in (50, 100, 200):
    rank, m = comps + 1, 2 * comps
    nys = verts[np.sort(np.random.default_rng(1024).choice(len(verts), m, replace=False))]
    t0 = time.perf_counter(); kmm = core.gpmm_kernel_matrix(ctx, nys, nys, kernel.terms); t1 = time.perf_counter()
    w, v = core.gpmm_eigen_psd(ctx, kmm, rank); t2 = time.perf_counter()
    basis, var = core.gpmm_nystrom_extend(ctx, verts, nys, kernel.terms, v, w); t3 = time.perf_counter()
    wl = np.linalg.eigvalsh(kmm)[::-1][:rank]; t4 = time.perf_counter()
    print(f"{comps:3d} components: kernel matrix {1e3 * (t1 - t0):7.1f} ms, Jacobi {3 * m}x{3 * m} {1e3 * (t2 - t1):7.1f} ms, extension {1e3 * (t3 - t2):7.1f} ms "
          f"(host LAPACK eigvalsh {1e3 * (t4 - t3):6.1f} ms); max rel eigenvalue error {np.abs(w / wl - 1).max():.1e}; "
          f"variance kept {var.sum():.1f}", flush=True)
