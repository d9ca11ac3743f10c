"""Second, independent CPU restatement in numpy/LAPACK (FP64) used to cross-check the C oracle.

TEST INFRASTRUCTURE ONLY (same rules as icp_oracle.h). PARITY UNPINNED. It follows the same
reference call sites but shares no code with icp_oracle.c: closest points are brute force and
vectorised, the linear algebra is LAPACK (numpy.linalg.svd / pinv) exactly where Breeze uses
``svd`` / ``pinv``. Paths relative to /root/reference/src/main/scala.
"""
from __future__ import annotations

import numpy as np

LOG_2PI = float(np.log(2.0 * np.pi))


# ---- geometry ---------------------------------------------------------------------------------
def closest_point_on_triangles(q, verts, tris):
    """Brute force exact closest point (Ericson, Real-Time Collision Detection 5.1.5), vectorised over
    (queries x triangles). Returns (tri, cp, d2)."""
    q = np.asarray(q, float).reshape(-1, 1, 3)
    a, b, c = verts[tris[:, 0]][None], verts[tris[:, 1]][None], verts[tris[:, 2]][None]
    ab, ac, ap = b - a, c - a, q - a
    d1, d2 = (ab * ap).sum(-1), (ac * ap).sum(-1)
    bp = q - b
    d3, d4 = (ab * bp).sum(-1), (ac * bp).sum(-1)
    cp_ = q - c
    d5, d6 = (ab * cp_).sum(-1), (ac * cp_).sum(-1)
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    with np.errstate(divide="ignore", invalid="ignore"):
        denom = 1.0 / (va + vb + vc)
        v, w = vb * denom, vc * denom
        res = a + ab * v[..., None] + ac * w[..., None]
        # regions, applied in reverse priority so that earlier tests win
        m = (va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0)
        wbc = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        res = np.where(m[..., None], b + wbc[..., None] * (c - b), res)
        m = (vb <= 0) & (d2 >= 0) & (d6 <= 0)
        wac = d2 / (d2 - d6)
        res = np.where(m[..., None], a + wac[..., None] * ac, res)
        m = (d6 >= 0) & (d5 <= d6)
        res = np.where(m[..., None], np.broadcast_to(c, res.shape), res)
        m = (vc <= 0) & (d1 >= 0) & (d3 <= 0)
        vab = d1 / (d1 - d3)
        res = np.where(m[..., None], a + vab[..., None] * ab, res)
        m = (d3 >= 0) & (d4 <= d3)
        res = np.where(m[..., None], np.broadcast_to(b, res.shape), res)
        m = (d1 <= 0) & (d2 <= 0)
        res = np.where(m[..., None], np.broadcast_to(a, res.shape), res)
    dd = ((q - res) ** 2).sum(-1)
    t = dd.argmin(1)
    idx = np.arange(len(t))
    return t, res[idx, t], dd[idx, t]


def closest_vertex(q, verts):
    d = ((np.asarray(q, float).reshape(-1, 1, 3) - verts[None]) ** 2).sum(-1)
    return d.argmin(1)


def boundary_flags(nv, tris):
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
    e.sort(axis=1)
    uniq, cnt = np.unique(e, axis=0, return_counts=True)
    f = np.zeros(nv, bool)
    f[uniq[cnt == 1].ravel()] = True
    return f


def vertex_normals(verts, tris):
    cn = np.cross(verts[tris[:, 1]] - verts[tris[:, 0]], verts[tris[:, 2]] - verts[tris[:, 0]])
    cn /= np.linalg.norm(cn, axis=1, keepdims=True)
    s = np.zeros_like(verts)
    cnt = np.zeros(len(verts))
    for k in range(3):
        np.add.at(s, tris[:, k], cn)
        np.add.at(cnt, tris[:, k], 1)
    s /= cnt[:, None]
    return s / np.linalg.norm(s, axis=1, keepdims=True)


# ---- model ------------------------------------------------------------------------------------
def pose_matrix(theta):
    phi, th, psi = theta[4], theta[5], theta[6]
    rz = np.array([[np.cos(phi), -np.sin(phi), 0], [np.sin(phi), np.cos(phi), 0], [0, 0, 1]])
    ry = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    rx = np.array([[1, 0, 0], [0, np.cos(psi), -np.sin(psi)], [0, np.sin(psi), np.cos(psi)]])
    return rz @ ry @ rx


class Model:
    def __init__(self, ref, tris, basis, variance, mean_def=None):
        self.ref = np.asarray(ref, float).reshape(-1, 3)
        self.tris = np.asarray(tris).reshape(-1, 3)
        self.U = np.asarray(basis, float)
        self.var = np.asarray(variance, float)
        self.Q = self.U * np.sqrt(self.var)
        self.N, self.K = len(self.ref), len(self.var)
        self.mean_def = np.zeros(3 * self.N) if mean_def is None else np.asarray(mean_def, float).reshape(-1)

    def transformed_mesh(self, theta):
        """api/sampling/ModelFittingParameters.scala:93-110"""
        theta = np.asarray(theta, float)
        p = self.ref + (self.mean_def + self.Q @ theta[10:]).reshape(-1, 3)
        r, t, c = pose_matrix(theta), theta[1:4], theta[7:10]
        return theta[0] * ((p - c) @ r.T + c + t)


def surface_noise_cov(normal, sd_n, sd_t):
    """api/sampling/SurfaceNoiseHelpers.scala:32-60"""
    n = np.asarray(normal, float) / np.linalg.norm(normal)
    cand = np.cross(n, [1.0, 0, 0])
    t1 = cand if cand @ cand < 1e-4 else np.cross(n, [0, 1.0, 0])
    with np.errstate(invalid="ignore", divide="ignore"):
        t1 = t1 / np.linalg.norm(t1)
        t2 = np.cross(n, t1)
        t2 = t2 / np.linalg.norm(t2)
    phi = np.stack([n, t1, t2], axis=1)
    return phi @ np.diag([sd_n ** 2, sd_t ** 2, sd_t ** 2]) @ phi.T


def regression(Qs, covs, resid):
    """Appendix A3. Qs (3n, K), covs (n, 3, 3), resid (3n,) -> (M, Minv, coefficients)."""
    n = len(covs)
    K = Qs.shape[1]
    QtL = Qs.T.copy()
    for i in range(n):
        QtL[:, 3 * i:3 * i + 3] = QtL[:, 3 * i:3 * i + 3] @ np.linalg.inv(covs[i])
    M = QtL @ Qs + np.eye(K)
    u, s, vt = np.linalg.svd(M)
    Minv = (u * np.where(s == 0, 0.0, 1.0 / s)) @ vt
    Minv = Minv.T
    return M, Minv, (Minv @ QtL) @ resid


def icp_observations(model, target_verts, target_tris, theta, direction, boundary_aware, ids, target_points, sd_t, sd_n):
    """api/sampling/proposals/NonRigidIcpProposal.scala:88-149"""
    theta = np.asarray(theta, float)
    cur = model.transformed_mesh(theta)
    nrm = vertex_normals(cur, model.tris)
    r, t, c = pose_matrix(theta), theta[1:4], theta[7:10]
    inv_pose = lambda x: (x - t - c) @ r + c
    obs = []
    if direction == 1:
        cur_b = boundary_flags(model.N, model.tris)
        vid = closest_vertex(target_points, cur)
        for p, i in zip(np.asarray(target_points).reshape(-1, 3), vid):
            if boundary_aware and cur_b[i]:
                continue
            obs.append((i, inv_pose(p) - model.ref[i], surface_noise_cov(nrm[i], sd_n, sd_t)))
    else:
        tb = boundary_flags(len(target_verts), target_tris)
        _, cp, _ = closest_point_on_triangles(cur[ids], target_verts, target_tris)
        tv = closest_vertex(cp, target_verts)
        for i, p, v in zip(ids, cp, tv):
            if boundary_aware and tb[v]:
                continue
            obs.append((i, inv_pose(p) - model.ref[i], surface_noise_cov(nrm[i], sd_n, sd_t)))
    return obs


def icp_posterior(model, obs):
    ids = np.array([o[0] for o in obs], int)
    rows = (3 * ids[:, None] + np.arange(3)).ravel()
    y = np.concatenate([o[1] for o in obs]) if obs else np.zeros(0)
    covs = np.array([o[2] for o in obs]).reshape(-1, 3, 3)
    M, Minv, mu = regression(model.Q[rows], covs, y - model.mean_def[rows])
    return dict(mu=mu, M=M, Minv=Minv)


def _full_coefficients(Qfull, resid):
    """model.coefficients: all N points, noise 1e-5 I (Appendix A5)"""
    K = Qfull.shape[1]
    QtL = Qfull.T / 1e-5
    M = QtL @ Qfull + np.eye(K)
    return (np.linalg.pinv(M, rcond=0.0, hermitian=False) @ QtL) @ resid


def posterior_basis(model, post):
    d = np.sqrt(model.var)
    sigma = d[:, None] * post["Minv"] * d[None, :]
    ubar, lam, _ = np.linalg.svd(sigma)
    # documented sign rule shared with the C oracle and the device (icp_oracle.c posterior_basis): the entry of largest
    # magnitude of every singular vector is positive
    im = np.abs(ubar).argmax(axis=0)
    ubar = ubar * np.where(ubar[im, np.arange(ubar.shape[1])] < 0, -1.0, 1.0)[None, :]
    return ubar, lam


def propose(model, post, theta, z, step):
    """NonRigidIcpProposal.scala:53-68 in the reference's structure"""
    theta = np.asarray(theta, float)
    ubar, lam = posterior_basis(model, post)
    up = model.U @ ubar
    field = model.mean_def + model.Q @ post["mu"] + up @ (np.sqrt(lam) * z)
    anew = _full_coefficients(model.Q, field - model.mean_def)
    out = theta.copy()
    out[10:] = theta[10:] + (anew - theta[10:]) * step
    return out


def log_transition(model, post_from, frm, to, step):
    """NonRigidIcpProposal.scala:71-85"""
    frm, to = np.asarray(frm, float), np.asarray(to, float)
    if not np.array_equal(frm[:10], to[:10]):
        return -np.inf
    ubar, lam = posterior_basis(model, post_from)
    qp = (model.U @ ubar) * np.sqrt(lam)
    comp = frm[10:] + (to[10:] - frm[10:]) / step
    resid = (model.mean_def + model.Q @ comp) - (model.mean_def + model.Q @ post_from["mu"])
    proj = _full_coefficients(qp, resid)
    return -0.5 * (model.K * LOG_2PI + proj @ proj)


# ---- evaluators ---------------------------------------------------------------------------------
def gauss_logpdf(x, mu, sd):
    return -((x - mu) ** 2) / (2 * sd * sd) - np.log(sd * np.sqrt(2 * np.pi))


def eval_independent(model, tverts, ttris, mode, mean, sd, ids, target_points, theta):
    cur = model.transformed_mesh(theta)
    m2t = t2m = 0.0
    if mode != 1:
        _, _, d2 = closest_point_on_triangles(cur[ids], tverts, ttris)
        m2t = gauss_logpdf(np.sqrt(d2), mean, sd).sum()
    if mode != 0:
        _, _, d2 = closest_point_on_triangles(target_points, cur, model.tris)
        t2m = gauss_logpdf(np.sqrt(d2), mean, sd).sum()
    return m2t if mode == 0 else t2m if mode == 1 else 0.5 * m2t + 0.5 * t2m


def eval_hausdorff(model, tverts, ttris, rate, theta):
    cur = model.transformed_mesh(theta)
    _, _, a = closest_point_on_triangles(cur, tverts, ttris)
    _, _, b = closest_point_on_triangles(tverts, cur, model.tris)
    hd = np.sqrt(max(a.max(), b.max()))
    return np.log(rate) - rate * hd


# ---- posterior variability maps -------------------------------------------

def samples_from_log(status, take_every_n=50, total=100, burn_in=0):
    """apps/util/LogHelper.scala:27-37: indices burn_in, burn_in + n, ... below min(len, total), each walked back to
    the last accepted entry (a rejected entry carries no parameters); at most `total` of them."""
    def get_log_index(i):
        while not status[i]:
            i -= 1          # the reference recurses to i - 1 (and fails below index 0, as this does)
            if i < 0:
                raise IndexError("no accepted sample at or before the requested index")
        return i
    idx = [get_log_index(i) for i in range(burn_in, min(len(status), total), take_every_n)]
    return idx[:min(total, len(idx))]


def posterior_variability(meshes, tris, ref_verts=None, sum_normals=True):
    """apps/util/PosteriorVariability.scala:30-73, loops kept as the reference writes them (per point id, folds over
    the samples): returns (mean N x 3, cov N x 3 x 3, total N, normal N)."""
    meshes = [np.asarray(m, float) for m in meshes]
    S, N = len(meshes), len(meshes[0])
    normals = [vertex_normals(m, tris) for m in meshes] if sum_normals else None
    ref_n = None if sum_normals else vertex_normals(np.asarray(ref_verts, float), tris)
    mean, cov, total, along = np.zeros((N, 3)), np.zeros((N, 3, 3)), np.zeros(N), np.zeros(N)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv1 = np.float64(1.0) / np.float64(S - 1)
        for pid in range(N):
            samples = [m[pid] for m in meshes]
            mu = np.zeros(3)
            for s in samples:
                mu = mu + s
            mu = mu * (1.0 / S)                                   # :42 / :58
            c = np.zeros((3, 3))
            for s in samples:
                c = c + np.outer(s - mu, s - mu)                  # :43
            c = c * inv1
            if sum_normals:
                n = np.zeros(3)
                for nm in normals:
                    n = n + nm[pid] / np.linalg.norm(nm[pid])     # :60-61 (unit normals, mean not re-normalised)
                n = n * (1.0 / S)
            else:
                n = ref_n[pid] / np.linalg.norm(ref_n[pid])       # :64
            acc = 0.0
            for s in samples:
                acc = acc + float(n @ (s - mu)) ** 2              # :66
            mean[pid], cov[pid], total[pid], along[pid] = mu, c, np.trace(c), acc * inv1
    return mean, cov, total, along


# ---- GPMM construction from analytic kernels -------------------------------

def gauss_mixture_kernel(x, y, terms):
    """apps/femur/CreateGPModel.scala:70-83 [S-recall: Scalismo GaussianKernel(sigma)(x, y) = exp(-|x - y|^2 / sigma^2)]:
    k(x, y) = sum_t scale_t g_t(x, y) A_t, evaluated pair by pair. terms: (scale, sigma, A or None). -> (3 nx) x (3 ny)."""
    x, y = np.asarray(x, float).reshape(-1, 3), np.asarray(y, float).reshape(-1, 3)
    out = np.zeros((3 * len(x), 3 * len(y)))
    for i, xi in enumerate(x):
        for j, yj in enumerate(y):
            d2 = float((xi - yj) @ (xi - yj))
            b = np.zeros((3, 3))
            for scale, sigma, a in terms:
                b = b + scale * np.exp(-d2 / (sigma * sigma)) * (np.eye(3) if a is None else np.asarray(a, float))
            out[3 * i:3 * i + 3, 3 * j:3 * j + 3] = b
    return out


def bspline3(t):
    """Centred cardinal cubic B-spline, Scalismo BSpline.nthOrderBSpline(3) [S-recall]: support (-2, 2)."""
    t = abs(float(t))
    if t >= 2.0:
        return 0.0
    if t >= 1.0:
        return (2.0 - t) ** 3 / 6.0
    return 2.0 / 3.0 - t * t + 0.5 * t ** 3


def bspline_kernel3d(a, b):
    """Scalismo BSplineKernel[_3D](order = 3, scale = 0) [S-recall]: sum over the integer lattice of the products of the
    tensor-product B-splines centred at the lattice points, evaluated at a and at b; factorises over the dimensions."""
    out = 1.0
    for d in range(3):
        kl, ku = int(np.ceil(max(a[d], b[d]) - 2.0)), int(np.floor(min(a[d], b[d]) + 2.0))
        out *= sum(bspline3(a[d] - k) * bspline3(b[d] - k) for k in range(kl, ku + 1))
    return out


def face_kernel(x, y, levels, scales, symmetric_weight=0.7, plain_weight=0.3, wx=None, wy=None, wy_mirror=None):
    """apps/bfm/FaceKernel.scala:26-104, pair by pair: SpatiallyVaryingMultiscaleKernel k(x, y) = sum_l scale_l w_l(x) w_l(y)
    B3(2^l x, 2^l y) I (:40-52) and FaceKernel = 0.7 symmetrize(k) + 0.3 k (:70), symmetrize(k)(x, y) = I k(x, y) +
    diag(-1, 1, 1) k(x, ybar) (:83-95). -> (3 nx) x (3 ny)."""
    x, y = np.asarray(x, float).reshape(-1, 3), np.asarray(y, float).reshape(-1, 3)
    out = np.zeros((3 * len(x), 3 * len(y)))
    ibar = np.diag([-1.0, 1.0, 1.0])

    def k(xi, yj, i, j, wyy):
        s = 0.0
        for li, (level, scale) in enumerate(zip(levels, scales)):
            c = 2.0 ** level
            s += bspline_kernel3d(xi * c, yj * c) * scale * (1.0 if wx is None else wx[li][i]) * (1.0 if wyy is None else wyy[li][j])
        return s

    for i, xi in enumerate(x):
        for j, yj in enumerate(y):
            kk = k(xi, yj, i, j, wy)
            blk = plain_weight * kk * np.eye(3)
            if symmetric_weight != 0.0:
                kb = k(xi, yj * np.array([-1.0, 1.0, 1.0]), i, j, wy_mirror)
                blk = blk + symmetric_weight * (np.eye(3) * kk + ibar * kb)
            out[3 * i:3 * i + 3, 3 * j:3 * j + 3] = blk
    return out


def nystrom_extend(kernel_nm, v, w):
    """[S-recall] LowRankGaussianProcess.approximateGPNystrom: with (w_i, v_i) the eigenpairs of the m-point kernel matrix,
    lambda_i = w_i / m and phi_i(x) = sqrt(m) / w_i * k(x, X_m) v_i. -> (basis, variance)."""
    m = kernel_nm.shape[1] // 3
    return kernel_nm @ v * (np.sqrt(m) / w), w / m
