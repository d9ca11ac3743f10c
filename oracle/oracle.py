"""ctypes binding of the C oracle (oracle/icp_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by bench.py's
cpu_baseline / --impl reference legs; never by the product package. PARITY UNPINNED (see
icp_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODEL_SAMPLING, TARGET_SAMPLING = 0, 1
MODEL_TO_TARGET, TARGET_TO_MODEL, SYMMETRIC = 0, 1, 2
PROP_ICP, PROP_RANDOM_SHAPE, PROP_ROTATION, PROP_TRANSLATION = 0, 1, 2, 3
EVAL_ACCEPT_ALL, EVAL_INDEPENDENT, EVAL_HAUSDORFF, EVAL_COLLECTIVE = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
_vp = C.c_void_p


def build(force=False):
    so = os.path.join(_HERE, "libicporacle.so")
    src = os.path.join(_HERE, "icp_oracle.c")
    hdr = os.path.join(_HERE, "icp_oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libicporacle.so"])
    return so


class _Component(C.Structure):
    _fields_ = [("kind", C.c_int), ("weight", C.c_double), ("icp", _vp), ("sd", C.c_double), ("axis", C.c_int)]


class _ChainDesc(C.Structure):
    _fields_ = [("model", _vp), ("target", _vp), ("n_components", C.c_int), ("components", C.POINTER(_Component)),
                ("use_prior", C.c_int), ("eval_kind", C.c_int), ("eval_mode", C.c_int),
                ("p0", C.c_double), ("p1", C.c_double), ("p2", C.c_double),
                ("n_ids", C.c_int), ("ids", _ip), ("n_tp", C.c_int), ("target_points", _dp),
                ("closed_form", C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_mesh_create.restype = _vp
        L.orc_mesh_create.argtypes = [C.c_int, _dp, C.c_int, _ip]
        L.orc_mesh_free.argtypes = [_vp]
        L.orc_closest_point_brute.argtypes = [C.c_int, _dp, C.c_int, _ip, C.c_int, _dp, _ip, _ip, _dp, _dp]
        L.orc_mesh_closest_point.argtypes = [_vp, C.c_int, _dp, _ip, _ip, _dp, _dp]
        L.orc_point_triangle_d2.restype = C.c_double
        L.orc_point_triangle_d2.argtypes = [_dp, _dp, _dp, _dp, _dp, _ip]
        L.orc_closest_vertex_brute.argtypes = [C.c_int, _dp, C.c_int, _dp, _ip, _dp]
        L.orc_mesh_closest_vertex.argtypes = [_vp, C.c_int, _dp, _ip, _dp]
        L.orc_mesh_boundary_flags.argtypes = [_vp, _bp]
        L.orc_mesh_vertex_normals.argtypes = [_vp, _dp]
        L.orc_model_create.restype = _vp
        L.orc_model_create.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _ip]
        L.orc_model_free.argtypes = [_vp]
        L.orc_transformed_mesh.argtypes = [_vp, _dp, _dp]
        L.orc_pose_matrix.argtypes = [_dp, _dp]
        L.orc_surface_noise_cov.argtypes = [_dp, C.c_double, C.c_double, _dp]
        L.orc_proposal_create.restype = _vp
        L.orc_proposal_create.argtypes = [_vp, _vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _ip, C.c_int, _dp]
        L.orc_proposal_free.argtypes = [_vp]
        L.orc_icp_posterior.restype = C.c_int
        L.orc_icp_posterior.argtypes = [_vp, _dp, _dp, _dp, _dp, _ip, _dp, _dp]
        L.orc_propose.argtypes = [_vp, _dp, _dp, _dp]
        L.orc_propose_closed_form.argtypes = [_vp, _dp, _dp, _dp]
        L.orc_log_transition.restype = C.c_double
        L.orc_log_transition.argtypes = [_vp, _dp, _dp]
        L.orc_log_transition_closed_form.restype = C.c_double
        L.orc_log_transition_closed_form.argtypes = [_vp, _dp, _dp]
        L.orc_std_icp_iteration.argtypes = [_vp, _vp, C.c_int, C.c_int, _ip, C.c_int, _dp, C.c_double, C.c_double, _dp, _dp]
        L.orc_std_icp_iteration_theta.argtypes = [_vp, _vp, C.c_int, C.c_int, _ip, C.c_int, _dp, C.c_double, C.c_double, _dp, _dp]
        L.orc_eval_independent.restype = C.c_double
        L.orc_eval_independent.argtypes = [_vp, _vp, C.c_int, C.c_double, C.c_double, C.c_int, _ip, C.c_int, _dp, _dp]
        L.orc_eval_hausdorff.restype = C.c_double
        L.orc_eval_hausdorff.argtypes = [_vp, _vp, C.c_double, _dp]
        L.orc_eval_collective.restype = C.c_double
        L.orc_eval_collective.argtypes = [_vp, _vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, _ip, C.c_int, _dp, _dp,
                                          C.POINTER(C.c_int), _dp]
        L.orc_eval_prior.restype = C.c_double
        L.orc_eval_prior.argtypes = [C.c_int, _dp]
        L.orc_random_walk_log_transition.restype = C.c_double
        L.orc_random_walk_log_transition.argtypes = [C.c_int, C.c_double, _dp, _dp]
        L.orc_pose_log_transition.restype = C.c_double
        L.orc_pose_log_transition.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp]
        L.orc_registration_metrics.argtypes = [_vp, _vp, _dp, _dp]
        L.orc_dice_coefficient.restype = C.c_double
        L.orc_dice_coefficient.argtypes = [_vp, _vp, _dp, C.c_int, _dp]
        L.orc_chain_run.restype = C.c_int
        L.orc_chain_run.argtypes = [C.POINTER(_ChainDesc), _dp, C.c_int, _dp, _dp, _dp, _ip, _bp, _dp, _dp]
        L.orc_use_blas.restype = C.c_int
        L.orc_use_blas.argtypes = [C.c_char_p]
        L.orc_blas_enabled.restype = C.c_int
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        _LIB = L
    return _LIB


def find_blas():
    """Path of an OpenBLAS build with CBLAS + LAPACKE entry points (the one scipy bundles), or None."""
    import glob
    try:
        import scipy
        hits = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so")))
    except Exception:
        hits = []
    for cand in hits + ["libopenblas.so.0", "libopenblas.so"]:
        return cand
    return None


def use_blas(enable=True):
    """Routes the oracle's dense linear algebra through OpenBLAS / LAPACK (timed CPU arm of bench.py). Returns the library
    path in use, or None when the built-in loops stay active."""
    if not enable:
        lib().orc_use_blas(None)
        return None
    path = find_blas()
    if path and lib().orc_use_blas(path.encode()) == 0:
        return path
    return None


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


class Mesh:
    """Scalismo TriangleMesh3D stand-in with its query helpers."""

    def __init__(self, verts, tris):
        self.verts, vp = _d(np.asarray(verts).reshape(-1, 3))
        self.tris, tp = _i(np.asarray(tris).reshape(-1, 3))
        self.h = lib().orc_mesh_create(len(self.verts), vp, len(self.tris), tp)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_mesh_free(self.h)
            self.h = None

    def closest_point(self, q, brute=False):
        q, qp = _d(np.asarray(q).reshape(-1, 3))
        n = len(q)
        tri = np.empty(n, np.int32); feat = np.empty(n, np.int32); cp = np.empty((n, 3)); d2 = np.empty(n)
        if brute:
            lib().orc_closest_point_brute(len(self.verts), self.verts.ctypes.data_as(_dp), len(self.tris),
                                          self.tris.ctypes.data_as(_ip), n, qp, tri.ctypes.data_as(_ip),
                                          feat.ctypes.data_as(_ip), cp.ctypes.data_as(_dp), d2.ctypes.data_as(_dp))
        else:
            lib().orc_mesh_closest_point(self.h, n, qp, tri.ctypes.data_as(_ip), feat.ctypes.data_as(_ip),
                                         cp.ctypes.data_as(_dp), d2.ctypes.data_as(_dp))
        return tri, feat, cp, d2

    def closest_vertex(self, q, brute=False):
        q, qp = _d(np.asarray(q).reshape(-1, 3))
        n = len(q)
        ids = np.empty(n, np.int32); d2 = np.empty(n)
        if brute:
            lib().orc_closest_vertex_brute(len(self.verts), self.verts.ctypes.data_as(_dp), n, qp,
                                           ids.ctypes.data_as(_ip), d2.ctypes.data_as(_dp))
        else:
            lib().orc_mesh_closest_vertex(self.h, n, qp, ids.ctypes.data_as(_ip), d2.ctypes.data_as(_dp))
        return ids, d2

    def boundary_flags(self):
        f = np.zeros(max(len(self.verts), 1), np.uint8)
        lib().orc_mesh_boundary_flags(self.h, f.ctypes.data_as(_bp))
        return f[:len(self.verts)].astype(bool)

    def vertex_normals(self):
        n = np.empty((len(self.verts), 3))
        lib().orc_mesh_vertex_normals(self.h, n.ctypes.data_as(_dp))
        return n


def point_triangle_d2(q, a, b, c):
    q, qp = _d(q); a, ap = _d(a); b, bp = _d(b); c, cp_ = _d(c)
    out = np.empty(3); f = C.c_int32(0)
    d2 = lib().orc_point_triangle_d2(qp, ap, bp, cp_, out.ctypes.data_as(_dp), C.byref(f))
    return d2, out, f.value


class Model:
    """Scalismo StatisticalMeshModel stand-in (reference mesh + low-rank GP)."""

    def __init__(self, ref, tris, basis, variance, mean_def=None):
        self.ref, rp = _d(np.asarray(ref).reshape(-1, 3))
        self.tris, tp = _i(np.asarray(tris).reshape(-1, 3))
        self.basis, bp = _d(basis)
        self.variance, vp = _d(variance)
        self.N, self.T, self.K = len(self.ref), len(self.tris), len(self.variance)
        assert self.basis.shape == (3 * self.N, self.K)
        self.mean_def, mp = _d(np.zeros(3 * self.N) if mean_def is None else np.asarray(mean_def).reshape(-1))
        self.h = lib().orc_model_create(self.N, self.T, self.K, rp, mp, bp, vp, tp)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_model_free(self.h)
            self.h = None

    def transformed_mesh(self, theta):
        th, thp = _d(theta)
        out = np.empty((self.N, 3))
        lib().orc_transformed_mesh(self.h, thp, out.ctypes.data_as(_dp))
        return out


def pose_matrix(theta):
    th, thp = _d(theta)
    r = np.empty((3, 3))
    lib().orc_pose_matrix(thp, r.ctypes.data_as(_dp))
    return r


def surface_noise_cov(normal, sd_normal, sd_tangent):
    n, np_ = _d(normal)
    out = np.empty((3, 3))
    lib().orc_surface_noise_cov(np_, sd_normal, sd_tangent, out.ctypes.data_as(_dp))
    return out


class IcpProposal:
    def __init__(self, model: Model, target: Mesh, step_length, tangential_noise, noise_along_normal, direction,
                 boundary_aware, ids, target_points):
        self.model, self.target = model, target
        self.ids, ip = _i(np.asarray(ids).reshape(-1))
        self.tp, tpp = _d(np.asarray(target_points).reshape(-1, 3))
        self.direction = direction
        self.h = lib().orc_proposal_create(model.h, target.h, step_length, tangential_noise, noise_along_normal,
                                           direction, int(boundary_aware), len(self.ids), ip, len(self.tp), tpp)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_proposal_free(self.h)
            self.h = None

    def posterior(self, theta, with_obs=False):
        K = self.model.K
        th, thp = _d(theta)
        mu = np.empty(K); M = np.empty((K, K)); Minv = np.empty((K, K))
        nmax = max(len(self.ids), len(self.tp), 1)
        if with_obs:
            oid = np.empty(nmax, np.int32); oy = np.empty((nmax, 3)); oc = np.empty((nmax, 3, 3))
            n = lib().orc_icp_posterior(self.h, thp, mu.ctypes.data_as(_dp), M.ctypes.data_as(_dp), Minv.ctypes.data_as(_dp),
                                        oid.ctypes.data_as(_ip), oy.ctypes.data_as(_dp), oc.ctypes.data_as(_dp))
            return dict(n=n, mu=mu, M=M, Minv=Minv, ids=oid[:n], y=oy[:n], cov=oc[:n])
        n = lib().orc_icp_posterior(self.h, thp, mu.ctypes.data_as(_dp), M.ctypes.data_as(_dp), Minv.ctypes.data_as(_dp), None, None, None)
        return dict(n=n, mu=mu, M=M, Minv=Minv)

    def propose(self, theta, z, closed_form=False):
        th, thp = _d(theta); zz, zp = _d(z)
        out = np.empty_like(th)
        (lib().orc_propose_closed_form if closed_form else lib().orc_propose)(self.h, thp, zp, out.ctypes.data_as(_dp))
        return out

    def log_transition(self, frm, to, closed_form=False):
        f, fp = _d(frm); t, tp = _d(to)
        fn = lib().orc_log_transition_closed_form if closed_form else lib().orc_log_transition
        return fn(self.h, fp, tp)


def std_icp_iteration(model, target, direction, ids, target_points, sigma2, step_length, alpha):
    ids, ip = _i(np.asarray(ids).reshape(-1)); tp, tpp = _d(np.asarray(target_points).reshape(-1, 3))
    a, ap = _d(alpha)
    out = np.empty_like(a)
    lib().orc_std_icp_iteration(model.h, target.h, direction, len(ids), ip, len(tp), tpp, sigma2, step_length, ap,
                                out.ctypes.data_as(_dp))
    return out


def std_icp_iteration_theta(model, target, direction, ids, target_points, sigma2, step_length, theta):
    ids, ip = _i(np.asarray(ids).reshape(-1)); tp, tpp = _d(np.asarray(target_points).reshape(-1, 3))
    th, thp = _d(theta)
    out = np.empty(model.K)
    lib().orc_std_icp_iteration_theta(model.h, target.h, direction, len(ids), ip, len(tp), tpp, sigma2, step_length, thp,
                                      out.ctypes.data_as(_dp))
    return out


def eval_independent(model, target, mode, mean, sd, ids, target_points, theta):
    ids, ip = _i(np.asarray(ids).reshape(-1)); tp, tpp = _d(np.asarray(target_points).reshape(-1, 3)); th, thp = _d(theta)
    return lib().orc_eval_independent(model.h, target.h, mode, mean, sd, len(ids), ip, len(tp), tpp, thp)


def eval_hausdorff(model, target, rate, theta):
    th, thp = _d(theta)
    return lib().orc_eval_hausdorff(model.h, target.h, rate, thp)


def eval_collective(model, target, mode, avg_mean, avg_sd, max_rate, ids, target_points, theta):
    ids, ip = _i(np.asarray(ids).reshape(-1)); tp, tpp = _d(np.asarray(target_points).reshape(-1, 3)); th, thp = _d(theta)
    st = C.c_int(0); am = np.empty(2)
    v = lib().orc_eval_collective(model.h, target.h, mode, avg_mean, avg_sd, max_rate, len(ids), ip, len(tp), tpp, thp,
                                  C.byref(st), am.ctypes.data_as(_dp))
    return v, st.value, am


def eval_prior(K, theta):
    th, thp = _d(theta)
    return lib().orc_eval_prior(K, thp)


def random_walk_log_transition(K, sd, frm, to):
    f, fp = _d(frm); t, tp = _d(to)
    return lib().orc_random_walk_log_transition(K, sd, fp, tp)


def pose_log_transition(K, kind, axis, sd, frm, to):
    f, fp = _d(frm); t, tp = _d(to)
    return lib().orc_pose_log_transition(K, kind, axis, sd, fp, tp)


def registration_metrics(model, target, theta):
    th, thp = _d(theta)
    out = np.empty(4)
    lib().orc_registration_metrics(model.h, target.h, thp, out.ctypes.data_as(_dp))
    return out


def dice_coefficient(model, target, theta, unit_samples):
    th, thp = _d(theta); u, up = _d(np.asarray(unit_samples).reshape(-1, 3))
    return lib().orc_dice_coefficient(model.h, target.h, thp, len(u), up)


def chain_run(model, target, components, use_prior, eval_kind, eval_mode, params, ids, target_points, theta0, n_steps,
              u_comp, z, u_acc, closed_form=False):
    """components: list of dicts(kind, weight, icp=IcpProposal|None, sd, axis)."""
    K = model.K
    comps = (_Component * len(components))()
    for i, c in enumerate(components):
        comps[i].kind = c["kind"]; comps[i].weight = c["weight"]
        comps[i].icp = c["icp"].h if c.get("icp") is not None else None
        comps[i].sd = c.get("sd", 0.0); comps[i].axis = c.get("axis", 0)
    ids, ip = _i(np.asarray(ids).reshape(-1)); tp, tpp = _d(np.asarray(target_points).reshape(-1, 3))
    p = list(params) + [0.0] * 3
    d = _ChainDesc(model.h, target.h, len(components), comps, int(use_prior), eval_kind, eval_mode, p[0], p[1], p[2],
                   len(ids), ip, len(tp), tpp, int(closed_form))
    th0, th0p = _d(theta0); uc, ucp = _d(u_comp); zz, zp = _d(z); ua, uap = _d(u_acc)
    comp = np.empty(n_steps, np.int32); acc = np.empty(n_steps, np.uint8)
    logv = np.empty((n_steps, 3)); thl = np.empty((n_steps, K + 10))
    n_acc = lib().orc_chain_run(C.byref(d), th0p, n_steps, ucp, zp, uap, comp.ctypes.data_as(_ip), acc.ctypes.data_as(_bp),
                                logv.ctypes.data_as(_dp), thl.ctypes.data_as(_dp))
    return dict(n_accepted=n_acc, comp=comp, accepted=acc.astype(bool), logv=logv, theta=thl)


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*[int(x) for x in ctr]); k = (C.c_uint32 * 2)(*[int(x) for x in key]); o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return np.array(list(o), dtype=np.uint32)
