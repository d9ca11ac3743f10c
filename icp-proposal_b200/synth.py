"""Synthetic meshes and GPMMs of the shapes BASELINE.json names (host-side data preparation).

Nothing here is on the hot path; it builds the arrays that cross the C ABI in bench.py and the
tests. Shapes follow SURVEY.md section 8d:

* femur twin: closed genus-0 mesh, N = 1622 vertices, T = 3240 triangles (2N - 4), semi-axes
  45 x 35 x 223 mm, GPMM from the reference's own kernel recipe
  (src/main/scala/apps/femur/CreateGPModel.scala:68-93: B*10*g90 + 5*g40 + 3*g10, Nystrom on 2K
  surface points, K + 1 basis functions).
* face-sized twin: 169 x 169 open height-field grid (N = 28 561, T = 56 448, with boundary),
  analytic multi-scale GPMM (cf. apps/bfm/FaceKernel.scala:63-74).
"""
from __future__ import annotations

import numpy as np


def fibonacci_ellipsoid_mesh(n=1622, semi_axes=(45.0, 35.0, 223.0)):
    """Closed, outward-oriented triangulation of n Fibonacci points on an ellipsoid."""
    from scipy.spatial import ConvexHull

    i = np.arange(n) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    golden = np.pi * (1.0 + 5.0 ** 0.5)
    sphere = np.stack([np.cos(golden * i) * np.sin(phi), np.sin(golden * i) * np.sin(phi), np.cos(phi)], axis=1)
    hull = ConvexHull(sphere)
    tris = hull.simplices.astype(np.int32)
    # orient outward (on the unit sphere the centroid direction is the outward normal)
    a, b, c = sphere[tris[:, 0]], sphere[tris[:, 1]], sphere[tris[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(b - a, c - a), a + b + c) < 0
    tris[flip] = tris[flip][:, [0, 2, 1]]
    order = np.lexsort((tris[:, 2], tris[:, 1], tris[:, 0]))  # deterministic triangle order
    tris = np.ascontiguousarray(tris[order])
    verts = sphere * np.asarray(semi_axes, dtype=np.float64)
    # store as float32-representable values like the reference's STL/HDF5 fixtures
    verts = verts.astype(np.float32).astype(np.float64)
    assert len(tris) == 2 * n - 4
    return verts, tris


def height_field_mesh(side=169, extent=160.0, bump=35.0):
    """Open face-sized grid surface with boundary: side^2 vertices, 2 (side-1)^2 triangles."""
    g = np.linspace(-extent / 2, extent / 2, side)
    x, y = np.meshgrid(g, g, indexing="ij")
    r2 = (x / (0.45 * extent)) ** 2 + (y / (0.55 * extent)) ** 2
    z = bump * np.exp(-r2) + 12.0 * np.exp(-((x / 14.0) ** 2 + ((y + 5.0) / 22.0) ** 2))  # face + nose bump
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32).astype(np.float64)
    idx = np.arange(side * side).reshape(side, side)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    tris = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.int32)
    return verts, np.ascontiguousarray(tris)


def _main_axes(verts):
    c = verts - verts.mean(0)
    u, _, _ = np.linalg.svd(c.T @ c / len(verts))
    return u


def nystrom_gpmm(verts, rank, kernels, seed=1024, n_nystrom=None, base_matrix=None):
    """Low-rank GP over ``verts`` by the Nystrom method.

    kernels: list of (sigma, scale, use_base_matrix); k(x, y) = sum_l scale_l g_{sigma_l}(x, y) A_l with
    g_s(x, y) = exp(-|x - y|^2 / s^2) and A_l = base_matrix or I_3.
    Returns (basis U (3N, rank) unscaled, variance (rank,)) with Q = U sqrt(variance) such that
    Q Q^T is the Nystrom approximation of the kernel matrix on the vertices.
    """
    rng = np.random.default_rng(seed)
    n = len(verts)
    m = min(n, n_nystrom if n_nystrom is not None else 2 * rank)
    sel = np.sort(rng.choice(n, size=m, replace=False))
    xm = verts[sel]
    eye = np.eye(3)
    bm = eye if base_matrix is None else base_matrix

    def kmat(x, y):
        d2 = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
        out = np.zeros((len(x), 3, len(y), 3))
        for sigma, scale, use_b in kernels:
            g = scale * np.exp(-d2 / (sigma * sigma))
            out += g[:, None, :, None] * (bm if use_b else eye)[None, :, None, :]
        return out.reshape(3 * len(x), 3 * len(y))

    kmm = kmat(xm, xm)
    w, v = np.linalg.eigh(kmm)
    order = np.argsort(w)[::-1][:rank]
    w, v = w[order], v[:, order]
    # deterministic sign: largest-magnitude entry positive
    sgn = np.sign(v[np.abs(v).argmax(0), np.arange(rank)])
    v = v * sgn
    q = np.empty((3 * n, rank))
    for s in range(0, n, 2048):
        q[3 * s:3 * min(n, s + 2048)] = kmat(verts[s:s + 2048], xm) @ (v / np.sqrt(w))
    variance = w / m
    basis = q / np.sqrt(variance)
    # float32-representable, like the statismo files
    return basis.astype(np.float32).astype(np.float64), variance.astype(np.float32).astype(np.float64)


def femur_twin(rank=101, n=1622, seed=1024):
    """Synthetic twin of the femur GPMM (config 1/3/4 shape). -> dict(ref, cells, basis, variance)."""
    verts, tris = fibonacci_ellipsoid_mesh(n)
    d = _main_axes(verts)
    base = d @ np.diag([10.0, 1.0, 1.0]) @ d.T
    basis, var = nystrom_gpmm(verts, rank, [(90.0, 10.0, True), (40.0, 5.0, False), (10.0, 3.0, False)], seed=seed,
                              base_matrix=base)
    return dict(ref=verts, cells=tris, basis=basis, variance=var)


def face_twin(rank=100, side=169, seed=1024):
    """Face-sized analytic GPMM (config 5 shape)."""
    verts, tris = height_field_mesh(side)
    basis, var = nystrom_gpmm(verts, rank, [(64.0, 12.0, False), (32.0, 6.0, False), (16.0, 2.0, False), (8.0, 0.6, False)],
                              seed=seed, n_nystrom=8 * rank)
    return dict(ref=verts, cells=tris, basis=basis, variance=var)


def model_instance(model, alpha):
    q = model["basis"] * np.sqrt(model["variance"])
    return model["ref"] + (q @ np.asarray(alpha, dtype=np.float64)).reshape(-1, 3)


def synthetic_target(model, seed=7, alpha_sd=0.5, vertex_noise=0.0):
    """Target = model instance with alpha ~ N(0, alpha_sd^2 I) (N(0, 0.25 I), seed 7)."""
    rng = np.random.default_rng(seed)
    alpha = rng.normal(0.0, alpha_sd, size=len(model["variance"]))
    verts = model_instance(model, alpha)
    if vertex_noise > 0:
        verts = verts + rng.normal(0.0, vertex_noise, size=verts.shape)
    return verts, model["cells"].copy(), alpha


def partial_target(model, seed=7, alpha_sd=0.5, hole_centers=((0.0, -5.0), (0.0, -40.0)), hole_radius=(18.0, 14.0)):
    """Partial target with removed discs (cf. apps/bfm/AlignShapes.scala:88-92): returns a compacted mesh."""
    verts, tris, alpha = synthetic_target(model, seed, alpha_sd)
    keep = np.ones(len(verts), bool)
    for (cx, cy), r in zip(hole_centers, hole_radius):
        keep &= ((verts[:, 0] - cx) ** 2 + (verts[:, 1] - cy) ** 2) > r * r
    tkeep = keep[tris].all(1)
    new_id = -np.ones(len(verts), np.int64)
    new_id[keep] = np.arange(keep.sum())
    return verts[keep], new_id[tris[tkeep]].astype(np.int32), alpha


def near_surface_queries(verts, tris, n, seed=11, sd=2.0):
    """Target-surface points displaced along the face normal by N(0, sd) (seed 11)."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, len(tris), size=n)
    a, b, c = verts[tris[t, 0]], verts[tris[t, 1]], verts[tris[t, 2]]
    r1, r2 = np.sqrt(rng.random(n)), rng.random(n)
    p = (1 - r1)[:, None] * a + (r1 * (1 - r2))[:, None] * b + (r1 * r2)[:, None] * c
    nrm = np.cross(b - a, c - a)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return p + nrm * rng.normal(0.0, sd, size=(n, 1))


def far_field_queries(verts, n, seed=12, scale=1.5):
    """Uniform in the scale x bounding box (seed 12)."""
    rng = np.random.default_rng(seed)
    lo, hi = verts.min(0), verts.max(0)
    c, h = 0.5 * (lo + hi), 0.5 * (hi - lo) * scale
    return c + (rng.random((n, 3)) * 2 - 1) * h
