import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def femur():
    """The reference's femur fixtures (committed as tests/golden/*.npz by tests/golden/make_golden.py)."""
    meshes = np.load(os.path.join(GOLDEN, "femur_meshes.npz"))
    out = {"ref": meshes["ref"].astype(np.float64), "cells": meshes["cells"].astype(np.int32),
           "target": meshes["target_aligned"].astype(np.float64), "target_cells": meshes["target_cells"].astype(np.int32)}
    for k in ("50", "100", "200"):
        g = np.load(os.path.join(GOLDEN, f"femur_gpmm_{k}.npz"))
        out[f"gpmm_{k}"] = {"basis": g["basis"].astype(np.float64), "variance": g["variance"].astype(np.float64)}
    with open(os.path.join(GOLDEN, "femur_golden.json")) as f:
        out["golden"] = json.load(f)
    return out


@pytest.fixture(scope="session")
def twin31():
    from icp_proposal_b200 import synth
    m = synth.femur_twin(rank=31)
    tv, tc, alpha = synth.synthetic_target(m)
    m["target"], m["target_cells"], m["target_alpha"] = tv, tc, alpha
    return m


@pytest.fixture(scope="session")
def twin101():
    from icp_proposal_b200 import synth
    m = synth.femur_twin(rank=101)
    tv, tc, alpha = synth.synthetic_target(m)
    m["target"], m["target_cells"], m["target_alpha"] = tv, tc, alpha
    return m


@pytest.fixture(scope="session")
def open_twin():
    """Small open surface (boundary present) + partial target: exercises the boundary-aware paths."""
    from icp_proposal_b200 import synth
    verts, tris = synth.height_field_mesh(side=33, extent=160.0)
    basis, var = synth.nystrom_gpmm(verts, 21, [(64.0, 12.0, False), (32.0, 6.0, False), (16.0, 2.0, False)], n_nystrom=120)
    m = dict(ref=verts, cells=tris, basis=basis, variance=var)
    tv, tc, alpha = synth.partial_target(m, seed=3, alpha_sd=0.4)
    m["target"], m["target_cells"], m["target_alpha"] = tv, tc, alpha
    return m


def random_theta(model, rng, n, alpha_sd=0.3, pose=False):
    K = len(model["variance"])
    th = np.zeros((n, K + 10))
    th[:, 0] = 1.0
    th[:, 7:10] = model["ref"].mean(0)
    th[:, 10:] = rng.normal(0, alpha_sd, (n, K))
    if pose:
        th[:, 1:4] = rng.normal(0, 0.5, (n, 3))
        th[:, 4:7] = rng.normal(0, 0.02, (n, 3))
    return th


@pytest.fixture(scope="session")
def ctx():
    """CUDA context of the product library. GPU tests fail (not skip) when the extension is missing."""
    from icp_proposal_b200 import core
    c = core.Context(0)
    yield c
    c.close()
