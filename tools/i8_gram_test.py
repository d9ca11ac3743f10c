"""tcgen05 INT8 self-test and MMA-rate measurement for the rank update's tile shape (csrc/tc_i8.cu)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp_proposal_b200 import _lib, core  # noqa: E402

ctx = core.Context(0)
lib = ctx.lib
rng = np.random.default_rng(0)
out = {}
for rows in (32, 96, 608):
    A = rng.integers(-127, 128, (rows, 128)).astype(np.int8)
    D = np.zeros((128, 128), np.int32)
    ms = C.c_double(0)
    _lib.check(lib.icp_debug_i8_gram(ctx.h, rows, A.ctypes.data, D.ctypes.data, 1, 1, C.byref(ms)), ctx.h)
    want = A.astype(np.int64).T @ A.astype(np.int64)
    ok = bool(np.array_equal(D[:, :112], want[:, :112]))
    out[f"rows_{rows}"] = {"exact": ok, "max_abs_diff": int(np.abs(D[:, :112] - want[:, :112]).max())}
    if not ok:
        bad = np.argwhere(D[:, :112] != want[:, :112])
        out[f"rows_{rows}"]["first_bad"] = bad[:5].tolist()
        out[f"rows_{rows}"]["got"] = D[:2, :4].tolist(); out[f"rows_{rows}"]["want"] = want[:2, :4].tolist()
A = rng.integers(-127, 128, (608, 128)).astype(np.int8)
for ctas, iters in ((1, 400), (148, 400), (296, 400)):
    ms = C.c_double(0)
    _lib.check(lib.icp_debug_i8_gram(ctx.h, 608, A.ctypes.data, None, iters, ctas, C.byref(ms)), ctx.h)
    ops = 2.0 * 128 * 112 * 608 * iters * ctas
    out[f"rate_ctas_{ctas}"] = {"ms": ms.value, "TOPS": ops / (ms.value * 1e-3) / 1e12, "us_per_608row_gram_per_cta": ms.value * 1e3 / iters}
print(json.dumps(out, indent=1))
