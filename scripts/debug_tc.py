"""Accuracy of the fused residual pass (generic FMA vs tcgen05 3xTF32 / 1xTF32) against float64 NumPy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pycmf_b200.device import CudaBackend

rng = np.random.RandomState(0)
def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)
for (n, d, kind) in ((700, 1300, "nonneg-fit"), (3000, 2000, "nonneg-fit"), (1000, 520, "randn")):
    if kind == "randn":
        U, V = 0.3 * rng.randn(n, 32), 0.3 * rng.randn(d, 32)
        X = U @ V.T + 0.05 * rng.randn(n, d)
    else:
        Ut, Vt = 0.6 * np.abs(rng.randn(n, 32)), 0.6 * np.abs(rng.randn(d, 32))
        X = Ut @ Vt.T + 0.05 * np.abs(rng.randn(n, d))
        U, V = Ut * (1 + 0.05 * rng.randn(n, 32)), Vt * (1 + 0.05 * rng.randn(d, 32))
    X = X.astype(np.float32).astype(np.float64)
    U = U.astype(np.float32).astype(np.float64); V = V.astype(np.float32).astype(np.float64)
    R = U @ V.T - X
    refL, refR, refsq = R @ V, R.T @ U, (R * R).sum()
    for path, splits in ((0, 0), (1, 0), (1, 1), (2, 0)):
        opts = {"dense_path": path}
        if splits: opts["tc_max_splits"] = splits
        be = CudaBackend(dtype="float32", options=opts)
        outL, outR, sq = be.resid_pass(be.to_device(U), be.to_device(V), be.ingest(X), "linear", want_sq=True)
        print("%-10s n=%d d=%d path=%d splits=%d  relL %.2e  relR %.2e  relsq %.2e" % (
            kind, n, d, path, splits, rel(be.to_host(outL), refL), rel(be.to_host(outR), refR),
            abs(float(be.to_host(sq)[0]) - refsq) / refsq))
        be.close()
