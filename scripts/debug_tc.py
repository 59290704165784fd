"""Debug helper for the tcgen05 kernel: structured inputs, prints error statistics per mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pycmf_b200.device import CudaBackend

np.set_printoptions(precision=4, linewidth=200, suppress=True)
rng = np.random.RandomState(0)
for path in (2, 1):
    for splits in (1, 0):
        be = CudaBackend(dtype="float32", options={"dense_path": path, "tc_max_splits": splits} if splits else {"dense_path": path})
        for (n, d) in ((256, 256), (1000, 520)):
            X = rng.randn(n, d).astype(np.float32).astype(np.float64)
            U, V = rng.randn(n, 32), rng.randn(d, 32)
            Xd = be.ingest(X)
            buf = be.to_host(be.mu_v_partial(Xd, be.to_device(U)))[:d]
            ref = X.T @ U
            print("path %d splits %d n %d d %d XtU: max|got| %.4f rel err %.3e" % (path, splits, n, d, np.abs(buf).max(), np.linalg.norm(buf - ref) / np.linalg.norm(ref)))
            if np.linalg.norm(buf - ref) / np.linalg.norm(ref) > 1e-2:
                print(" got[:4,:6]\n", buf[:4, :6], "\n ref[:4,:6]\n", ref[:4, :6])
            F = np.ones((n, 32)); Fd = be.to_device(F)
            be.mu_left(Fd, be.to_device(V), Xd, 0.0, 0.0)
            got = be.to_host(Fd) * (F @ (V.T @ V)); ref = X @ V
            print("                              XV : max|got| %.4f rel err %.3e" % (np.abs(got).max(), np.linalg.norm(got - ref) / np.linalg.norm(ref)))
            if np.linalg.norm(got - ref) / np.linalg.norm(ref) > 1e-2:
                print(" got[:4,:6]\n", got[:4, :6], "\n ref[:4,:6]\n", ref[:4, :6])
            Us, Vs = 0.3 * U, 0.3 * V
            for link in ("linear", "logit"):
                gx, Hx, pr = be.newton_v_xpart(be.to_device(Vs), be.to_device(Us), Xd, 0, d, link, 1.0)
                S = Us @ Vs.T
                est = 1 / (1 + np.exp(-S)) if link == "logit" else S
                ref = (est - X).T @ Us
                g = be.to_host(gx)
                print("                              resid_right %s: rel err %.3e" % (link, np.linalg.norm(g - ref) / np.linalg.norm(ref)))
                if np.linalg.norm(g - ref) / np.linalg.norm(ref) > 1e-2:
                    print(" got[:4,:6]\n", g[:4, :6], "\n ref[:4,:6]\n", ref[:4, :6])
        be.close()
