"""Where the end-to-end time of fit_iterative_update goes on C2 (host float64 pinned arrays -> factors back on the host).
   python scripts/e2e_breakdown.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pycmf_b200 import workloads as W
from pycmf_b200.cmf_solvers import NewtonSolver
from pycmf_b200.device import CudaBackend

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
be = CudaBackend(dtype="float32")
cfg = W.describe("c2", 1.0)
data = W.generate(be, "c2", 0, cfg["n"], 1.0)
U, V, Z = W.finish_init(be, data, data["x_sum"])


def pinned(t):
    h = torch.empty(t.shape, dtype=torch.float64, pin_memory=True)
    h.copy_(t.double())
    return h.numpy()


Xh, Yh, Uh, Vh, Zh = pinned(data["X"].t), pinned(data["Y"].t), pinned(U), pinned(V), pinned(Z)
del data
p = dict(W.SOLVER_PARAMS["c2"])
s = NewtonSolver(tol=0, x_link="linear", y_link="logit", dtype="float32", backend=be, max_iter=steps, **p)


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(3):
    Uc, Vc, Zc = Uh.copy(), Vh.copy(), Zh.copy()
    t0 = sync()
    st = s.prepare(Xh, Yh, Uc, Vc, Zc)
    t1 = sync()
    step = s.make_stepper(st)
    step(); step()
    t2 = sync()
    step()                      # capture + first replay
    t3 = sync()
    cap_ms = s.capture_seconds_ * 1e3
    for _ in range(steps - 3):
        step()
    t4 = sync()
    for host, dev in ((Uc, st.U), (Vc, st.V), (Zc, st.Z)):
        host[...] = be.to_host(dev)
    t5 = sync()
    print("rep %d: prepare (H2D %.0f MB + casts) %.2f ms (%.1f GB/s) | 2 eager iterations %.2f ms | capture + replay %.2f ms (capture call %.2f) | "
          "%d replays %.2f ms (%.3f ms each) | read-back %.2f ms | total %.2f ms" % (
              rep, Xh.nbytes / 1e6, (t1 - t0) * 1e3, Xh.nbytes / (t1 - t0) / 1e9, (t2 - t1) * 1e3, (t3 - t2) * 1e3, cap_ms, steps - 3,
              (t4 - t3) * 1e3, (t4 - t3) * 1e3 / max(steps - 3, 1), (t5 - t4) * 1e3, (t5 - t0) * 1e3), flush=True)
    del st
# the same through the public call
for rep in range(2):
    Uc, Vc, Zc = Uh.copy(), Vh.copy(), Zh.copy()
    t0 = sync()
    s.fit_iterative_update(Xh, Yh, Uc, Vc, Zc)
    t1 = sync()
    print("fit_iterative_update: %.2f ms (graph capture %.2f ms)" % ((t1 - t0) * 1e3, s.capture_seconds_ * 1e3), flush=True)
