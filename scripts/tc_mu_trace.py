"""Pipeline timeline of CTA 0 of the MU tensor-core kernel (clock64 stamps per tile, relative to the first X TMA issue).
   python scripts/tc_mu_trace.py [k] [n d]"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pycmf_b200.device import CudaBackend, DenseMatrix
from pycmf_b200 import _lib

k = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n, d = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (20000, 50000)
be = CudaBackend(dtype="float32", options={"dense_path": 1, "tc_trace": 1})
torch = be.torch
Xd = DenseMatrix(torch.rand(n, d, device=be.device))
U, V = torch.rand(n, k, device=be.device), torch.rand(d, k, device=be.device)
names = ["mma_top", "mma_q_seen", "mma_r_seen", "mma_issued", "x_issue", "q_issue", "conv_x_seen", "conv_rfree", "conv_done"]
NT = 96


def show(tag):
    buf = np.zeros(10 * NT + 320, dtype=np.int64)
    _lib.check(be.lib.pycmf_debug_tc_trace(be.ctx, buf.ctypes.data_as(ctypes.c_void_p), buf.size))
    t = buf[:9 * NT].reshape(9, NT)
    t0 = t[4, 0]
    print("==", tag)
    print("tile " + " ".join("%11s" % s for s in names))
    for it in list(range(0, 8)) + list(range(8, NT, 4)):
        print("%4d " % it + " ".join("%11d" % (t[e, it] - t0 if t[e, it] else -1) for e in range(9)))
    steady = slice(40, 90)
    per = np.diff(t[3, steady]).mean()
    print("steady state: %.0f clk per tile; MMA warp waits: Q %.0f, R %.0f, issue %.0f clk per tile" % (
        per, (t[1, steady] - t[0, steady]).mean(), (t[2, steady] - t[1, steady]).mean(), (t[3, steady] - t[2, steady]).mean()))
    print("converter warp 0: X wait -> RFREE seen %.0f clk, RFREE -> done %.0f clk; X issue -> converter sees X %.0f clk; "
          "Q issue -> MMA sees Q %.0f clk" % ((t[7, steady] - t[6, steady]).mean(), (t[8, steady] - t[7, steady]).mean(),
                                               (t[6, steady] - t[4, steady]).mean(), (t[1, steady] - t[5, steady]).mean()))


out = be.empty(d + k, k)
for _ in range(2):
    be.mu_v_partial(Xd, U, out=out)
show("X^T U (RIGHT), k = %d" % k)
F = torch.ones(n, k, device=be.device)
be.mu_left(F, V, Xd, 0.0, 0.0)
show("X V (LEFT), k = %d" % k)
