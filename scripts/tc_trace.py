"""Prints the pipeline timeline (cycles relative to the first TMA issue) of CTA (0,0) of a tcgen05 pass."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pycmf_b200.device import CudaBackend
from pycmf_b200 import _lib

n, d = 20000, 5000
splits = int(sys.argv[1]) if len(sys.argv) > 1 else 0
path = int(sys.argv[2]) if len(sys.argv) > 2 else 1
opts = {"dense_path": path, "tc_trace": 1}
if splits:
    opts["tc_max_splits"] = splits
be = CudaBackend(dtype="float32", options=opts)
torch = be.torch
X = torch.rand(n, d, device=be.device)
from pycmf_b200.device import DenseMatrix
Xd = DenseMatrix(X)
U = torch.rand(n, 32, device=be.device); V = torch.rand(d, 32, device=be.device)
names = ["tma_issue", "g1_issue", "s_seen", "r_done", "g2_issue", "empty_seen", "full_seen_epi", "qt_start", "qt_done"]
def show(tag):
    NT = 96
    buf = np.zeros(10 * NT + 320, dtype=np.int64)
    _lib.check(be.lib.pycmf_debug_tc_trace(be.ctx, buf.ctypes.data_as(ctypes.c_void_p), buf.size))
    gt = buf[10 * NT:].reshape(160, 2)[:148]
    t = buf[:10 * NT].reshape(10, NT)
    t0 = t[0, 0]
    print("==", tag)
    print("CTA 0: entry %d, setup done %d, end %d (cycles relative to the first TMA issue)" % tuple(t[7, :3] - t0))
    st, en = gt[:, 0] - gt[:, 0].min(), gt[:, 1] - gt[:, 0].min()
    print("per-CTA global timer (ns): start min/median/max %d/%d/%d, end min/median/max %d/%d/%d" % (
        st.min(), np.median(st), st.max(), en.min(), np.median(en), en.max()))
    print("end time by CTA (us): " + " ".join("%.0f" % (e / 1e3) for e in en))
    print("tile " + " ".join("%13s" % s for s in names))
    for it in list(range(0, 12)) + list(range(12, NT, 4)):
        if t[0, it]:
            print("%4d " % it + " ".join("%13d" % (t[e, it] - t0 if t[e, it] else -1) for e in (0, 1, 2, 3, 4, 5, 6, 8, 9)))
for rep in range(2):
    gx, Hx, pr = be.newton_v_xpart(V, U, Xd, 0, d, "linear", 1.0)
show("resid RIGHT")
F = U.clone()
be.newton_left(F, V, Xd, 0.5, 0.0, 0.1, "linear", False, 0.2, False)
show("after newton_left (resid LEFT)")
