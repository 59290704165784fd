"""Times the two tcgen05 MU contractions (X^T U: family tc_xtu, X V: family tc_xv) for wide factors on a C5-shaped slice,
and checks them against torch fp64 on a sub-block.   python scripts/tc_mu_bench.py [n d] [k,k,...] [key=val ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pycmf_b200.device import CudaBackend, DenseMatrix

pos = [a for a in sys.argv[1:] if "=" not in a]
n, d = (int(pos[0]), int(pos[1])) if len(pos) >= 2 else (20000, 50000)
ks = [int(x) for x in pos[2].split(",")] if len(pos) >= 3 else [256, 128, 64]
opts = {"dense_path": 1}
for a in sys.argv[1:]:
    if "=" in a:
        key, v = a.split("=")
        opts[key] = float(v)
be = CudaBackend(dtype="float32", options=opts)
X = torch.rand(n, d, device=be.device)
Xd = DenseMatrix(X)
for k in ks:
    U = torch.rand(n, k, device=be.device)
    V = torch.rand(d, k, device=be.device)
    out = be.empty(d + k, k)
    F = torch.ones(n, k, device=be.device)
    for _ in range(2):
        be.mu_v_partial(Xd, U, out=out)
    # accuracy on a sub-block against float64
    ref = (X[:, :512].double().T @ U.double())
    err_r = float(((out[:512].double() - ref).norm() / ref.norm()).item())
    Fd = F.clone()
    be.mu_left(Fd, V, Xd, 0.0, 0.0)
    G = V.double().T @ V.double()
    got = Fd[:512].double() * (F[:512].double() @ G)
    ref = X[:512].double() @ V.double()
    err_l = float(((got - ref).norm() / ref.norm()).item())
    be.profile(True); be.profile_reset()
    for _ in range(5):
        be.mu_v_partial(Xd, U, out=out)
        Fd.fill_(1.0)
        be.mu_left(Fd, V, Xd, 0.0, 0.0)
    res = {}
    for fam in ("tc_xtu", "tc_xv"):
        ms, cnt = be.profile_query(fam)
        res[fam] = ms / max(cnt, 1)
    be.profile(False)
    flop = 2.0 * n * d * k
    print("k=%d  X^T U %.3f ms (%.1f TFLOP/s fp32-equivalent, %.0f GB/s of X, rel err %.1e)   "
          "X V %.3f ms (%.1f TFLOP/s, %.0f GB/s, rel err %.1e)" % (
              k, res["tc_xtu"], flop / res["tc_xtu"] / 1e9, n * d * 4 / res["tc_xtu"] / 1e6, err_r,
              res["tc_xv"], flop / res["tc_xv"] / 1e9, n * d * 4 / res["tc_xv"] / 1e6, err_l), flush=True)
be.close()
