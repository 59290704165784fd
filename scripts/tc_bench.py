"""Times the two tcgen05 passes over the C2-sized X alone (CUDA events, 20 launches each) for a set of backend options.
   python scripts/tc_bench.py key=val[,val...] ...      e.g.  tc_prefetch=0,1,2"""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pycmf_b200.device import CudaBackend, DenseMatrix

n, d = 20000, 5000
grid = []
for a in sys.argv[1:]:
    k, v = a.split("=")
    grid.append([(k, float(x)) for x in v.split(",")])
X = None
for combo in itertools.product(*grid) if grid else [()]:
    opts = {"dense_path": 1}
    opts.update(dict(combo))
    be = CudaBackend(dtype="float32", options=opts)
    if X is None:
        X = torch.rand(n, d, device=be.device)
        U = torch.rand(n, 32, device=be.device) * 0.1
        V = torch.rand(d, 32, device=be.device) * 0.1
    Xd = DenseMatrix(X)
    if "copy" not in globals():
        # box calibration: a plain device copy of X (read + write bytes) and the SM clock right after it
        Y = torch.empty_like(X)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            Y.copy_(X)
        e0.record()
        for _ in range(10):
            Y.copy_(X)
        e1.record(); torch.cuda.synchronize()
        copy = 2 * n * d * 4 * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        del Y
        print("calibration: torch copy of X %.0f GB/s (read + write)" % copy, flush=True)
    res = {}
    for name, wl, wr in (("left", True, False), ("right", False, True)):
        for _ in range(3):
            be.resid_pass(U, V, Xd, "linear", want_left=wl, want_right=wr)
        be.profile(True); be.profile_reset()
        for _ in range(20):
            be.resid_pass(U, V, Xd, "linear", want_left=wl, want_right=wr)
        ms, cnt = be.profile_query("tc_resid_" + name)
        be.profile(False)
        res[name] = ms / cnt
    gb = n * d * 4 / 1e9
    print(dict(combo), "left %.1f us (%.0f GB/s)  right %.1f us (%.0f GB/s)" % (
        res["left"] * 1e3, gb / (res["left"] * 1e-3), res["right"] * 1e3, gb / (res["right"] * 1e-3)), flush=True)
    be.close()
