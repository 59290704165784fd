"""Which phase of the fp32 Newton iteration loses accuracy on C2 at full size? (GPU; oracle = checker)"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycmf_b200 import workloads as W
from pycmf_b200.device import CudaBackend
from pycmf_b200.cmf_solvers import NewtonSolver, FitState
from pycmf_b200.sharding import Comm
from oracle import cmf_oracle as O

rf = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
cfg = W.describe("c2", scale)
p = dict(W.SOLVER_PARAMS["c2"])
for dtype, path in (("float64", 1), ("float32", 0), ("float32", 1)):
    be = CudaBackend(device=0, dtype=dtype, options={"dense_path": path})
    data = W.generate(be, "c2", 0, cfg["n"], scale)
    U, V, Z = W.finish_init(be, data, data["x_sum"])
    f64 = lambda t: t.detach().cpu().numpy().astype(np.float64)
    Xh, Yh = f64(data["X"].t), f64(data["Y"].t)
    Uh, Vh, Zh = f64(U), f64(V), f64(Z)
    st = FitState(be, Comm(), data["X"], data["Y"], U, V, Z, cfg["n"], (0, cfg["n"]))
    for it in range(3):
        for phase in ("U", "Z", "V"):
            s = NewtonSolver(tol=0, x_link="linear", y_link="logit", dtype=dtype, backend=be, max_iter=1,
                             update_U=phase == "U", update_Z=phase == "Z", update_V=phase == "V", **p)
            # same inputs on both sides: the oracle restarts every phase from the GPU's current factors
            Uh, Vh, Zh = f64(st.U), f64(st.V), f64(st.Z)
            st.iteration += 1
            s._step(st)
            O.newton_step(Xh, Yh, Uh, Vh, Zh, x_link="linear", y_link="logit", update_U=phase == "U",
                          update_Z=phase == "Z", update_V=phase == "V", **p)
            got = {"U": st.U, "Z": st.Z, "V": st.V}[phase]
            ref = {"U": Uh, "Z": Zh, "V": Vh}[phase]
            print(dtype, "path", path, "iter", it, "phase", phase, "rel err of the updated factor %.3e" % rf(f64(got), ref),
                  "objective %.6f" % O.compute_error(Xh, Yh, Uh, Vh, Zh, p["alpha"], "linear", "logit"), flush=True)
    be.close()
    del data, st, U, V, Z
    torch.cuda.empty_cache()
