"""Debug helper: per-phase comparison of the CUDA path against the oracle along the oracle's trajectory."""
import sys, os, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import test_gpu_parity as T
from oracle import cmf_oracle as O
from pycmf_b200.cmf_solvers import NewtonSolver

name = sys.argv[1] if len(sys.argv) > 1 else "nt_csr_logit_logit_k20"
dtype = sys.argv[2] if len(sys.argv) > 2 else "float64"
solver, n, d, l, k, sparse, params = T.MID[name]
case = T._mid_case(solver, n, d, l, k, sparse, seed=zlib.crc32(name.encode()) % 1000, **params)
p = dict(case["params"]); p.pop("solver")
X, Y = case["X"], case["Y"]
U, V, Z = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
pert = p.get("hessian_pertubation", 0.2)
for it in range(7):
    for phase in ("U", "Z", "V"):
        s = NewtonSolver(max_iter=1, tol=0, dtype=dtype, update_U=phase == "U", update_Z=phase == "Z",
                         update_V=phase == "V", **p)
        Ug, Vg, Zg = U.copy(), V.copy(), Z.copy()
        s.fit_iterative_update(X, Y, Ug, Vg, Zg)
        Uo, Vo, Zo = U.copy(), V.copy(), Z.copy()
        O.newton_step(X, Y, Uo, Vo, Zo, update_U=phase == "U", update_Z=phase == "Z", update_V=phase == "V",
                      **{k_: v for k_, v in p.items()})
        F_g, F_o = dict(U=Ug, V=Vg, Z=Zg)[phase], dict(U=Uo, V=Vo, Z=Zo)[phase]
        diff = np.abs(F_g - F_o)
        r = diff.max(axis=1).argmax()
        print("it %d phase %s max abs diff %.3e (row %d) |F|max %.3e" % (it, phase, diff.max(), r, np.abs(F_o).max()))
        if diff.max() > 1e-6 * max(1.0, np.abs(F_o).max()):
            # rebuild that row's Hessian on the host and show its spectrum
            if phase == "U":
                a, B, w, l2d = U[r], V, p["alpha"], 0.0 if p.get("x_link") == "logit" else p["l2_reg"]
                link = p.get("x_link", "linear")
            elif phase == "Z":
                a, B, w, l2d, link = Z[r], V, 1 - p["alpha"], p["l2_reg"], p.get("y_link", "linear")
            else:
                a = V[r]; B = None
            if B is not None:
                est = B @ a
                ww = O.d_sigmoid(est) if link == "logit" else np.ones_like(est)
                H = w * (B * ww[:, None]).T @ B + l2d * np.eye(len(a))
                print("   eig:", np.array2string(np.linalg.eigvalsh(H), precision=6))
            print("   got:", np.array2string(F_g[r], precision=6))
            print("   ref:", np.array2string(F_o[r], precision=6))
        U, V, Z = Uo if phase == "U" else U, Vo if phase == "V" else V, Zo if phase == "Z" else Z
