"""Parity of the CUDA fit loop against the CPU oracle ON THE BASELINE CONFIGS -- TEST INFRASTRUCTURE.

`run_parity(name, ...)` generates BASELINE.json workload `name` with bench.py's own generator
(`pycmf_b200/workloads.py`), runs the product solver on the GPU(s) and the float64 oracle
(`oracle/cmf_oracle.py`, pinned against the unmodified reference by tests/golden) on the host from the SAME initial
factors, and returns the north_star figures: max relative error of the per-iteration objective and the relative
Frobenius error of the final U, V, Z.  C1 and C2 run at full size; C3 / C4 / C5 are row-scaled (C4 also column-scaled)
to what the oracle finishes in seconds -- the scale is part of the result.  Used by bench.py's `parity` block and by
tests/test_gpu_configs.py; the oracle is the checker here, never the thing measured.
"""
import numpy as np
import scipy.sparse as sp

from . import cmf_oracle as O

# name -> (row scale, column scale, iterations): what the oracle finishes in a few seconds
PARITY_SHAPES = {
    "c1": (1.0, 1.0, 100),        # 1000 x 500 in full, the 100 iterations BASELINE.json names
    "c2": (1.0, 1.0, 3),          # 20000 x 5000 in full
    "c3": (0.01, 1.0, 3),         # 20000 rows of the 2M, all 200000 columns (V at full size)
    "c4": (0.001, 0.01, 2),       # 2000 x 2000, k = 128, sg = 0.1 with the reference's NumPy masks injected
    "c5": (0.01, 1.0, 3),         # 2000 rows of the 200000, all 50000 columns (V, Y, Z at full size)
}
BARS = {"float32": {"objective_rel": 1e-4, "factor_rel_fro": 1e-3},
        "float64": {"objective_rel": 1e-9, "factor_rel_fro": 1e-9}}


def rel_fro(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-300))


def _host_problem(raw, U, V, Z):
    """float64 host copies of the (fp32-valued) device problem: the oracle sees exactly the numbers the GPU sees."""
    f64 = lambda t: t.detach().to("cpu").numpy().astype(np.float64)  # noqa: E731
    if raw["csr"] is not None:
        rowptr, colidx, vals = raw["csr"]
        n, d = raw["rows"][1] - raw["rows"][0], raw["shape"][1]
        X = sp.csr_matrix((f64(vals), colidx.cpu().numpy(), rowptr.cpu().numpy()), shape=(n, d))
    else:
        X = f64(raw["X"])
    return X, f64(raw["Y"]), f64(U), f64(V), f64(Z)


def run_parity(name, be, comm, make_solver, dtype="float32", shape=None, seed=1234):
    """Returns the parity dict on rank 0 (None elsewhere).  `make_solver(cfg, params, **kw)` builds the product solver."""
    import torch
    from pycmf_b200 import workloads as W
    from pycmf_b200.cmf_solvers import FitState
    from pycmf_b200.sharding import row_range

    scale, col_scale, iters = shape or PARITY_SHAPES[name]
    cfg = W.describe(name, scale, col_scale)
    params = dict(W.SOLVER_PARAMS[name])
    n, d, l, k = cfg["n"], cfg["d"], cfg["l"], cfg["k"]
    r0, r1 = row_range(n, comm.rank, comm.world)
    data = W.generate(be, name, r0, r1, scale, seed=seed, col_scale=col_scale)
    xs = torch.tensor([data["x_sum"]], dtype=torch.float64, device=be.device)
    comm.all_reduce_sum(xs)
    x_sum = float(xs.item())
    U, V, Z = W.finish_init(be, data, x_sum)

    ratio = cfg.get("sg_sample_ratio", 1.0)
    masks = None
    if ratio < 1.0:
        np.random.seed(0)               # every rank draws the same stream, in the reference's call order
        masks = [O.draw_newton_masks(n, d, l, ratio) for _ in range(iters)]

    host = None
    if comm.rank == 0:
        if comm.world == 1:
            raw_all, U_all = data, U
            raw_all = dict(data)
            if data["csr"] is None:
                raw_all["X"] = data["X"].t
            raw_all["Y"] = data["Y"].t
        else:
            raw_all = W.generate_raw(torch, be.device, name, 0, n, scale, col_scale, seed, dtype=be.tdtype)
            U_all, _, _ = W.finish_init(be, raw_all, x_sum)
        host = _host_problem(raw_all, U_all, V, Z)
        del raw_all, U_all

    st = FitState(be, comm, data["X"], data["Y"], U, V, Z, n, (r0, r1))
    solver = make_solver(cfg, params, dtype=dtype, backend=be, comm=comm, max_iter=iters)
    solver.history = []
    solver.masks_per_iter = masks
    solver.fit_device(st)
    U_gpu = comm.all_gather_rows(st.U, n)
    if comm.rank != 0:
        return None

    Xh, Yh, Uh, Vh, Zh = host
    hist = []
    kw = dict(params)
    kw.pop("sg_sample_ratio", None)
    if cfg["solver"] == "mu":
        O.fit_iterative_update(Xh, Yh, Uh, Vh, Zh, solver="mu", max_iter=iters, tol=0, history=hist,
                               l1_reg=kw.get("l1_reg", 0.), l2_reg=kw.get("l2_reg", 0.))
    else:
        O.fit_iterative_update(Xh, Yh, Uh, Vh, Zh, solver="newton", max_iter=iters, tol=0, history=hist,
                               x_link=cfg["x_link"], y_link=cfg["y_link"], sg_sample_ratio=ratio,
                               masks_per_iter=masks, **kw)
    got = np.asarray(solver.history, dtype=np.float64)
    ref = np.asarray(hist, dtype=np.float64)
    obj_err = float(np.max(np.abs(got - ref) / np.abs(ref)))
    fro = {"U": rel_fro(be.to_host(U_gpu), Uh), "V": rel_fro(be.to_host(st.V), Vh), "Z": rel_fro(be.to_host(st.Z), Zh)}
    bars = BARS[str(np.dtype(dtype))]
    ok = obj_err <= bars["objective_rel"] and max(fro.values()) <= bars["factor_rel_fro"]
    return {"against": "oracle/cmf_oracle.py (float64 NumPy restatement, pinned to the unmodified reference by tests/golden)",
            "workload": W.label(name, cfg), "row_scale": scale, "col_scale": col_scale, "iterations": iters,
            "n_ranks": comm.world, "objective_max_rel_err": obj_err, "objective_last": float(got[-1]),
            "objective_last_oracle": float(ref[-1]),
            "factor_rel_fro": {key: float("%.3e" % val) for key, val in fro.items()},
            "bars": bars, "pass": bool(ok)}
