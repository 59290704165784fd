#!/usr/bin/env python
"""bench.py -- CMF fit iterations/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W [--workload c5] [--dtype float32] [--impl reference]

A "step" is one full solver iteration (`update_step`: V, U, Z for MU; U, Z, V for Newton) over a synthetic
BASELINE.json workload.  The top-level line is C5 (configs[4]) AT FULL SIZE -- dense X 200000 x 50000 fp32 (40 GB,
the largest configuration that fits one B200 and the one the tensor-core roofline and the strong-scaling target are
stated on), rows of X / U sharded over the N ranks.  `others` carries one entry per further configuration, each with
its own value / e2e / roofline / cpu_baseline / parity: C1 and C2 in full, C3 as a true row shard (250000 rows per
rank = n/8, all 200000 columns: the full 2M x 200k problem at N = 8, "scaling": "weak").

  value : K iterations timed on the device (CUDA events, max over ranks), inputs resident in HBM.
  e2e   : the same metric through the reference-facing seam `solver.fit_iterative_update(X, Y, U, V, Z)` with HOST
          (pinned) arrays: one call of K iterations, H2D of X/Y/U/V/Z and D2H of U/V/Z inside the timed region.
  roofline    : dominant kernel family, timed live with CUDA events by the library's per-family timers; the tensor
                (TF32) and fp64 GEMM peaks are measured in this run with cuBLAS, the HBM peak is MEASURED_PEAKS.json's.
  parity      : product vs the float64 CPU oracle from the same initial factors on that workload (oracle/parity.py).
  cpu_baseline: the UNMODIFIED reference (baseline/_ref, kind "reference"; the oracle port if it is absent) on the
                host cores, rank 0 at N = 1 only, on a bounded sample of the workload.
  --impl reference : the reference arm -- the unmodified reference's solvers on the host cores, same workloads.
"""
import argparse
import importlib.util
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cmf_fit_iterations_per_sec"
UNIT = "it/s"
DEFAULT_OTHERS = "c1,c2,c3,c4"
LTS_CAP_BYTES_PER_CLK = 6300.0      # measured full-chip L2 -> SM throughput cap (guides/B300_MICROARCH.md, "LTS throughput cap")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--others", default=None, help="comma list of further workloads (default c1,c2,c3 next to c5; 'none')")
    ap.add_argument("--scale", type=float, default=None, help="row scale of the headline workload")
    ap.add_argument("--col-scale", type=float, default=None, help="column scale (toy slices of the sparse workloads)")
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--dense-path", type=int, default=None, help="0 generic FMA, 1 tcgen05 3xTF32, 2 tcgen05 1xTF32")
    ap.add_argument("--opt", action="append", default=[], help="backend option key=value (repeatable)")
    ap.add_argument("--v-phase", default="rows", choices=["rows", "columns"],
                    help="N > 1, Newton with per-row Hessians (c4): 'columns' re-partitions the resident row shards into "
                         "column blocks (all-to-all) and runs the column-sharded V phase; default 'rows' = the measured path")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-peaks", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    return ap.parse_args()


def default_scale(name, world):
    """Row / column scale per workload: c3 is a row shard per rank (n/8 rows each, the full problem at 8 ranks); c4 needs
    tensor-core Hessian builds to run at width and is a toy slice."""
    return {"c1": (1.0, 1.0), "c2": (1.0, 1.0), "c3": (min(world, 8) / 8.0, 1.0), "c4": (0.02, 0.02),
            "c5": (1.0, 1.0)}[name]


def load_workloads():
    """pycmf_b200/workloads.py as a stand-alone module (no package import: the reference arm must not load the product)."""
    spec = importlib.util.spec_from_file_location("_cmf_workloads", os.path.join(ROOT, "pycmf_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _reasons(self):
        nv = self.nv
        try:
            bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:  # noqa: BLE001
            bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        return [n for b, n in names.items() if bits & b]

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.reasons.update(self._reasons())
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------- peaks
def measure_gemm_peaks(torch, seconds=1.0):
    """cuBLAS GEMM peaks measured in THIS run (the denominators MEASURED_PEAKS.json does not carry): TF32 8192^3 (best of
    6 = burst; back to back for `seconds` = sustained) and fp64 4096^3.  Library calls, denominators only."""
    out = {}

    def run(dtype, n, tf32, sustained):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            a = torch.randn(n, n, device="cuda", dtype=dtype)
            b = torch.randn(n, n, device="cuda", dtype=dtype)
            c = torch.empty(n, n, device="cuda", dtype=dtype)
            for _ in range(2):
                torch.matmul(a, b, out=c)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(a, b, out=c)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res = {"burst": round(2.0 * n ** 3 / best / 1e9, 1)}
            if sustained > 0:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0, cnt = time.perf_counter(), 0
                e0.record()
                while time.perf_counter() - t0 < sustained:
                    for _ in range(10):
                        torch.matmul(a, b, out=c)
                    cnt += 10
                    torch.cuda.synchronize()
                e1.record()
                torch.cuda.synchronize()
                res["sustained"] = round(2.0 * n ** 3 * cnt / e0.elapsed_time(e1) / 1e9, 1)
            return res
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    out["tf32_tflops"] = run(torch.float32, 8192, True, seconds)
    out["fp64_tflops"] = run(torch.float64, 4096, False, 0.0)
    out["how"] = "torch.matmul (cuBLAS) in this run: TF32 8192^3 best of 6 (burst) and back to back for %.1f s " \
                 "(sustained); fp64 4096^3 best of 6" % seconds
    torch.cuda.empty_cache()
    return out


def load_measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return {}


# ------------------------------------------------------------------------------------ host workload
def pinned_like(torch, t, dtype):
    """Device tensor -> pinned host ndarray of `dtype`, in row blocks (a strided 40 GB X must not be made contiguous on
    the device first)."""
    h = torch.empty(t.shape, dtype=dtype, pin_memory=True)
    if t.dim() == 2 and t.numel() * t.element_size() > (1 << 30):
        step = max(1, (1 << 28) // max(1, t.shape[1] * t.element_size()))
        for r0 in range(0, t.shape[0], step):
            blk = t[r0:r0 + step]
            h[r0:r0 + step].copy_(blk.to(dtype).contiguous() if blk.dtype != dtype else blk.contiguous())
    else:
        h.copy_(t.to(dtype) if t.dtype != dtype else t)
    return h.numpy()


def host_problem(torch, raw, U, V, Z, x_dtype=None, pinned=True):
    """Device problem -> host arrays as a user of the reference would hold them (float64; `x_dtype` float32 for the 40 GB
    X of C5, which a float64 copy would double).  Dense arrays live in pinned memory when `pinned`."""
    import scipy.sparse as sp
    f64 = torch.float64

    def to_host(t, dt=f64):
        if pinned:
            return pinned_like(torch, t, dt)
        return t.detach().to("cpu").to(dt).numpy()

    if raw["csr"] is not None:
        rowptr, colidx, vals = raw["csr"]
        r0, r1 = raw["rows"]
        if pinned:      # the CSR arrays of the user's matrix in pinned memory, like the dense arrays
            parts = (pinned_like(torch, vals, f64), pinned_like(torch, colidx, torch.int32),
                     pinned_like(torch, rowptr, torch.int32))
        else:
            parts = (vals.cpu().numpy().astype(np.float64), colidx.cpu().numpy(), rowptr.cpu().numpy())
        Xh = sp.csr_matrix(parts, shape=(r1 - r0, raw["shape"][1]), copy=False)
        Xh.has_canonical_format = True          # generated sorted and duplicate-free (workloads.py)
    else:
        X = raw["X"].t if hasattr(raw["X"], "t") and not torch.is_tensor(raw["X"]) else raw["X"]
        Xh = to_host(X, x_dtype or f64)
    Y = raw["Y"].t if not torch.is_tensor(raw["Y"]) else raw["Y"]
    return Xh, to_host(Y), to_host(U), to_host(V), to_host(Z)


def make_solver(cfg, params, **kw):
    from pycmf_b200.cmf_solvers import MUSolver, NewtonSolver
    if cfg["solver"] == "mu":
        return MUSolver(tol=0, l1_reg=params.get("l1_reg", 0.), l2_reg=params.get("l2_reg", 0.), **kw)
    p = dict(params)
    p.setdefault("sg_sample_ratio", cfg.get("sg_sample_ratio", 1.0))
    return NewtonSolver(tol=0, x_link=cfg["x_link"], y_link=cfg["y_link"], **p, **kw)


# ----------------------------------------------------------------------------------- CPU step functions
def oracle_step_fn(cfg, params):
    """One CPU iteration of the oracle port on host arrays."""
    from oracle import cmf_oracle as O
    if cfg["solver"] == "mu":
        return lambda X, Y, U, V, Z: O.mu_step(X, Y, U, V, Z, params.get("l1_reg", 0.), params.get("l2_reg", 0.))
    p = dict(params)
    p.setdefault("sg_sample_ratio", cfg.get("sg_sample_ratio", 1.0))
    return lambda X, Y, U, V, Z: O.newton_step(X, Y, U, V, Z, x_link=cfg["x_link"], y_link=cfg["y_link"], **p)


def reference_step_fn(cfg, params):
    """One iteration of the UNMODIFIED reference (baseline/_ref: its own MUSolver / NewtonSolver.update_step,
    cmf_solvers.py:248-263 / :510-522) on host arrays, or None when the install is absent."""
    from oracle.ref_loader import load_reference
    ref = load_reference()
    if ref is None:
        return None
    S = ref.cmf_solvers
    l1, l2, alpha = params.get("l1_reg", 0.), params.get("l2_reg", 0.), params.get("alpha", 0.5)
    if cfg["solver"] == "mu":
        s = S.MUSolver(max_iter=1, tol=0, l1_reg=l1, l2_reg=l2)
    else:
        p = dict(params)
        p.setdefault("sg_sample_ratio", cfg.get("sg_sample_ratio", 1.0))
        s = S.NewtonSolver(max_iter=1, tol=0, x_link=cfg["x_link"], y_link=cfg["y_link"], **p)
    return lambda X, Y, U, V, Z: s.update_step(X, Y, U, V, Z, l1, l2, alpha)


def cpu_step(cfg, params):
    fn = reference_step_fn(cfg, params)
    if fn is not None:
        return fn, "reference"
    return oracle_step_fn(cfg, params), "port"


def threads_used():
    try:
        from threadpoolctl import threadpool_info
        blas = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        return max(blas) if blas else 1
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1: give the BLAS behind NumPy every host core back (the reference is BLAS-bound
    in its products and Python-bound in its per-row loops)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:  # noqa: BLE001
        pass


def row_sample(X, U, rows):
    return X[:rows], U[:rows].copy()


def time_cpu_workload(name, cfg, params, host, budget_s, max_steps, warmup=1):
    """Seconds per FULL-SIZE iteration of the CPU implementation and a description of what was run.

    The reference cannot run C3 / C5 at full size (its MU U update forms the n x d product U V^T, cmf_solvers.py:233:
    80 GB / 3.2 TB of float64), so those are timed on two row samples (n_s and 2 n_s rows of X / U, everything on the
    V side at full size) and extrapolated linearly in n: t(n) = a + b n -- the V-side work a does not shrink with the
    sample, so a plain n / n_s scaling would overstate the reference's time."""
    Xh, Yh, Uh, Vh, Zh = host
    step, kind = cpu_step(cfg, params)
    n_full = cfg["n_full"]
    n_have = Xh.shape[0]

    def run(rows, steps_cap, budget, warm_steps):
        X, U = (Xh, Uh.copy()) if rows >= n_have else row_sample(Xh, Uh, rows)
        if not hasattr(X, "tocsr") and X.dtype != np.float64:
            X = X.astype(np.float64)          # the reference computes in float64 (check_array(dtype=float), cmf.py:386)
        V, Z = Vh.copy(), Zh.copy()
        t0 = time.perf_counter()
        for _ in range(warm_steps):
            step(X, Yh, U, V, Z)
        warm = time.perf_counter() - t0
        times = []
        while len(times) < steps_cap and (len(times) < 1 or warm + sum(times) + np.mean(times) < budget):
            t0 = time.perf_counter()
            step(X, Yh, U, V, Z)
            times.append(time.perf_counter() - t0)
        return float(np.mean(times)), len(times)

    who = "the unmodified reference (baseline/_ref)" if kind == "reference" else "the NumPy/BLAS oracle port"
    sample_rows = {"c3": 2000, "c5": 1000}.get(name)
    if sample_rows is None or 2 * sample_rows >= n_full:
        # C2: one reference iteration takes ~10 s (per-row Python loops): no untimed warm-up iteration
        sec, cnt = run(n_have, max_steps, budget_s, 0 if cfg["solver"] == "newton" else warmup)
        return sec, kind, "%d full-size iteration(s) of %s" % (cnt, who)
    t1, c1 = run(sample_rows, max(2, max_steps // 2), budget_s / 3, warmup)
    t2, c2 = run(2 * sample_rows, max(2, max_steps // 2), 2 * budget_s / 3, warmup)
    b = max(t2 - t1, 0.0) / sample_rows
    a = max(t1 - b * sample_rows, 0.0)
    sec = a + b * n_full
    what = ("row samples of %d and %d of the %d rows (V, Y, Z at full size): %.3f s and %.3f s per iteration (%d + %d "
            "iterations of %s), extrapolated linearly in n: t(n) = %.3f + %.3e n" % (
                sample_rows, 2 * sample_rows, n_full, t1, t2, c1, c2, who, a, b))
    return sec, kind, what


# --------------------------------------------------------------------------------------------- our arm
class Env:
    pass


def bench_workload(env, name, scale, col_scale, steps, warmup, args, headline):
    """One workload on all ranks: returns the result dict on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from pycmf_b200 import workloads as W
    from pycmf_b200.cmf_solvers import FitState
    from pycmf_b200.sharding import row_range

    be, comm, world, rank = env.be, env.comm, env.world, env.rank
    cfg = W.describe(name, scale, col_scale)
    params = W.SOLVER_PARAMS[name]
    n = cfg["n"]
    r0, r1 = row_range(n, rank, world)
    data = W.generate(be, name, r0, r1, scale, col_scale=col_scale)
    xs = torch.tensor([data["x_sum"]], dtype=torch.float64, device=be.device)
    comm.all_reduce_sum(xs)
    x_sum = float(xs.item())
    U, V, Z = W.finish_init(be, data, x_sum)
    U0, V0, Z0 = U.clone(), V.clone(), Z.clone()
    solver = make_solver(cfg, params, dtype=args.dtype, backend=be, comm=comm, max_iter=steps, sharded_input=True,
                         v_phase=args.v_phase)
    xcol, cols = solver._prepare_column_block(be, comm, data["X"], data["X"], cfg["d"], r0)   # None unless --v-phase columns
    st = FitState(be, comm, data["X"], data["Y"], U, V, Z, n, (r0, r1), xcol, cols)
    v_phase_used = "columns" if xcol is not None else "rows"
    del xcol
    obj_first = solver.device_error(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region A (the product path: fit_device's stepper = eager warm-up, then CUDA-graph replay)
    step = solver.make_stepper(st, force_graph=True)
    for i in range(max(warmup, 3)):
        step()
    barrier()
    clocks = ClockSampler(env.local_rank)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clock_info = clocks.finish()
    # ---- timed region B (same iterations launched eagerly with the library's per-kernel-family event timers on:
    #      per-launch kernel durations for the roofline, and the launch count).  The branches that normally run side by
    #      side (U / Z updates, shared-Hessian side streams) are serialised here so that an event pair brackets one family.
    n_prof = min(steps, 20)
    be.set_option("side_streams", 0)
    be.profile(True)
    be.profile_reset()
    launches0 = be.launch_count()
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evp0.record()
    for i in range(n_prof):
        st.iteration += 1
        solver._step(st)
    evp1.record()
    torch.cuda.synchronize()
    ms_prof = evp0.elapsed_time(evp1)
    launches = int(round((be.launch_count() - launches0) / n_prof * steps))
    fams = {}
    for fam in env.families:
        tot, cnt = be.profile_query(fam)
        if cnt:
            fams[fam] = (tot, cnt)
    be.profile(False)
    be.profile_reset()
    be.set_option("side_streams", 1)
    t = torch.tensor([ms], dtype=torch.float64, device=be.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    obj_last = solver.device_error(st)
    value = steps / (ms / 1e3)

    # ---- roofline of the dominant kernel family
    sb = 4 if args.dtype == "float32" else 8
    n_loc, d, l, k = r1 - r0, cfg["d"], cfg["l"], cfg["k"]
    nnz_loc = data["X"].nnz if cfg["sparse"] else None
    roofline = build_roofline(env, name, cfg, fams, n_prof, ms_prof, n_loc, nnz_loc, sb, scale, args, clock_info)
    if roofline is not None and name != "c4":
        # the whole iteration against the same peak: SURVEY 8d's algorithmic flops / bytes of ONE iteration of the full
        # problem (all ranks) over the measured time per iteration (region A), per GPU
        cfull = dict(cfg)
        nnz_all = None
        if cfg["sparse"]:
            tnnz = torch.tensor([float(nnz_loc)], dtype=torch.float64, device=be.device)
            comm.all_reduce_sum(tnnz)
            nnz_all = float(tnnz.item())
        fl, by = W.algorithmic_work(cfull, sb, nnz_all)
        sec = ms / steps / 1e3
        if roofline["bound"] == "tensor":
            ach = fl / sec / 1e12 / world
            roofline["step_level"] = {"achieved": round(ach, 1), "unit": "TFLOP/s per GPU", "frac": round(ach / roofline["peak"], 4)
                                      if roofline.get("peak") else None, "algorithmic_flops_per_iteration": fl}
        else:
            ach = by / sec / 1e9 / world
            roofline["step_level"] = {"achieved": round(ach, 1), "unit": "GB/s per GPU", "frac": round(ach / roofline["peak"], 4)
                                      if roofline.get("peak") else None, "algorithmic_bytes_per_iteration": by}

    # ---- e2e through the solver seam with host buffers
    e2e = None
    host = None
    x_host_dtype = torch.float32 if (not cfg["sparse"] and n_loc * d * 8 > 16e9) else torch.float64
    # C4: the reference draws a fresh permutation and copies X[:, mask] for every one of the 44000 rows of an iteration
    # (cmf_solvers.py:328-344): minutes per iteration even on this slice -- no CPU arm for it
    want_cpu = world == 1 and not args.no_cpu and name != "c4"
    do_e2e = not args.no_e2e and name != "c4"          # (C4: the device-resident figure only)
    if do_e2e or want_cpu:
        host = host_problem(torch, data, U0, V0, Z0, x_dtype=x_host_dtype)
    del st, step, U, V, Z
    x_bytes_dev = (nnz_loc * (sb + 4) * 2 if cfg["sparse"] else n_loc * d * sb)
    if do_e2e:
        del data
        torch.cuda.empty_cache()
        Xh, Yh, Uh, Vh, Zh = host
        s2 = make_solver(cfg, params, dtype=args.dtype, backend=be, comm=comm, max_iter=steps, sharded_input=True)
        # untimed warm-up call (allocator, scratch growth, pinned staging)
        s2.max_iter = min(2, steps)
        s2.fit_iterative_update(Xh, Yh, Uh.copy(), Vh.copy(), Zh.copy())
        s2.max_iter = steps
        Uc, Vc, Zc = Uh.copy(), Vh.copy(), Zh.copy()
        torch.cuda.empty_cache()
        barrier()
        t0 = time.perf_counter()
        s2.fit_iterative_update(Xh, Yh, Uc, Vc, Zc)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=be.device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        xbytes = (Xh.data.nbytes + Xh.indices.nbytes + Xh.indptr.nbytes) if cfg["sparse"] else Xh.nbytes
        h2d = xbytes + Yh.nbytes + Uh.nbytes + Vh.nbytes + Zh.nbytes
        d2h = (Uh.size + Vh.size + Zh.size) * sb
        e2e = {"value": round(steps / dt, 3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d / steps), "d2h_bytes_per_step": int(d2h / steps),
               "call": "fit_iterative_update(X, Y, U, V, Z) with pinned host arrays (X %s, factors float64), max_iter=%d; "
                       "per rank: its own row block" % ("CSR float64" if cfg["sparse"] else str(x_host_dtype).split(".")[1],
                                                        steps),
               "seconds": round(dt, 4), "h2d_bytes_total_per_rank": int(h2d),
               "graph_capture_ms": round(getattr(s2, "capture_seconds_", 0.0) * 1e3, 2)}
        del s2
    else:
        del data
    torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N = 1): the unmodified reference on the host cores, bounded
    cpu = None
    if want_cpu:
        use_all_cores()
        c2 = dict(cfg)
        c2["n_full"] = n
        sec, kind, what = time_cpu_workload(name, c2, params, host, args.cpu_seconds, 10)
        cpu = {"value": float("%.5g" % (1.0 / sec)), "unit": UNIT, "cores": threads_used(), "kind": kind,
               "sample": what + "; host cpu_count=%d" % (os.cpu_count() or 0)}
    del host

    # ---- parity against the oracle on this workload
    parity = None
    if not args.no_parity:
        from oracle.parity import run_parity
        try:
            parity = run_parity(name, be, comm, make_solver, dtype=args.dtype)
        except Exception as e:  # noqa: BLE001   -- reported, never hidden
            parity = {"error": repr(e), "pass": False}
        torch.cuda.empty_cache()

    if rank != 0:
        return None
    return {
        "workload": name, "value": round(value, 3), "unit": UNIT, "ms_per_step": round(ms / steps, 5), "steps": steps,
        "warmup": max(warmup, 3), "scaling": "weak" if name in ("c3", "c4") and scale < 1.0 else "strong",
        "config": W.bench_config(name, scale, col_scale, world),
        "details": {"objective_first": round(obj_first, 6), "objective_last": round(obj_last, 6),
                    "x_shard_mb": round(x_bytes_dev / 1e6, 1), "dense_path": args.dense_path,
                    "cuda_graph": bool(solver._graphable(FitStateProbe(comm, be), True)),
                    "cuda_graph_in_e2e": bool(solver._graphable(FitStateProbe(comm, be))),
                    "v_phase": v_phase_used},
        "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "parity": parity,
    }


class FitStateProbe:
    """What `_graphable` looks at."""

    def __init__(self, comm, be):
        self.comm, self.be = comm, be


def build_roofline(env, name, cfg, fams, n_prof, ms_prof, n_loc, nnz_loc, sb, scale, args, clock_info):
    if not fams:
        return None
    peaks, run_peaks = env.peaks, env.run_peaks
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "MEASURED_PEAKS.json") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    d, l, k = cfg["d"], cfg["l"], cfg["k"]
    # the roofline is reported for the kernel that carries the traffic / flops: the slowest of the passes over X
    streaming = [f for f in fams if (f.startswith("tc_") and f not in ("tc_factor", "tc_ytv")) or f.startswith("resid_")
                 or f.startswith("dmma_resid")
                 or f in ("spmm", "sddmm", "dmma_gemm")]
    big = [f for f in streaming if not f.startswith("resid_")] or streaming
    dom = max(big or fams, key=lambda f: fams[f][0] / fams[f][1])
    tot, cnt = fams[dom]
    if dom in ("row_grad_hess", "safe_solve", "newton_finish_small"):
        alg_bytes = (3 * d * k + d * l + k * k) * sb
    elif cfg["sparse"]:
        alg_bytes = nnz_loc * (sb + 4) + (n_loc + 1) * 4 + (n_loc + min(d, nnz_loc)) * k * sb
    else:
        alg_bytes = (n_loc * d + (n_loc + d) * k + (n_loc if "left" in dom or dom == "tc_xv" else d) * k) * sb
    per_launch_ms = tot / cnt
    achieved = alg_bytes / (per_launch_ms / 1e3) / 1e9
    traffic = None
    if env.world == 1:
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic = tj.get("%s@%s:%s:%s" % (name, scale, args.dtype, dom))
        except Exception:  # noqa: BLE001
            pass
    common = {"traffic": traffic, "avg_launch_ms": round(per_launch_ms, 5), "launches_timed": cnt,
              # device time of this family / device time of all timed families of the step (serialised, like the
              # ncu launch list it is to be compared with; the host-side gaps of the eager region are excluded)
              "share_of_step": round(tot / sum(v[0] for v in fams.values()), 4),
              "timed_in": "region B: %d eager, serialised iterations with per-family CUDA-event timers "
                          "(%.5f ms/step); value is region A (CUDA-graph replay, independent updates on side "
                          "streams)" % (n_prof, ms_prof / n_prof),
              "families_ms_per_step": {f: round(v[0] / n_prof, 5) for f, v in fams.items()}}
    if dom in ("tc_xv", "tc_xtu") and k >= 128 and args.dtype == "float32":
        # wide factors: the dense MU contraction is tensor-core bound.  fp32 accuracy on TF32 tensor cores costs three
        # MMAs per product (3xTF32: hi*hi + hi*lo + lo*hi), so the fp32-equivalent peak is a third of the TF32 peak,
        # which is measured in this run with a cuBLAS TF32 GEMM.  A kernel timed inside a long step (C5: ~50 ms steps,
        # the board runs under its power cap) is held against the sustained figure; the burst figure is listed next to it.
        alg_flops = 2.0 * n_loc * d * k
        tf = (run_peaks or {}).get("tf32_tflops") or {}
        tf_s, tf_b = tf.get("sustained"), tf.get("burst")
        src = "cuBLAS TF32 GEMM measured in this run, sustained / 3 (3xTF32)"
        if not tf_s:
            bf = peaks.get("bf16_tflops_sustained") or 1400.0
            tf_s, tf_b = bf / 2.0, (peaks.get("bf16_tflops") or 1640.0) / 2.0
            src = "MEASURED_PEAKS.json bf16 / 2 / 3 (no in-run TF32 measurement)"
        ach = alg_flops / (per_launch_ms / 1e3) / 1e12
        roofline = {"bound": "tensor", "kernel": dom, "achieved": round(ach, 1), "peak": round(tf_s / 3.0, 1),
                    "peak_source": src, "unit": "TFLOP/s", "frac": round(ach / (tf_s / 3.0), 4),
                    "algorithmic_flops_per_launch": alg_flops, "executed_tf32_tflops": round(3 * ach, 1),
                    "tf32_peak_sustained": tf_s, "tf32_peak_burst": tf_b,
                    "frac_of_burst": round(3 * ach / tf_b, 4) if tf_b else None, "hbm_gbs_of_x": round(achieved, 1)}
    elif dom == "dmma_gemm" and not cfg["sparse"]:
        # float64 path: the two passes over X are DMMA GEMMs (mma.sync.m8n8k4.f64); the family also holds the
        # factor-sized products, so the flops are the step's 4 n d k + 4 d l k over the family's time per step
        fl = 4.0 * n_loc * d * k + 4.0 * d * l * k
        ms_step = tot / n_prof
        pk = ((run_peaks or {}).get("fp64_tflops") or {}).get("burst")
        ach = fl / (ms_step / 1e3) / 1e12
        roofline = {"bound": "tensor", "kernel": dom, "achieved": round(ach, 2), "peak": pk,
                    "peak_source": "cuBLAS fp64 GEMM 4096^3 measured in this run", "unit": "TFLOP/s",
                    "frac": round(ach / pk, 4) if pk else None, "algorithmic_flops_per_step": fl,
                    "hbm_gbs_of_x": round(2.0 * n_loc * d * sb / (ms_step / 1e3) / 1e9, 1)}
    else:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": hbm_peak,
                    "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
                    "algorithmic_bytes_per_launch": alg_bytes}
        if cfg["sparse"] and dom == "spmm":
            # what actually bounds the SpMM: one k * s-byte factor row is gathered from L2 per nonzero (at 0.05 %
            # density no shared-memory tile holds more than one nonzero per staged row)
            gather = nnz_loc * k * sb
            mhz = clock_info.get("sm_mhz") or 1965.0
            cap = LTS_CAP_BYTES_PER_CLK * mhz * 1e6 / 1e9
            roofline["l2_gather"] = {"bytes_per_launch": gather, "achieved_gbs": round(gather / (per_launch_ms / 1e3) / 1e9, 1),
                                     "lts_cap_gbs": round(cap, 1), "frac_of_lts_cap": round(gather / (per_launch_ms / 1e3) / 1e9 / cap, 4),
                                     "note": "L2 -> SM gather of factor rows; cap = 6300 B/clk (guide, measured) x SM clock"}
    roofline.update(common)
    return roofline


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pycmf_b200.device import CudaBackend
    from pycmf_b200.sharding import Comm, TorchComm

    env = Env()
    env.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    env.rank = rank = int(os.environ.get("RANK", "0"))
    env.local_rank = local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        env.comm = TorchComm()
    else:
        env.comm = Comm()
    opts = {}
    if args.dense_path is not None:
        opts["dense_path"] = args.dense_path
    for kv in args.opt:
        key, val = kv.split("=")
        opts[key] = float(val)
    env.be = CudaBackend(device=local_rank, dtype=args.dtype, options=opts)
    env.families = ("resid_left", "resid_right", "gemm", "spmm", "sddmm", "row_grad_hess", "safe_solve",
                    "apply_shared_inverse", "newton_finish_small", "tc_xv", "tc_xtu", "tc_factor", "tc_ytv",
                    "tc_resid_left", "tc_resid_right", "dmma_gemm", "dmma_resid_left", "dmma_resid_right", "mu_fused")
    env.peaks = load_measured_peaks()
    env.run_peaks = None
    if not args.no_peaks:
        # every rank measures (keeps the ranks in step; the GPUs are independent)
        env.run_peaks = measure_gemm_peaks(torch)
    if world > 1:
        dist.barrier()

    scale, col_scale = default_scale(args.workload, world)
    if args.scale is not None:
        scale = args.scale
    if args.col_scale is not None:
        col_scale = args.col_scale
    head = bench_workload(env, args.workload, scale, col_scale, args.steps, args.warmup, args, True)
    others = []
    names = args.others
    if names is None:
        names = DEFAULT_OTHERS if args.workload == "c5" and args.scale is None else "none"
    for nm in [x for x in names.split(",") if x and x != "none"]:
        s, cs = default_scale(nm, world)
        # short steps: more of them, so that the timed region spans several clock samples
        # (c4: a 0.35 s iteration -- a quarter of the steps)
        k_steps = args.steps * (25 if nm == "c1" else 5 if nm == "c2" else 2 if nm == "c3" else 1)
        if nm == "c4":
            k_steps = max(3, args.steps // 4)
        try:
            res = bench_workload(env, nm, s, cs, k_steps, args.warmup, args, False)
        except Exception as e:  # noqa: BLE001 -- a secondary entry must not cost the headline line
            if world > 1:       # several ranks: the others may be inside a collective of this workload -- fail fast
                raise
            res = {"workload": nm, "error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            print("bench: workload %s failed on rank %d: %r" % (nm, rank, e), file=sys.stderr, flush=True)
        if rank == 0:
            others.append(res)

    if rank == 0:
        out = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": head["scaling"], "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64",
            "data": "synthetic", "config": head["config"], "details": head["details"], "clocks": head["clocks"],
            "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"],
            "parity": head["parity"], "peaks": {"measured_in_run": env.run_peaks, "MEASURED_PEAKS.json": {
                key: env.peaks.get(key) for key in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained")}},
            "others": others,
        }
        print(json.dumps(out), flush=True)
    # drop every captured graph (they hold NCCL kernels) before the process group goes away
    env.be.close()
    del env.be
    import gc
    gc.collect()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------- reference arm
def reference_workload(W, torch, device, name, world, steps, warmup, budget_s):
    """The unmodified reference's update_step on the host cores for workload `name` (same generator, same initial
    factors, same solver parameters as our arm)."""
    scale, col_scale = default_scale(name, world)
    cfg = W.describe(name, scale, col_scale)
    params = W.SOLVER_PARAMS[name]
    n = cfg["n"]
    sample_rows = {"c3": 2000, "c5": 1000}.get(name)
    rows = n if sample_rows is None else 2 * sample_rows
    # the first `rows` rows of the workload (the generator is row-block based: they are the same rows our arm sees);
    # the init scaling uses the mean of the generated rows (the full sum would need the whole matrix on this host)
    raw = W.generate_raw(torch, device, name, 0, rows, scale, col_scale, dtype=torch.float32)
    U, V, Z = W.finish_init(torch.float64, raw, raw["x_sum"] * (float(n) / rows))
    host = host_problem(torch, raw, U, V, Z, pinned=False)
    del raw
    if device.type == "cuda":
        torch.cuda.empty_cache()
    c2 = dict(cfg)
    c2["n_full"] = n
    use_all_cores()
    sec, kind, what = time_cpu_workload(name, c2, params, host, budget_s, steps, warmup=min(warmup, 1))
    value = float("%.5g" % (1.0 / sec))
    return {"workload": name, "value": value, "unit": UNIT, "ms_per_step": round(sec * 1e3, 3),
            "scaling": "weak" if name in ("c3", "c4") and scale < 1.0 else "strong",
            "config": W.bench_config(name, scale, col_scale, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads_used(), "kind": kind,
                             "sample": what + "; host cpu_count=%d" % (os.cpu_count() or 0)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_reference(args):
    """The reference arm: the UNMODIFIED reference (baseline/_ref) through its own solver classes on the host cores.
    Rank 0 only; nothing of pycmf_b200 is imported (the workload generator is loaded as a stand-alone file)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import torch
    W = load_workloads()
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if torch.cuda.is_available() \
        else torch.device("cpu")
    names = args.others
    if names is None:
        names = DEFAULT_OTHERS if args.workload == "c5" and args.scale is None else "none"
    try:
        head = reference_workload(W, torch, device, args.workload, world, args.steps, args.warmup, 60.0)
        others = [reference_workload(W, torch, device, nm, world, args.steps, args.warmup, 30.0)
                  for nm in names.split(",") if nm and nm not in ("none", "c4")]      # c4: minutes per reference iteration
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": "reference arm failed: %r" % (e,)}))
        return
    out = {"impl": "reference", "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
           "scaling": head["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": head["config"], "cpu_baseline": head["cpu_baseline"], "e2e": head["e2e"], "gpu_launches": 0,
           "others": others}
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
