#!/usr/bin/env python
"""bench.py -- CMF fit iterations/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W [--workload c2] [--dtype float32] [--impl reference]

A "step" is one full solver iteration (`update_step`: U, Z, V for Newton; V, U, Z for MU) over the
synthetic workload.  Default workload = BASELINE.json configs[1] (C2: dense 20000x5000 + 5000x50, k=32,
Newton, x linear / y logit), rows of X / U sharded over the N ranks (strong scaling).

  value : K iterations timed on the device (CUDA events, max over ranks), inputs resident in HBM.
  e2e   : the same metric through the reference-facing seam `solver.fit_iterative_update(X, Y, U, V, Z)`
          with HOST (pinned, float64) arrays: one call of K iterations, H2D of X/Y/U/V/Z and D2H of
          U/V/Z inside the timed region (wall clock + device sync, max over ranks).
  roofline    : dominant kernel family, timed live with CUDA events by the library's per-family timers.
  cpu_baseline: the CPU oracle (oracle/cmf_oracle.py, NumPy/BLAS port of the reference's algorithm),
                rank 0 at N=1 only, a bounded number of full-size iterations.
  --impl reference : times that CPU port with all host threads on the same workload.
"""
import argparse
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=None, help="row (and, if sparse, column) scale of the workload")
    ap.add_argument("--col-scale", type=float, default=1.0, help="column scale (toy slices of the sparse workloads)")
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--dense-path", type=int, default=None, help="0 generic FMA, 1 tcgen05 3xTF32, 2 tcgen05 1xTF32")
    ap.add_argument("--opt", action="append", default=[], help="backend option key=value (repeatable)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    return ap.parse_args()


def default_scale(name):
    # c3 / c4 are 8-GPU configurations: on one GPU the default is a 1/8 row slice; c4 additionally needs the
    # tensor-core Hessian kernels (next round) to run at full width, so it is column-scaled too.
    return {"c1": 1.0, "c2": 1.0, "c3": 1.0, "c4": 0.02, "c5": 1.0}[name]


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
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


# ------------------------------------------------------------------------------------ host workload
def host_copy(be, t, pinned=True, dtype=np.float64):
    """Device tensor -> host ndarray (float64 like the reference's arrays) living in pinned memory."""
    import torch
    h = torch.empty(t.shape, dtype=torch.float64 if dtype == np.float64 else torch.float32, pin_memory=pinned)
    h.copy_(t.to(h.dtype))
    return h.numpy()


def make_solver(name, cfg, params, **kw):
    from pycmf_b200.cmf_solvers import MUSolver, NewtonSolver
    if cfg["solver"] == "mu":
        return MUSolver(tol=0, l1_reg=params.get("l1_reg", 0.), l2_reg=params.get("l2_reg", 0.), **kw)
    p = dict(params)
    p.setdefault("sg_sample_ratio", cfg.get("sg_sample_ratio", 1.0))
    return NewtonSolver(tol=0, x_link=cfg["x_link"], y_link=cfg["y_link"], **p, **kw)


def oracle_step_fn(cfg, params):
    """One CPU iteration of the oracle on host arrays (the 'port' of the reference's update_step)."""
    from oracle import cmf_oracle as O
    if cfg["solver"] == "mu":
        return lambda X, Y, U, V, Z: O.mu_step(X, Y, U, V, Z, params.get("l1_reg", 0.), params.get("l2_reg", 0.))
    p = dict(params)
    p.setdefault("sg_sample_ratio", cfg.get("sg_sample_ratio", 1.0))
    return lambda X, Y, U, V, Z: O.newton_step(X, Y, U, V, Z, x_link=cfg["x_link"], y_link=cfg["y_link"], **p)


def time_cpu(step, X, Y, U, V, Z, budget_s, max_steps):
    """Runs 1 warm-up + as many full-size iterations as fit in budget_s (>= 1, <= max_steps)."""
    t0 = time.perf_counter()
    step(X, Y, U, V, Z)
    warm = time.perf_counter() - t0
    times = []
    while len(times) < max_steps and (not times or sum(times) + warm + np.mean(times) < budget_s):
        t0 = time.perf_counter()
        step(X, Y, U, V, Z)
        times.append(time.perf_counter() - t0)
    return times


def threads_used():
    try:
        from threadpoolctl import threadpool_info
        blas = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        return max(blas) if blas else 1
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def host_problem_from_device(be, data, U, V, Z):
    import scipy.sparse as sp
    X = data["X"]
    if X.is_sparse:
        Xh = sp.csr_matrix((be.to_host(X.vals).astype(np.float64), be.to_host(X.colidx), be.to_host(X.rowptr)),
                           shape=X.shape)
    else:
        Xh = host_copy(be, X.t)
    return Xh, host_copy(be, data["Y"].t), host_copy(be, U), host_copy(be, V), host_copy(be, Z)


# --------------------------------------------------------------------------------------------- arms
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pycmf_b200 import workloads as W
    from pycmf_b200.cmf_solvers import FitState
    from pycmf_b200.device import CudaBackend
    from pycmf_b200.sharding import Comm, TorchComm, row_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = TorchComm()
    else:
        comm = Comm()
    opts = {}
    if args.dense_path is not None:
        opts["dense_path"] = args.dense_path
    for kv in args.opt:
        key, val = kv.split("=")
        opts[key] = float(val)
    be = CudaBackend(device=local_rank, dtype=args.dtype, options=opts)
    scale = args.scale if args.scale is not None else default_scale(args.workload)
    cfg = W.describe(args.workload, scale, args.col_scale)
    params = W.SOLVER_PARAMS[args.workload]
    n = cfg["n"]
    r0, r1 = row_range(n, rank, world)
    data = W.generate(be, args.workload, r0, r1, scale, col_scale=args.col_scale)
    xs = torch.tensor([data["x_sum"]], dtype=torch.float64, device=be.device)
    comm.all_reduce_sum(xs)
    U, V, Z = W.finish_init(be, data, float(xs.item()))
    U0, V0, Z0 = U.clone(), V.clone(), Z.clone()
    st = FitState(be, comm, data["X"], data["Y"], U, V, Z, n, (r0, r1))
    solver = make_solver(args.workload, cfg, params, dtype=args.dtype, backend=be, comm=comm, max_iter=args.steps)
    obj_first = solver.device_error(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region A (the product path: fit_device's stepper = eager warm-up, then CUDA-graph replay)
    step = solver.make_stepper(st)
    for i in range(max(args.warmup, 3)):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clock_info = clocks.finish()
    # ---- timed region B (same iterations launched eagerly with the library's per-kernel-family event timers on:
    #      per-launch kernel durations for the roofline, and the launch count)
    n_prof = min(args.steps, 20)
    # the branches that normally run side by side (U / Z updates, shared-Hessian side streams) are serialised here so
    # that an event pair brackets one kernel family only
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
    launches = int(round((be.launch_count() - launches0) / n_prof * args.steps))
    fams = {}
    for fam in ("resid_left", "resid_right", "gemm", "spmm", "sddmm", "row_grad_hess", "safe_solve",
                "apply_shared_inverse", "newton_finish_small", "tc_xv", "tc_xtu", "tc_factor", "tc_ytv", "tc_resid_left",
                "tc_resid_right"):
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
    value = args.steps / (ms / 1e3)

    # ---- roofline of the dominant kernel family
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    sb = 4 if args.dtype == "float32" else 8
    n_loc, d, l, k = r1 - r0, cfg["d"], cfg["l"], cfg["k"]
    roofline = None
    if fams:
        # the roofline is reported for the kernel that carries the HBM traffic: the slowest of the passes over X
        # (every family's time is listed next to it)
        streaming = [f for f in fams if (f.startswith("tc_") and f not in ("tc_factor", "tc_ytv")) or f.startswith("resid_") or f in ("spmm", "sddmm")]
        big = [f for f in streaming if not f.startswith("resid_")] or streaming
        dom = max(big or fams, key=lambda f: fams[f][0] / fams[f][1])
        tot, cnt = fams[dom]
        if dom in ("row_grad_hess", "safe_solve", "newton_finish_small"):
            # factor-sized kernels (latency / ALU bound): bytes = factors in + out, label rows, per-row k x k where used
            alg_bytes = (3 * d * k + d * l + k * k) * sb
        elif cfg["sparse"]:
            nnz = data["X"].nnz
            alg_bytes = nnz * (sb + 4) + (n_loc + 1) * 4 + (n_loc + min(d, nnz)) * k * sb
        else:
            alg_bytes = (n_loc * d + (n_loc + d) * k + (n_loc if "left" in dom or dom == "tc_xv" else d) * k) * sb
        per_launch_ms = tot / cnt
        achieved = alg_bytes / (per_launch_ms / 1e3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
                traffic = tj.get("%s@%s:%s:%s" % (args.workload, scale, args.dtype, dom),
                                 tj.get("%s:%s:%s" % (args.workload, args.dtype, dom)) if scale == 1.0 else None)
        except Exception:  # noqa: BLE001
            pass
        common = {"traffic": traffic, "avg_launch_ms": round(per_launch_ms, 5), "launches_timed": cnt,
                  # device time of this family / device time of all timed families of the step (serialised, like the
                  # ncu launch list it is to be compared with; the host-side gaps of the eager region are excluded)
                  "share_of_step": round(tot / sum(v[0] for v in fams.values()), 4),
                  "timed_in": "region B: %d eager, serialised iterations with per-family CUDA-event timers "
                              "(%.5f ms/step); value is region A (CUDA-graph replay, U / Z updates and shared-Hessian "
                              "branches on side streams)" % (n_prof, ms_prof / n_prof),
                  "families_ms_per_step": {f: round(v[0] / n_prof, 5) for f, v in fams.items()}}
        if dom in ("tc_xv", "tc_xtu") and k >= 128 and args.dtype == "float32":
            # wide factors: the dense MU contraction is tensor-core bound.  fp32 accuracy on TF32 tensor cores costs three
            # MMAs per product (3xTF32: hi*hi + hi*lo + lo*hi), so the fp32-equivalent peak is a third of the TF32 peak;
            # the TF32 peak is taken as half of the measured dense bf16 rate (same pipe, half the elements per clock).
            alg_flops = 2.0 * n_loc * d * k
            bf16_s, bf16_b = peaks.get("bf16_tflops_sustained"), peaks.get("bf16_tflops")
            src = "measured (bf16 sustained / 2 / 3)"
            if not bf16_s:
                bf16_s, bf16_b, src = 1400.0, 1640.0, "fallback (bf16 1400 sustained / 2 / 3)"
            ach = alg_flops / (per_launch_ms / 1e3) / 1e12
            peak = bf16_s / 2.0 / 3.0
            roofline = {"bound": "tensor", "kernel": dom, "achieved": round(ach, 1), "peak": round(peak, 1),
                        "peak_source": src, "unit": "TFLOP/s", "frac": round(ach / peak, 4),
                        "algorithmic_flops_per_launch": alg_flops, "executed_tf32_tflops": round(3 * ach, 1),
                        "tf32_peak_sustained": round(bf16_s / 2.0, 1), "tf32_peak_burst": round((bf16_b or 0) / 2.0, 1),
                        "frac_of_burst": round(3 * ach / (bf16_b / 2.0), 4) if bf16_b else None,
                        "hbm_gbs_of_x": round(achieved, 1)}
        else:
            roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": hbm_peak,
                        "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
                        "algorithmic_bytes_per_launch": alg_bytes}
        roofline.update(common)

    # ---- e2e through the solver seam with host buffers
    e2e = None
    host = None
    if not args.no_e2e:
        st.U.copy_(U0); st.V.copy_(V0); st.Z.copy_(Z0)
        host = host_problem_from_device(be, data, U0, V0, Z0)
        Xh, Yh, Uh, Vh, Zh = host
        s2 = make_solver(args.workload, cfg, params, dtype=args.dtype, backend=be, comm=comm,
                         max_iter=args.steps, sharded_input=True)
        # untimed warm-up call (allocator, scratch growth)
        s2.max_iter = min(2, args.steps)
        s2.fit_iterative_update(Xh, Yh, Uh.copy(), Vh.copy(), Zh.copy())
        s2.max_iter = args.steps
        Uc, Vc, Zc = Uh.copy(), Vh.copy(), Zh.copy()
        barrier()
        t0 = time.perf_counter()
        s2.fit_iterative_update(Xh, Yh, Uc, Vc, Zc)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=be.device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        xbytes = (Xh.data.nbytes + Xh.indices.nbytes + Xh.indptr.nbytes) if cfg["sparse"] else Xh.nbytes   # CSC is built on the device
        h2d = xbytes + Yh.nbytes + Uh.nbytes + Vh.nbytes + Zh.nbytes
        d2h = (Uh.size + Vh.size + Zh.size) * sb
        e2e = {"value": round(args.steps / dt, 3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": int(d2h / args.steps),
               "call": "fit_iterative_update(X, Y, U, V, Z) with pinned float64 host arrays, max_iter=%d" % args.steps,
               "seconds": round(dt, 4),
               "graph_capture_ms": round(getattr(s2, "capture_seconds_", 0.0) * 1e3, 2)}

    # ---- CPU baseline (rank 0, N = 1)
    cpu = None
    if world == 1 and not args.no_cpu:
        if host is None:
            host = host_problem_from_device(be, data, U0, V0, Z0)
        Xh, Yh, Uh, Vh, Zh = host
        step = oracle_step_fn(cfg, params)
        times = time_cpu(step, Xh, Yh, Uh.copy(), Vh.copy(), Zh.copy(), args.cpu_seconds, 10)
        cpu = {"value": round(1.0 / float(np.mean(times)), 5), "unit": UNIT, "cores": threads_used(),
               "kind": "port", "sample": "%d full-size iterations of the NumPy/BLAS oracle after 1 warm-up "
               "(host cpu_count=%d)" % (len(times), os.cpu_count() or 0)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 5), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64",
            "data": "synthetic",
            "config": {"workload": "%s: %s X %dx%d + Y %dx%d, k=%d, solver=%s, x_link=%s, y_link=%s%s" % (
                args.workload, "CSR" if cfg["sparse"] else "dense", n, d, d, l, k, cfg["solver"], cfg["x_link"],
                cfg["y_link"], (", sg=%.2f" % cfg["sg_sample_ratio"]) if "sg_sample_ratio" in cfg else ""),
                "scale": scale, "sharding": "rows of X/U over %d rank(s), V/Z/Y replicated" % world,
                "l2_policy": "inputs larger than L2 (X shard %.0f MB)" % (
                    (data["X"].nnz * (sb + 4) if cfg["sparse"] else n_loc * d * sb) / 1e6),
                "solver_params": params, "objective_first": round(obj_first, 6), "objective_last": round(obj_last, 6),
                "dense_path": args.dense_path, "cuda_graph": bool(solver._graphable(st))},
            "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        # The captured iteration holds NCCL kernels inside live CUDA graphs; tearing the process group down under them
        # was seen to hang.  Everything is synchronised and printed: leave without the interpreter's teardown.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_reference(args):
    """The reference's algorithm on the host cores (NumPy/BLAS oracle port; the Python reference itself cannot
    travel to the GPU box).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from pycmf_b200 import workloads as W
    scale = args.scale if args.scale is not None else default_scale(args.workload)
    cfg = W.describe(args.workload, scale, args.col_scale)
    params = W.SOLVER_PARAMS[args.workload]
    n, d, l, k = cfg["n"], cfg["d"], cfg["l"], cfg["k"]
    try:
        import torch
        from pycmf_b200.device import CudaBackend
        be = CudaBackend(device=int(os.environ.get("LOCAL_RANK", "0")), dtype=args.dtype)
        data = W.generate(be, args.workload, 0, n, scale, col_scale=args.col_scale)
        U, V, Z = W.finish_init(be, data, data["x_sum"])
        host = host_problem_from_device(be, data, U, V, Z)
        del data, be
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": "cannot generate the workload: %r" % (e,)}))
        return
    Xh, Yh, Uh, Vh, Zh = host
    step = oracle_step_fn(cfg, params)
    budget = 150.0
    t0 = time.perf_counter()
    step(Xh, Yh, Uh, Vh, Zh)
    first = time.perf_counter() - t0
    warm_done = 1
    while warm_done < args.warmup and first * (warm_done + 3) < 0.3 * budget:
        step(Xh, Yh, Uh, Vh, Zh)
        warm_done += 1
    times = []
    while len(times) < args.steps and (len(times) < 2 or (sum(times) + first * warm_done + np.mean(times)) < budget):
        t0 = time.perf_counter()
        step(Xh, Yh, Uh, Vh, Zh)
        times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    value = round(1.0 / sec, 5)
    cores = threads_used()
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s X %dx%d + Y %dx%d, k=%d, solver=%s, x_link=%s, y_link=%s" % (
            args.workload, "CSR" if cfg["sparse"] else "dense", n, d, d, l, k, cfg["solver"], cfg["x_link"],
            cfg["y_link"]), "scale": scale, "steps_run": len(times), "warmup_run": warm_done,
            "solver_params": params},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d of %d requested full-size iterations (time budget %.0f s), float64 NumPy/BLAS "
                                   "oracle port of cmf_solvers.py; host cpu_count=%d" % (
                                       len(times), args.steps, budget, os.cpu_count() or 0)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
