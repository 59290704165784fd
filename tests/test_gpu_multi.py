"""Row-sharded fits over 2 GPUs (NCCL) must reproduce the single-GPU result (shard-count invariance)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
from helpers import load_golden, draw_masks_for_case, rel_fro
from test_gpu_parity import _mid_case, MID, NT
import zlib
from pycmf_b200.cmf_solvers import MUSolver, NewtonSolver
MUSolver.SHARD_V_MIN = 0          # exercise the row-sharded V update on the small test shapes
MUSolver.V_SLABS = (4, 5, 3, 2)   # and its slab-overlapped variant (opt-in in the product)
from pycmf_b200.sharding import TorchComm, Comm

def run(case, dtype, comm, masks=None, dense_path=0, history=True, **extra):
    p = dict(case["params"]); solver = p.pop("solver")
    p.update(extra)
    cls = MUSolver if solver == "mu" else NewtonSolver
    # dense_path 0: both shard counts use the same (FMA) arithmetic, so only the summation order differs; the tcgen05
    # path switches on above a size threshold and would be compared against the FMA path on the smaller shards
    # v_phase='rows': this worker covers the row-sharded partial-Hessian exchange; the column-sharded V phase ('auto' picks
    # it for per-row Hessians when every rank holds the whole host X) has its own worker below
    s = cls(max_iter=case["iters"], tol=0, random_state=case["rng_seed"], dtype=dtype, comm=comm,
            backend_options={"dense_path": dense_path}, v_phase="rows", **p)
    s.history = [] if history else None; s.masks_per_iter = masks
    U, V, Z = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
    s.fit_iterative_update(case["X"], case["Y"], U, V, Z)
    return np.asarray(s.history if history else []), U, V, Z

names = ["mu_dense", "mu_csr_reg", "nt_lin_logit", "nt_logit_logit", "nt_csr_lin_logit", "nt_sg_logit_logit"]
for name in names:
    case, g = load_golden(name)
    masks = draw_masks_for_case(case)
    hist, U, V, Z = run(case, "float64", TorchComm(), masks)
    assert np.allclose(hist, g["objective"][1:], rtol=1e-9, atol=1e-11), name
    for got, ref in ((U, g["U"]), (V, g["V"]), (Z, g["Z"])):
        assert rel_fro(got, ref) < 1e-9, name
for name in ["mu_dense_k32", "nt_signed_l1_lin_logit_k32", "mu_csr_k64"]:   # contractive trajectories only
    solver, n, d, l, k, sparse, params = MID[name]
    case = _mid_case(solver, n, d, l, k, sparse, seed=zlib.crc32(name.encode()) % 1000, **params); case["iters"] = 4
    h2, U2, V2, Z2 = run(case, "float32", TorchComm())
    h1, U1, V1, Z1 = run(case, "float32", Comm())
    assert np.abs(h2 - h1).max() / np.abs(h1).max() < 1e-5, (name, h1, h2)
    for a, b in ((U1, U2), (V1, V2), (Z1, Z2)):
        assert rel_fro(a, b) < 1e-4, name
# tensor-core MU contractions (k = 64) on both shard counts: every shard is large enough for the tcgen05 path
case = _mid_case("mu", 2000, 600, 10, 64, False, seed=7); case["iters"] = 4
h2, U2, V2, Z2 = run(case, "float32", TorchComm(), dense_path=1)
h1, U1, V1, Z1 = run(case, "float32", Comm(), dense_path=1)
assert np.abs(h2 - h1).max() / np.abs(h1).max() < 1e-4, ("mu_tc_k64", h1, h2)
for a, b in ((U1, U2), (V1, V2), (Z1, Z2)):
    assert rel_fro(a, b) < 1e-3, "mu_tc_k64"
# the same fit through the default stepper: two eager iterations, then CUDA-graph replay with the NCCL collectives (and the
# communication-stream overlap of the V update) captured inside the graph -- must equal the eager, per-iteration run
case["iters"] = 7
_, Ug, Vg, Zg = run(case, "float32", TorchComm(), dense_path=1, history=False, use_cuda_graph=True)
_, Ue, Ve, Ze = run(case, "float32", TorchComm(), dense_path=1)
for a, b in ((Ug, Ue), (Vg, Ve), (Zg, Ze)):
    assert rel_fro(a, b) < 1e-6, "graph replay vs eager on 2 ranks"
# one slab (no overlap) and four slabs give the same V update up to summation order
MUSolver.V_SLABS = (1,)
_, U1s, V1s, Z1s = run(case, "float32", TorchComm(), dense_path=1)
MUSolver.V_SLABS = (4, 5, 3, 2)
for a, b in ((U1s, Ue), (V1s, Ve), (Z1s, Ze)):
    assert rel_fro(a, b) < 1e-5, "slab-overlapped V update vs single pass"
# on-device sampler: the sample sets are keyed by the GLOBAL row, so 2 ranks draw what 1 rank draws
solver, n, d, l, k, sparse, params = MID["nt_signed_l1_lin_logit_k32"]
case = _mid_case(solver, 600, 200, 6, 16, False, seed=3, **params); case["iters"] = 3
case["params"]["sg_sample_ratio"] = 0.5
h2, U2, V2, Z2 = run(case, "float64", TorchComm(), sampler="device")
h1, U1, V1, Z1 = run(case, "float64", Comm(), sampler="device")
assert np.abs(h2 - h1).max() / np.abs(h1).max() < 1e-9, ("device sampler, 2 ranks vs 1", h1, h2)
for a, b in ((U1, U2), (V1, V2), (Z1, Z2)):
    assert rel_fro(a, b) < 1e-9, "device sampler, 2 ranks vs 1"
dist.barrier(); dist.destroy_process_group()
if rank == 0: print("MULTI_OK")
'''


def test_two_gpu_row_sharding_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29600 + os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", port, str(script), ROOT]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    errors = [ln for ln in out.stdout.splitlines() if "Error" in ln and "ChildFailed" not in ln]
    assert out.returncode == 0 and "MULTI_OK" in out.stdout, "\n".join(errors[:6]) or out.stdout[-3000:]


_COLUMNS_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
from helpers import load_golden, draw_masks_for_case, rel_fro
from test_gpu_parity import _mid_case, MID
from pycmf_b200.cmf_solvers import NewtonSolver
from pycmf_b200.sharding import TorchComm, Comm

def run(case, dtype, comm, masks=None, **extra):
    p = dict(case["params"]); p.pop("solver"); p.update(extra)
    s = NewtonSolver(max_iter=case["iters"], tol=0, random_state=case["rng_seed"], dtype=dtype, comm=comm, **p)
    s.history = []; s.masks_per_iter = masks
    U, V, Z = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
    s.fit_iterative_update(case["X"], case["Y"], U, V, Z)
    return np.asarray(s.history), U, V, Z

# the reference's trajectories (golden = unmodified reference), float64, per-row Hessians: dense / CSR, logit x link, sampled
for name in ["nt_logit_logit", "nt_csr_logit_lin", "nt_sg_logit_logit", "nt_sg_csr_lin_logit", "nt_sg_zero_ysample"]:
    case, g = load_golden(name)
    hist, U, V, Z = run(case, "float64", TorchComm(), draw_masks_for_case(case), v_phase="columns")
    assert np.allclose(hist, g["objective"][1:], rtol=1e-9, atol=1e-11), name
    for got, ref in ((U, g["U"]), (V, g["V"]), (Z, g["Z"])):
        assert rel_fro(got, ref) < 1e-9, name
# mid-size float32 (k = 72: tensor-core Hessians, tridiagonal clamped solve): column-sharded == row-sharded == one rank
solver, n, d, l, k, sparse, params = MID["nt_logit_lin_k72"]
case = _mid_case(solver, n, d, l, k, sparse, seed=11, **params); case["iters"] = 3
hc, Uc, Vc, Zc = run(case, "float32", TorchComm(), v_phase="columns")
hr, Ur, Vr, Zr = run(case, "float32", TorchComm(), v_phase="rows")
h1, U1, V1, Z1 = run(case, "float32", Comm())
for h in (hc, hr):                          # the float32 parity bars (objective 1e-4, factors 1e-3)
    assert np.abs(h - h1).max() / np.abs(h1).max() < 1e-4, ("columns / rows vs one rank", h1, hr, hc)
for a, b in ((Uc, U1), (Vc, V1), (Zc, Z1), (Ur, U1), (Vr, V1), (Zr, Z1)):
    assert rel_fro(a, b) < 1e-3, "column- / row-sharded V phase vs one rank"
# device sampler: the sets are keyed by the GLOBAL row of V, so the column-sharded phase draws what one rank draws
solver, n, d, l, k, sparse, params = MID["nt_signed_l1_lin_logit_k32"]
case = _mid_case(solver, 600, 201, 6, 16, False, seed=3, **params); case["iters"] = 3      # d odd: uneven column blocks
case["params"]["sg_sample_ratio"] = 0.5
h2, U2, V2, Z2 = run(case, "float64", TorchComm(), sampler="device", v_phase="columns")
h1, U1, V1, Z1 = run(case, "float64", Comm(), sampler="device")
assert np.abs(h2 - h1).max() / np.abs(h1).max() < 1e-9, ("device sampler, columns on 2 ranks vs 1", h1, h2)
for a, b in ((U1, U2), (V1, V2), (Z1, Z2)):
    assert rel_fro(a, b) < 1e-9, "device sampler, columns on 2 ranks vs 1"
dist.barrier(); dist.destroy_process_group()
if rank == 0: print("COLUMNS_OK")
"""


def test_two_gpu_column_sharded_newton_v_phase(tmp_path):
    """for per-row Hessians (logit x link / sg < 1) the V phase re-partitions -- all-gather U, every rank updates
    its d / G rows of V against its column block of X, all-gather V -- instead of all-reducing d k^2 Hessian entries."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker_columns.py"
    script.write_text(_COLUMNS_WORKER)
    port = str(27600 + os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", port, str(script), ROOT]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    errors = [ln for ln in out.stdout.splitlines() if "Error" in ln and "ChildFailed" not in ln]
    assert out.returncode == 0 and "COLUMNS_OK" in out.stdout, "\n".join(errors[:6]) or out.stdout[-3000:]


_REPART_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
from helpers import load_golden, draw_masks_for_case, rel_fro
from pycmf_b200.cmf_solvers import NewtonSolver
from pycmf_b200.device import CudaBackend
from pycmf_b200.sharding import TorchComm, row_range
comm = TorchComm()
world = comm.world
for name in ["nt_logit_logit", "nt_sg_csr_lin_logit", "nt_csr_logit_lin"]:
    case, g = load_golden(name)
    p = dict(case["params"]); p.pop("solver")
    n, d = case["X"].shape
    r0, r1 = row_range(n, rank, world)
    # (a) every rank holds its rows only: the resident row shards are re-partitioned by one all-to-all over NVLink
    s = NewtonSolver(max_iter=case["iters"], tol=0, random_state=case["rng_seed"], dtype="float64", comm=comm,
                     sharded_input=True, v_phase="columns", **p)
    s.history, s.masks_per_iter = [], draw_masks_for_case(case)
    U, V, Z = case["U0"][r0:r1].copy(), case["V0"].copy(), case["Z0"].copy()
    s.fit_iterative_update(case["X"][r0:r1], case["Y"], U, V, Z)
    assert np.allclose(s.history, g["objective"][1:], rtol=1e-9, atol=1e-11), (name, "all-to-all")
    for got, ref in ((U, g["U"][r0:r1]), (V, g["V"]), (Z, g["Z"])):
        assert rel_fro(got, ref) < 1e-9, (name, "all-to-all")
    # (b) the whole matrix already in HBM on every rank (device initialisation): the column block is a view
    be = CudaBackend(dtype="float64")
    s = NewtonSolver(max_iter=case["iters"], tol=0, random_state=case["rng_seed"], dtype="float64", comm=comm,
                     backend=be, v_phase="columns", **p)
    s.history, s.masks_per_iter = [], draw_masks_for_case(case)
    U, V, Z = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
    s.fit_iterative_update(be.ingest(case["X"]), case["Y"], U, V, Z)
    assert np.allclose(s.history, g["objective"][1:], rtol=1e-9, atol=1e-11), (name, "resident")
    for got, ref in ((U, g["U"]), (V, g["V"]), (Z, g["Z"])):
        assert rel_fro(got, ref) < 1e-9, (name, "resident")
dist.barrier(); dist.destroy_process_group()
if rank == 0: print("REPART_OK")
"""


def test_two_gpu_column_phase_from_resident_row_shards(tmp_path):
    """The device-side sources of the column block (explicit v_phase='columns'): per-rank row shards re-partitioned by an
    all-to-all under NCCL (sharded_input=True), and a matrix already resident on every rank (a view)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker_repart.py"
    script.write_text(_REPART_WORKER)
    port = str(25600 + os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", port, str(script), ROOT]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    errors = [ln for ln in out.stdout.splitlines() if "Error" in ln and "ChildFailed" not in ln]
    assert out.returncode == 0 and "REPART_OK" in out.stdout, "\n".join(errors[:6]) or out.stdout[-3000:]
