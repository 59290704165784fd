"""CPU tests of the host side: C-ABI exports, solver orchestration against the golden fixtures through a
NumPy stand-in backend, row sharding under gloo (world_size 2), API validation and initialisation."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import CASES, draw_masks_for_case, load_golden, rel_fro
from fake_backend import FakeBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_every_declared_symbol():
    from pycmf_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pycmf_b200.h")).read()
    declared = set(re.findall(r"\b(pycmf_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    assert os.path.exists(_lib.LIB_PATH), "run python -m pycmf_b200._build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.pycmf_abi_version() == _lib.ABI_VERSION


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pycmf_b200 import CMF
    from pycmf_b200._lib import BackendError
    X, Y = np.abs(np.random.RandomState(0).randn(6, 5)), np.abs(np.random.RandomState(1).randn(5, 3))
    with pytest.raises(BackendError, match="no CPU fallback"):
        CMF(n_components=2, max_iter=2).fit_transform(X, Y)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pycmf_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle|/oracle", src, re.M), fn


def _run_fake(case, masks=None, comm=None, **extra):
    from pycmf_b200.cmf_solvers import MUSolver, NewtonSolver
    MUSolver.SHARD_V_MIN = 0          # exercise the row-sharded V update on the small test shapes
    p = dict(case["params"])
    solver = p.pop("solver")
    cls = MUSolver if solver == "mu" else NewtonSolver
    s = cls(max_iter=case["iters"], tol=0, random_state=case["rng_seed"], dtype="float64", backend=FakeBackend(),
            comm=comm, **p, **extra)
    s.history, s.masks_per_iter = [], masks
    U, V, Z = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
    e0 = s.compute_error(case["X"], case["Y"], U, V, Z)
    s.fit_iterative_update(case["X"], case["Y"], U, V, Z)
    return np.asarray([e0] + s.history), U, V, Z


@pytest.mark.parametrize("name", sorted(CASES))
def test_solver_orchestration_matches_golden(name):
    """MUSolver / NewtonSolver phase decomposition (partial -> all-reduce -> apply; x-part -> finish, chunked
    V rows) reproduces the reference trajectories when the phases are computed exactly."""
    case, g = load_golden(name)
    hist, U, V, Z = _run_fake(case, draw_masks_for_case(case))
    assert np.allclose(hist, g["objective"], rtol=1e-9, atol=1e-11)
    for got, ref in ((U, g["U"]), (V, g["V"]), (Z, g["Z"])):
        assert rel_fro(got, ref) < 1e-9


def test_numpy_sampler_follows_reference_rng_stream():
    """Without injected masks the solver draws from np.random in the reference's order (golden = reference)."""
    case, g = load_golden("nt_sg_lin_lin")
    hist, U, V, Z = _run_fake(case, masks=None)
    assert np.allclose(hist, g["objective"], rtol=1e-9, atol=1e-11)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch.distributed as dist
from helpers import load_golden, draw_masks_for_case, rel_fro
from test_host_logic import _run_fake
from pycmf_b200.sharding import TorchComm
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
extra = {"v_phase": sys.argv[5]} if len(sys.argv) > 5 else {}
from pycmf_b200.cmf_solvers import NewtonSolver
calls, inner = [0], NewtonSolver._step_v_columns
def counted(self, *a):
    calls[0] += 1
    return inner(self, *a)
NewtonSolver._step_v_columns = counted
for name in sys.argv[4].split(","):
    case, g = load_golden(name)
    calls[0] = 0
    hist, U, V, Z = _run_fake(case, draw_masks_for_case(case), comm=TorchComm(), **extra)
    p = case["params"]
    per_row = p["solver"] == "newton" and p.get("update_V", True) and \
        (p.get("x_link", "linear") == "logit" or p.get("sg_sample_ratio", 1.) < 1.)
    assert calls[0] == (case["iters"] if extra.get("v_phase") != "rows" and per_row else 0), (name, calls[0])
    assert np.allclose(hist, g["objective"], rtol=1e-9, atol=1e-11), name
    for got, ref in ((U, g["U"]), (V, g["V"]), (Z, g["Z"])):
        assert rel_fro(got, ref) < 1e-9, name
dist.destroy_process_group()
print("OK")
'''


def test_row_sharding_world2_gloo_is_shard_count_invariant(tmp_path):
    names = "mu_dense,mu_csr_reg,nt_lin_logit,nt_logit_logit,nt_csr_lin_logit,nt_sg_logit_logit,nt_sg_zero_ysample"
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), names, "rows"],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "OK" in o, o[-3000:]


@pytest.mark.parametrize("mode", ["columns", "auto"])
def test_column_sharded_newton_v_phase_world2_gloo(tmp_path, mode):
    """SURVEY 8e, Newton with a logit x link / sg < 1: for the V phase every rank owns d / 2 rows of V and the matching
    column block of X over all rows, U is all-gathered and the new V rows are all-gathered (v_phase='columns').  Must
    reproduce the reference trajectories exactly like the row-sharded partial-Hessian all-reduce does; the cases with a
    shared Hessian (linear x link, sg = 1) must not re-partition at all."""
    names = ("nt_logit_logit,nt_logit_lin,nt_csr_logit_lin,nt_sg_logit_logit,nt_sg_lin_lin,nt_sg_csr_lin_logit,"
             "nt_sg_zero_ysample,nt_lin_logit,nt_no_V,mu_dense")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(33500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), names, mode],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "OK" in o, o[-3000:]


_REPART_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
rank, world = int(sys.argv[3]), int(sys.argv[4])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=world)
from pycmf_b200.device import DenseMatrix, SparseMatrix, column_block_from_row_shards
from pycmf_b200.sharding import TorchComm, row_range
from helpers import load_golden, draw_masks_for_case, rel_fro
from fake_backend import FakeBackend
from pycmf_b200.cmf_solvers import NewtonSolver
comm = TorchComm()

def csc_arrays(A):                       # the column access path of a SparseMatrix, as device.py builds it
    C = sp.csc_matrix(A); C.sort_indices()
    return (torch.from_numpy(C.indptr.astype(np.int32)), torch.from_numpy(C.indices.astype(np.int32)),
            torch.from_numpy(C.data.astype(np.float64)))

# 1. the torch-level re-partition itself (CPU tensors): dense and CSR, uneven row and column blocks, an empty column
rng = np.random.RandomState(5)
for n, d in ((23, 11), (8, 5), (6, 3)):
    A = rng.randn(n, d) * (rng.rand(n, d) < 0.4)
    A[:, d // 2] = 0.0
    if n == 8:
        A[:row_range(n, 0, world)[1]] = 0.0                  # rank 0's shard has no nonzeros at all
    r0, r1 = row_range(n, rank, world)
    ranges = [row_range(d, g, world) for g in range(world)]
    c0, c1 = ranges[rank]
    blk = column_block_from_row_shards(DenseMatrix(torch.from_numpy(A[r0:r1].copy())), comm, r0, ranges)
    assert blk.shape == (n, c1 - c0) and np.array_equal(blk.t.numpy(), A[:, c0:c1]), "dense"
    S = SparseMatrix.__new__(SparseMatrix)
    S.shape = (r1 - r0, d)
    S.rowptr = S.colidx = S.vals = None
    S.colptr, S.rowidx, S.cvals = csc_arrays(A[r0:r1])
    if S.cvals.numel() == 0:                                 # as SparseMatrix stores an empty shard: one dummy element
        S.rowidx, S.cvals = torch.zeros(1, dtype=torch.int32), torch.zeros(1, dtype=torch.float64)
    blk = column_block_from_row_shards(S, comm, r0, ranges)
    want = csc_arrays(A[:, c0:c1])
    assert blk.shape == (n, c1 - c0) and blk.nnz == want[2].numel(), "sparse shape"
    for got, ref in zip((blk.colptr, blk.rowidx[:blk.nnz], blk.cvals[:blk.nnz]), want):
        assert got.dtype == ref.dtype and torch.equal(got, ref), "sparse arrays"

# 2. the solver on per-rank row blocks (sharded_input=True): explicit v_phase='columns' re-partitions the resident shards
for name in sys.argv[5].split(","):
    case, g = load_golden(name)
    p = dict(case["params"]); p.pop("solver")
    n, d = case["X"].shape
    # the caller's own, UNEVEN row blocks (sharded_input does not require row_range's balanced split)
    cuts = [0] + [int(n * f) for f in ((0.65,) if world == 2 else (0.5, 0.6))] + [n]
    r0, r1 = cuts[rank], cuts[rank + 1]
    for mode in ("columns", "auto"):
        s = NewtonSolver(max_iter=case["iters"], tol=0, random_state=case["rng_seed"], dtype="float64",
                         backend=FakeBackend(), comm=comm, sharded_input=True, v_phase=mode, **p)
        s.history, s.masks_per_iter = [], draw_masks_for_case(case)
        seen = []
        inner = s._step_v_columns
        s._step_v_columns = lambda *a: (seen.append(1), inner(*a))[1]
        U, V, Z = case["U0"][r0:r1].copy(), case["V0"].copy(), case["Z0"].copy()
        s.fit_iterative_update(case["X"][r0:r1], case["Y"], U, V, Z)
        assert len(seen) == case["iters"], (name, mode, len(seen))
        assert np.allclose(s.history, g["objective"][1:], rtol=1e-9, atol=1e-11), (name, mode)
        for got, ref in ((U, g["U"][r0:r1]), (V, g["V"]), (Z, g["Z"])):
            assert rel_fro(got, ref) < 1e-9, (name, mode)
dist.destroy_process_group()
print("OK")
'''


@pytest.mark.parametrize("world", [2, 3])
def test_row_shards_to_column_blocks_all_to_all_gloo(tmp_path, world):
    """The device-side source of the column-sharded V phase: row shards resident on the ranks are re-partitioned into
    column blocks by one all-to-all (pycmf_b200.device.column_block_from_row_shards, torch only) -- checked on CPU tensors
    against scipy slicing -- and the solver takes it for per-rank inputs (sharded_input=True) with v_phase='columns'."""
    script = tmp_path / "repart_worker.py"
    script.write_text(_REPART_WORKER)
    port = str(35500 + os.getpid() % 2000 + world)
    names = "nt_logit_logit,nt_sg_csr_lin_logit,nt_sg_zero_ysample"
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), str(world), names],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "OK" in o, o[-3000:]


def test_column_phase_selection_and_errors():
    from pycmf_b200.cmf_solvers import MUSolver, NewtonSolver
    be = FakeBackend()
    assert NewtonSolver(x_link="logit", v_phase="columns", backend=be)._wants_columns(2)
    assert NewtonSolver(sg_sample_ratio=0.5, v_phase="columns", backend=be)._wants_columns(4)
    assert not NewtonSolver(x_link="logit", v_phase="columns", backend=be)._wants_columns(1)      # one rank: nothing to do
    assert not NewtonSolver(x_link="linear", v_phase="columns", backend=be)._wants_columns(2)     # shared Hessian
    assert not NewtonSolver(x_link="logit", v_phase="columns", update_V=False, backend=be)._wants_columns(2)
    assert NewtonSolver(x_link="logit", backend=be)._wants_columns(2)                             # 'auto': when it can
    assert not NewtonSolver(x_link="logit", v_phase="rows", backend=be)._wants_columns(2)
    with pytest.raises(ValueError, match="v_phase"):
        NewtonSolver(x_link="logit", v_phase="diagonal", backend=be)._wants_columns(2)

    class TwoRanks:
        rank, world = 1, 2
    A = np.arange(24.).reshape(4, 6)
    assert NewtonSolver(x_link="logit", backend=be, v_phase="rows", sharded_input=True)._prepare_column_block(
        be, TwoRanks(), A, be.ingest(A), 6, 4) == (None, None)
    blk, (c0, c1) = NewtonSolver(x_link="logit", backend=be)._prepare_column_block(
        be, TwoRanks(), be.ingest(A), be.ingest(A[2:]), 6, 2)                 # matrix already resident: a view of it
    assert (c0, c1) == (3, 6) and np.array_equal(blk.a, A[:, 3:6])
    s = NewtonSolver(x_link="logit", backend=be)
    blk, (c0, c1) = s._prepare_column_block(be, TwoRanks(), A, be.ingest(A[2:]), 6, 2)
    assert (c0, c1) == (3, 6) and np.array_equal(blk.a, A[:, 3:6])
    import scipy.sparse as sp
    blk, _ = s._prepare_column_block(be, TwoRanks(), sp.csr_matrix(A), be.ingest(A[2:]), 6, 2)
    assert blk.is_sparse and np.array_equal(blk.a.toarray(), A[:, 3:6])
    assert MUSolver(backend=be)._prepare_column_block(be, TwoRanks(), A, be.ingest(A), 6, 0) == (None, None)


def test_row_range_and_localize():
    from pycmf_b200.sharding import localize_indices, row_range
    for n, w in ((10, 3), (7, 8), (2000000, 8), (5, 1)):
        blocks = [row_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1
    out = localize_indices(np.array([[0, 5, 9], [4, 3, 7]]), 3, 8)
    assert out.tolist() == [[-1, 2, -1], [1, 0, 4]] and out.dtype == np.int32


def test_api_validation_errors_match_reference_messages():
    from pycmf_b200 import CMF, collective_matrix_factorization
    X, Y = np.ones((5, 2)), np.ones((5, 2))
    msg = "Expected X.shape[1] == Y.shape[0], found X.shape = {}, Y.shape = {}".format(X.shape, Y.shape)
    with pytest.raises(ValueError, match=re.escape(msg)):           # reference tests/test_cmf.py:28-34
        CMF(solver='mu', beta_loss=2).fit(X, Y)
    Y = np.ones((2, 3))
    with pytest.raises(ValueError, match="No such link"):
        collective_matrix_factorization(X, Y, n_components=2, x_link="probit")
    with pytest.raises(ValueError, match="No such solver"):
        collective_matrix_factorization(X, Y, n_components=2, solver="sgd")
    with pytest.raises(ValueError, match="Invalid init argument"):
        collective_matrix_factorization(X, Y, n_components=2, x_init="bogus")
    from sklearn.base import clone
    est = clone(CMF(n_components=3, solver="newton", dtype="float64"))
    assert est.get_params()["dtype"] == "float64" and est.get_params()["hessian_pertubation"] == 0.2


@pytest.mark.parametrize("init", [None, "random", "svd", "nndsvd", "nndsvda", "nndsvdar"])
def test_initialisation_matches_reference(init):
    from oracle.ref_loader import load_reference
    ref = load_reference()
    if ref is None:
        pytest.skip("reference not mounted")
    import warnings
    from pycmf.cmf import _initialize_mf as ref_init
    from pycmf_b200.init import _initialize_mf
    rng = np.random.RandomState(3)
    for shape, k in (((12, 9), 4), ((6, 5), 7)):
        Mx = np.abs(rng.randn(*shape))
        nonneg = init not in ("svd",)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                rA, rB = ref_init(Mx, k, init=init, random_state=0, non_negative=nonneg)
            except Exception as e:   # e.g. nndsvd with k > min(shape): same failure expected from ours
                with pytest.raises(type(e)):
                    _initialize_mf(Mx, k, init=init, random_state=0, non_negative=nonneg)
                continue
            A, B = _initialize_mf(Mx, k, init=init, random_state=0, non_negative=nonneg)
        assert np.allclose(A, rA) and np.allclose(B, rB)


def test_topic_terms_format(capsys):
    from pycmf_b200.analysis import _print_topic_terms_with_importances_from_matrices
    U = np.array([[0.1, 3.0], [2.0, 0.2], [1.0, 0.1]])
    Z = np.array([[0.5, 0.25]])
    _print_topic_terms_with_importances_from_matrices(U, Z, np.array(["a", "b", "c"]), topn_words=2)
    out = capsys.readouterr().out.strip().splitlines()
    assert out == ["Topic 1 [0.500]: c,b", "Topic 2 [0.250]: b,a"]


_RNG_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
from fake_backend import FakeBackend
from pycmf_b200.cmf_solvers import NewtonSolver, FitState
from pycmf_b200.sharding import TorchComm
rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=2)
np.random.seed(1000 + rank)                       # the ranks' global NumPy streams start DIFFERENT (random_state=None)
be, comm = FakeBackend(), TorchComm()
n, d, l, k = 12, 9, 4, 3
st = FitState(be, comm, None, None, torch.zeros(n // 2, k, dtype=torch.float64), torch.zeros(d, k, dtype=torch.float64),
              torch.zeros(l, k, dtype=torch.float64), n, (rank * n // 2, (rank + 1) * n // 2))
s = NewtonSolver(sg_sample_ratio=0.5, random_state=None, sampler="numpy", backend=be, comm=comm)
st.iteration = 1
m = s._masks(st)
# replicated masks (Z, Vy) must be identical on both ranks; Vx is localised to the rank's rows (-1 elsewhere)
for key in ("Z", "Vy"):
    mine = m[key].to(torch.float64).clone(); other = mine.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(mine, other), key
full = torch.where(m["Vx"] >= 0, m["Vx"] + st.r0, torch.zeros_like(m["Vx"])).to(torch.float64)
dist.all_reduce(full)
assert int((m["Vx"] >= 0).sum()) > 0 and float(full.max()) < n
dist.destroy_process_group()
print("OK")
'''


def test_numpy_sampler_is_synchronised_across_ranks_without_an_integer_seed(tmp_path):
    """with random_state=None every rank owns a different global NumPy stream; the solver re-seeds all ranks from
    rank 0 before the first draw, otherwise the replicated V / Z masks silently diverge."""
    script = tmp_path / "rng_worker.py"
    script.write_text(_RNG_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "OK" in o, o[-3000:]


def test_bench_config_and_algorithmic_work_follow_the_survey():
    """bench.py's `config` object is built by ONE function for both arms, and the step-level roofline uses SURVEY 8d's figures."""
    from pycmf_b200 import workloads as W
    c5 = W.describe("c5")
    fl, by = W.algorithmic_work(c5, 4)
    assert abs(fl - 1.036e13) / 1.036e13 < 5e-3 and abs(by - 8.12e10) / 8.12e10 < 5e-3
    c3 = W.describe("c3")
    fl3, by3 = W.algorithmic_work(c3, 4)
    assert abs(fl3 - 8.7e10) / 8.7e10 < 0.03 and abs(by3 - 5.06e9) / 5.06e9 < 0.03
    shard = W.describe("c3", 0.125)
    assert (shard["n"], shard["d"]) == (250000, 200000)            # a row shard keeps every column
    cfg = W.bench_config("c5", 1.0, 1.0, 8)
    assert cfg["workload"].startswith("c5: dense X 200000x50000") and "larger than L2" in cfg["l2_policy"]
    assert "L2-resident" in W.bench_config("c1", 1.0, 1.0, 1)["l2_policy"]
    assert W.bench_config("c2", 1.0, 1.0, 1)["deviation_from_SURVEY_8d"]


def test_n_components_above_the_backend_limit_is_rejected_before_any_device_work():
    from pycmf_b200.cmf_solvers import MUSolver
    s = MUSolver(max_iter=1)
    with pytest.raises(ValueError, match="n_components"):
        s.prepare(np.ones((4, 300)), np.ones((300, 2)), np.ones((4, 300)), np.ones((300, 300)), np.ones((2, 300)))


def _edge_names():
    from oracle.cases import EDGE
    return sorted(EDGE)


@pytest.mark.parametrize("name", _edge_names())
def test_solver_orchestration_on_degenerate_shapes(name):
    """One row, one label column, rank one, an all-zero CSR row, sample sets of one / zero indices, single-factor updates:
    the host orchestration (phases, chunked V rows, write-back of the updated factors only) against the oracle."""
    from helpers import run_oracle
    from oracle.cases import EDGE, make_edge_case
    params = EDGE[name][-1]
    case = make_edge_case(name)
    masks = draw_masks_for_case(case)
    hist_o, Uo, Vo, Zo = run_oracle(case, masks)
    hist, U, V, Z = _run_fake(case, masks)
    assert np.allclose(hist, hist_o, rtol=1e-9, atol=1e-11)
    for got, ref, upd in ((U, Uo, "update_U"), (V, Vo, "update_V"), (Z, Zo, "update_Z")):
        assert np.allclose(got, ref, rtol=1e-9, atol=1e-12)
        if not params.get(upd, True):
            assert np.array_equal(got, case[upd[-1] + "0"])            # a factor held fixed is never written back
