"""GPU parity tests: the CUDA path (through the C ABI) against the reference-generated golden
fixtures and the CPU oracle.  Tolerances are north_star's: fp32 objective 1e-4 rel / factors 1e-3
rel Frobenius; fp64 1e-9 for both."""
import zlib

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import CASES, draw_masks_for_case, load_golden, rel_fro, run_oracle
from oracle import cmf_oracle as O

pytestmark = pytest.mark.gpu

TOL = {"float32": (1e-4, 1e-3), "float64": (1e-9, 1e-9)}


def run_ours(case, dtype, masks=None, **extra):
    from pycmf_b200.cmf_solvers import MUSolver, NewtonSolver
    p = dict(case["params"])
    solver = p.pop("solver")
    cls = MUSolver if solver == "mu" else NewtonSolver
    s = cls(max_iter=case["iters"], tol=0, random_state=case["rng_seed"], dtype=dtype, **p, **extra)
    s.history = []
    s.masks_per_iter = masks
    U, V, Z = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
    e0 = s.compute_error(case["X"], case["Y"], U, V, Z)
    U2, V2, Z2, n_iter = s.fit_iterative_update(case["X"], case["Y"], U, V, Z)
    assert U2 is U and V2 is V and Z2 is Z and n_iter == case["iters"]
    return np.asarray([e0] + s.history), U, V, Z


def assert_parity(hist, U, V, Z, ref_hist, rU, rV, rZ, dtype):
    otol, ftol = TOL[dtype]
    rel = np.abs(hist - ref_hist) / np.abs(ref_hist)
    assert rel.max() < otol, "objective rel err %.3e at iter %d" % (rel.max(), rel.argmax())
    for name, got, ref in (("U", U, rU), ("V", V, rV), ("Z", Z, rZ)):
        assert rel_fro(got, ref) < ftol, "%s rel fro %.3e" % (name, rel_fro(got, ref))


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_parity(name, dtype):
    case, g = load_golden(name)
    masks = draw_masks_for_case(case)
    hist, U, V, Z = run_ours(case, dtype, masks)
    assert_parity(hist, U, V, Z, g["objective"], g["U"], g["V"], g["Z"], dtype)


@pytest.mark.parametrize("name", ["nt_logit_logit", "nt_noreg_clamp", "nt_lin_logit"])
def test_jacobi_only_matches_cholesky_fastpath(name):
    case, g = load_golden(name)
    hist, U, V, Z = run_ours(case, "float64", backend_options={"chol_fastpath": 0})
    assert_parity(hist, U, V, Z, g["objective"], g["U"], g["V"], g["Z"], "float64")


def _mid_case(solver, n, d, l, k, sparse, seed, **params):
    rng = np.random.RandomState(seed)
    Ut, Vt, Zt = 0.6 * np.abs(rng.randn(n, k)), 0.6 * np.abs(rng.randn(d, k)), 0.6 * rng.randn(l, k)
    x_logit = params.get("x_link") == "logit"
    X = O.expit(Ut @ Vt.T - 1.0) if x_logit else Ut @ Vt.T + 0.05 * np.abs(rng.randn(n, d))
    if sparse:
        X = sp.csr_matrix(X * (rng.rand(n, d) < 0.05))
    Y = O.expit(Vt @ Zt.T) if params.get("y_link") == "logit" else np.abs(Vt @ Zt.T)
    sx, sy = np.sqrt(np.abs(X.mean()) / k), np.sqrt(np.abs(Y.mean()) / k)
    U0, V0 = sx * np.abs(rng.randn(n, k)), (sx * np.abs(rng.randn(d, k)) + sy * np.abs(rng.randn(d, k))) / 2
    Z0 = sy * rng.randn(l, k) if params.get("Z_non_negative") is False else sy * np.abs(rng.randn(l, k))
    return dict(X=X, Y=Y, U0=U0, V0=V0, Z0=Z0, params=dict(solver=solver, **params), iters=params.pop("iters", 8),
                rng_seed=seed)


# l1_reg = 0 wherever the non-negativity projection is on: l1 * sign(f) is discontinuous at the clamped zeros
# (sign(0) = 0 vs sign(1e-17) = 1), which makes the reference trajectory itself ill-conditioned (a 1e-12
# perturbation of U0 moves the objective by 2e-3 on nt_csr_logit_logit_k20; measured with the oracle).
NT = dict(alpha=0.4, l1_reg=0.0, l2_reg=0.1, Z_non_negative=False)
NT_SIGNED_L1 = dict(alpha=0.4, l1_reg=0.01, l2_reg=0.1, U_non_negative=False, V_non_negative=False,
                    Z_non_negative=False)
MID = {
    "mu_dense_k32": ("mu", 700, 300, 12, 32, False, dict(l1_reg=0.01, l2_reg=0.01)),
    "mu_csr_k64": ("mu", 900, 400, 6, 64, True, dict()),
    "mu_dense_k130": ("mu", 300, 280, 9, 130, False, dict()),
    "mu_dense_k256": ("mu", 520, 300, 5, 256, False, dict()),
    "nt_lin_logit_k32": ("newton", 600, 250, 10, 32, False, dict(NT, y_link="logit")),
    "nt_logit_logit_k16": ("newton", 300, 200, 6, 16, False, dict(NT, x_link="logit", y_link="logit")),
    "nt_csr_lin_lin_k24": ("newton", 500, 300, 6, 24, True, dict(NT)),
    "nt_csr_logit_logit_k20": ("newton", 260, 180, 5, 20, True, dict(NT, x_link="logit", y_link="logit",
                                                                        U_non_negative=False, V_non_negative=False)),
    "nt_lin_lin_k130": ("newton", 700, 600, 8, 130, False, dict(NT, l2_reg=1.0)),
    "nt_logit_lin_k72": ("newton", 150, 140, 4, 72, False, dict(NT, x_link="logit")),
    # d >= 1024 with few label rows: the grouped-row Hessian kernel of the Z update (two row groups, ragged k)
    "nt_lin_logit_d1100_k24": ("newton", 300, 1100, 20, 24, False, dict(NT_SIGNED_L1, y_link="logit", l1_reg=0.0)),
    "nt_signed_l1_lin_logit_k32": ("newton", 400, 200, 8, 32, False, dict(NT_SIGNED_L1, y_link="logit")),
    "nt_signed_l1_csr_logit_k16": ("newton", 300, 200, 6, 16, True, dict(NT_SIGNED_L1, x_link="logit", y_link="logit")),
}


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", sorted(MID))
def test_midsize_parity_vs_oracle(name, dtype):
    solver, n, d, l, k, sparse, params = MID[name]
    case = _mid_case(solver, n, d, l, k, sparse, seed=zlib.crc32(name.encode()) % 1000, **params)
    case["iters"] = 6
    ref_hist, rU, rV, rZ = run_oracle(case)
    hist, U, V, Z = run_ours(case, dtype)
    assert_parity(hist, U, V, Z, ref_hist, rU, rV, rZ, dtype)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_sampled_midsize_parity(dtype):
    case = _mid_case("newton", 120, 90, 7, 12, False, seed=5, **dict(NT, x_link="logit", y_link="logit",
                                                                      sg_sample_ratio=0.4))
    case["iters"] = 4
    masks = draw_masks_for_case(case)
    ref_hist, rU, rV, rZ = run_oracle(case, masks_per_iter=masks)
    hist, U, V, Z = run_ours(case, dtype, masks)
    assert_parity(hist, U, V, Z, ref_hist, rU, rV, rZ, dtype)


# ---- primitives ---------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def be64():
    from pycmf_b200.device import CudaBackend
    return CudaBackend(dtype="float64")


@pytest.mark.parametrize("shape", [(1, 1, 1), (70, 33, 129), (5, 300, 4000), (257, 64, 64)])
@pytest.mark.parametrize("trans", [False, True])
def test_gemm(be64, shape, trans):
    m, q, p = shape
    rng = np.random.RandomState(0)
    A = rng.randn(p, m) if trans else rng.randn(m, p)
    B = rng.randn(p, q)
    got = be64.to_host(be64.gemm(be64.to_device(A), be64.to_device(B), trans_a=trans))
    ref = (A.T if trans else A) @ B
    assert np.allclose(got, ref, rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("shape", [(1000, 32, 700), (300, 50, 4000), (129, 256, 65), (5000, 32, 20000), (64, 8, 64),
                                   (777, 10, 333), (2000, 130, 1500)])
@pytest.mark.parametrize("trans", [False, True])
def test_dmma_gemm_matches_numpy(be64, shape, trans):
    """float64 tensor-core GEMM (mma.sync.m8n8k4.f64, dmma.cu): ragged tiles, split-K, alpha / beta, both operand
    layouts -- against NumPy and against the FMA kernel it replaces."""
    from pycmf_b200.device import CudaBackend
    m, q, p = shape
    rng = np.random.RandomState(1)
    A = rng.randn(p, m) if trans else rng.randn(m, p)
    if A.shape[1] % 2:                                   # odd pitch: not eligible (16-byte chunks); pad the pitch instead
        A = np.ascontiguousarray(np.pad(A, ((0, 0), (0, 1))))[:, :-1]
    B, C0 = rng.randn(p, q + q % 2)[:, :q], rng.randn(m, q)
    ref = 0.7 * ((A.T if trans else A) @ B) - 1.3 * C0
    Ad = be64.to_device(np.ascontiguousarray(np.pad(A, ((0, 0), (0, A.shape[1] % 2)))))[:, :A.shape[1]]
    Bd = be64.to_device(np.ascontiguousarray(np.pad(B, ((0, 0), (0, q % 2)))))[:, :q]
    launches = be64.launch_count()
    be64.profile(True); be64.profile_reset()
    got = be64.to_host(be64.gemm(Ad, Bd, trans_a=trans, alpha=0.7, beta=-1.3, out=be64.to_device(C0)))
    _, cnt = be64.profile_query("dmma_gemm")
    be64.profile(False)
    assert cnt == 1, "the DMMA kernel did not run"
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max() * max(1, p // 100)
    fma = CudaBackend(dtype="float64", options={"dense_path": 0})
    got0 = fma.to_host(fma.gemm(fma.to_device(np.ascontiguousarray(A)), fma.to_device(np.ascontiguousarray(B)),
                                trans_a=trans, alpha=0.7, beta=-1.3, out=fma.to_device(C0)))
    assert np.abs(got - got0).max() <= 1e-12 * np.abs(ref).max() * max(1, p // 100)


@pytest.mark.parametrize("k", [1, 3, 10, 32, 64, 100, 256])
def test_spmm_and_transpose(be64, k):
    rng = np.random.RandomState(k)
    S = sp.random(200, 150, density=0.05, random_state=rng, format="csr")
    S[7, :] = 0
    S.eliminate_zeros()
    Sd = be64.ingest(S)
    B, A = rng.randn(150, k), rng.randn(200, k)
    assert np.allclose(be64.to_host(be64.spmm(Sd, be64.to_device(B))), S @ B, atol=1e-12)
    assert np.allclose(be64.to_host(be64.spmm(Sd, be64.to_device(A), transposed=True)), S.T @ A, atol=1e-12)


@pytest.mark.parametrize("k", [1, 2, 7, 32, 33, 64, 128, 150])
@pytest.mark.parametrize("chol,path", [(0, 2), (1, 2), (0, 1), (1, 1), (0, 0), (1, 0)])
def test_safe_solve_matches_eigh_clamp(k, chol, path):
    """x = S(H) g against eigh: Cholesky fast path, Frobenius shortcut, and the clamp-active solvers (path 1 = tridiagonalisation
    + bisection + inverse iteration, path 2 = the same with the Householder steps on a register-resident matrix for k = 64 / 128
    (the default), path 0 = one-sided Jacobi) on rank-deficient, indefinite, multiple-eigenvalue, clustered and clamp-level
    spectra."""
    from pycmf_b200.device import CudaBackend
    be = CudaBackend(dtype="float64", options={"chol_fastpath": chol, "solve_path": path})
    rng = np.random.RandomState(k)
    batch = 13
    H = np.empty((batch, k, k))
    for b in range(batch):
        r = max(1, (b * k) // batch)                      # ranks from deficient to full
        A = rng.randn(k, r)
        H[b] = A @ A.T * (0.05 if b % 3 == 0 else 1.0)
        if b == 4:
            H[b] -= 0.7 * np.eye(k)                       # indefinite: abs() of negative eigenvalues
        if b == 5:
            H[b] += 3.0 * np.eye(k)                       # safely positive definite: Cholesky path
        if b == 6:
            H[b] *= 0.15 / np.linalg.norm(H[b])           # ||H||_F < pert: every eigenvalue clamped, S(H) = I / pert
        if b == 7:
            H[b] *= 0.21 / np.linalg.norm(H[b])           # just above the Frobenius shortcut: the full path
        if b >= 9:
            Q, _ = np.linalg.qr(rng.randn(k, k))
            lam = [np.where(np.arange(k) % 2 == 0, 1.0, 0.05),                       # a k/2-fold eigenvalue above the clamp
                   0.2 + 1e-9 * rng.randn(k),                                        # everything at the clamp level
                   np.repeat(np.linspace(0.3, 5, (k + 1) // 2), 2)[:k] + 1e-8 * rng.randn(k),     # tight pairs
                   np.concatenate([np.linspace(0.25, 40, k - k // 3), 1e-4 * rng.rand(k // 3)])][b - 9]   # bulk below the clamp
            H[b] = (Q * lam) @ Q.T
            H[b] = 0.5 * (H[b] + H[b].T)
    g = rng.randn(batch, k)
    ref = np.einsum('bij,bj->bi', O.safe_invert(H, 0.2), g)
    got = be.to_host(be.safe_solve(be.to_device(H), be.to_device(g), 0.2))
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-10


@pytest.mark.parametrize("k", [48, 32, 64, 128, 256])
def test_spmm_hot_rows_and_columns_are_split(be64, k):
    """tf-idf-like skew: one row / one column holding thousands of nonzeros goes through the row-splitting path
    (k = 48: generic kernel; 32 .. 256: vector kernel, contiguous chunk ranges per warp)."""
    rng = np.random.RandomState(5)
    S = sp.random(3000, 2500, density=0.002, random_state=rng, format="lil")
    S[17, :] = rng.rand(2500)            # hot row: 2500 nonzeros = 5 chunks
    S[:, 33] = rng.rand(3000, 1)         # hot column: 3000 nonzeros in the CSC copy
    S = sp.csr_matrix(S)
    Sd = be64.ingest(S)
    B, A = rng.randn(2500, k), rng.randn(3000, k)
    C0 = rng.randn(3000, k)
    got = be64.to_host(be64.spmm(Sd, be64.to_device(B), alpha=0.5, beta=2.0, out=be64.to_device(C0)))
    assert np.allclose(got, 0.5 * (S @ B) + 2.0 * C0, atol=1e-10)
    assert np.allclose(be64.to_host(be64.spmm(Sd, be64.to_device(A), transposed=True)), S.T @ A, atol=1e-10)


def test_sqerr_dense_and_sparse(be64):
    rng = np.random.RandomState(3)
    A, B = rng.rand(90, 11), rng.rand(70, 11)
    T = rng.rand(90, 70) * (rng.rand(90, 70) < 0.2)
    for link in ("linear", "logit"):
        ref = O.compute_factorization_error(T, A, B.T, link) ** 2
        for tgt in (T, sp.csr_matrix(T)):
            got = float(be64.to_host(be64.sqerr(be64.to_device(A), be64.to_device(B), be64.ingest(tgt), link))[0])
            assert abs(got - ref) / ref < 1e-11


def test_device_sampler_draws_distinct_in_range(be64):
    idx = be64.to_host(be64.sample_indices(50, 1000, 300, seed=7, stream_id=3))
    assert idx.min() >= 0 and idx.max() < 1000
    assert all(len(set(r)) == 300 for r in idx)
    idx2 = be64.to_host(be64.sample_indices(50, 1000, 300, seed=7, stream_id=4))
    assert (idx != idx2).mean() > 0.9
    # roughly uniform marginals
    counts = np.bincount(idx.ravel(), minlength=1000)
    assert counts.max() < 40 and counts.min() >= 2


def test_cmf_api_end_to_end():
    """The reference's own usage (tests/test_cmf.py:56-64, :271-291): CMF.fit_transform, dense == sparse."""
    from pycmf_b200 import CMF
    rng = np.random.mtrand.RandomState(42)
    X = np.abs(rng.randn(10, 8))
    X[:, 2 * np.arange(4)] = 0
    Y = np.abs(rng.randn(8, 5))
    for solver in ("mu", "newton"):
        est1 = CMF(solver=solver, n_components=5, x_init='random', y_init='random', random_state=0, tol=1e-2,
                   dtype="float64")
        est2 = CMF(solver=solver, n_components=5, x_init='random', y_init='random', random_state=0, tol=1e-2,
                   dtype="float64")
        U1, V1, Z1 = est1.fit_transform(X, Y)
        U2, V2, Z2 = est2.fit_transform(sp.csr_matrix(X), Y)
        for a, b in ((U1, U2), (V1, V2), (Z1, Z2)):
            np.testing.assert_array_almost_equal(a, b, decimal=6)
        assert est1.reconstruction_err_ > 0 and est1.n_iter_ >= 1
        assert not ((U1 < 0).any() or (V1 < 0).any() or (Z1 < 0).any())


# ---- tcgen05 dense path (fp32, k == 32) ------------------------------------------------------------
def _tc_backend(path, splits=0):
    from pycmf_b200.device import CudaBackend
    opts = {"dense_path": path}
    if splits:
        opts["tc_max_splits"] = splits
    return CudaBackend(dtype="float32", options=opts)


@pytest.mark.parametrize("shape", [(1000, 520), (300, 2052), (4100, 260)])
@pytest.mark.parametrize("path,tol", [(0, 3e-6), (1, 2e-5), (2, 3e-3)])
@pytest.mark.parametrize("splits", [0, 1])
def test_tc_mu_products_match_numpy(shape, path, tol, splits):
    """X^T U and X V on the tensor cores (3xTF32 ~ fp32 accuracy, 1xTF32 ~ 1e-3) vs float64 NumPy."""
    n, d = shape
    rng = np.random.RandomState(n + d)
    X, U, V = rng.randn(n, d), rng.randn(n, 32), rng.randn(d, 32)
    be = _tc_backend(path, splits)
    Xd = be.ingest(X)
    buf = be.to_host(be.mu_v_partial(Xd, be.to_device(U)))
    if splits == 1 and path == 1:
        tol = 1e-4      # one long fp32 TMEM accumulation chain (production caps the chain at 32 tiles)
    assert rel_fro(buf[:d], X.T @ U) < tol
    assert rel_fro(buf[d:], U.T @ U) < 3e-6
    # X V through the MU left update: F <- F * (X V) / (F (V^T V)) with F = 1  =>  X V = F_new * (1 (V^T V))
    F = np.ones((n, 32))
    Fd = be.to_device(F)
    be.mu_left(Fd, be.to_device(V), Xd, 0.0, 0.0)
    got = be.to_host(Fd) * (F @ (V.T @ V))
    assert rel_fro(got, X @ V) < max(tol, 1e-5)


@pytest.mark.parametrize("link", ["linear", "logit"])
@pytest.mark.parametrize("path,tol", [(0, 5e-6), (1, 5e-5), (2, 3e-2)])
@pytest.mark.parametrize("splits", [0, 1])
def test_tc_fused_residual_right_matches_numpy(link, path, tol, splits):
    """gx = alpha (f(U V^T) - X)^T U from the fused tcgen05 kernel vs float64 NumPy (X never leaves fp32)."""
    n, d = 1100, 516
    rng = np.random.RandomState(7)
    U, V = 0.3 * rng.randn(n, 32), 0.3 * rng.randn(d, 32)
    X = (O.expit(U @ V.T) if link == "logit" else U @ V.T) + 0.05 * rng.randn(n, d)
    be = _tc_backend(path, splits)
    gx, Hx, per_row = be.newton_v_xpart(be.to_device(V), be.to_device(U), be.ingest(X), 0, d, link, 0.7)
    ref = 0.7 * (O.inverse(U @ V.T, link) - X.astype(np.float32).astype(np.float64)).T @ U
    assert rel_fro(be.to_host(gx), ref) < tol


@pytest.mark.parametrize("path", [1, 2])
def test_tc_long_tile_loop_newton_step(path):
    """Single split => every CTA walks all tiles (multi-phase mbarrier ring); one Newton step vs the oracle."""
    # signed factors: with the projection this shape bounces (objective 1.8e3 -> 6.2e3 -> 6.8e2) and amplifies any
    # rounding difference 5x per iteration, which tests the trajectory's conditioning rather than the kernel
    case = _mid_case("newton", 700, 1300, 9, 32, False, seed=11,
                     **dict(NT, y_link="logit", U_non_negative=False, V_non_negative=False))
    case["iters"] = 3
    ref_hist, rU, rV, rZ = run_oracle(case)
    hist, U, V, Z = run_ours(case, "float32", backend_options={"dense_path": path, "tc_max_splits": 1})
    otol, ftol = (1e-4, 1e-3) if path == 1 else (2e-3, 2e-2)
    assert (np.abs(hist - ref_hist) / np.abs(ref_hist)).max() < otol
    for got, ref in ((U, rU), (V, rV), (Z, rZ)):
        assert rel_fro(got, ref) < ftol


@pytest.mark.parametrize("ctas", [1, 3, 7, 0])
@pytest.mark.parametrize("chain", [0, 2, 5])
@pytest.mark.parametrize("link", ["linear", "logit"])
def test_tc_persistent_ranges_cross_own_tiles(ctas, chain, link):
    """The persistent tcgen05 pass: few CTAs walk many tiles, crossing own-tile and chain boundaries inside one CTA
    (P reloaded into tensor memory, partials per (CTA, own tile) summed by the reduce kernel).  Both directions
    and the fused sum of squares against float64 NumPy; ragged edges in both dimensions."""
    from pycmf_b200.device import CudaBackend
    n, d = 1100, 709
    rng = np.random.RandomState(ctas * 10 + chain)
    U, V = 0.3 * rng.randn(n, 32), 0.3 * rng.randn(d, 32)
    X = ((O.expit(U @ V.T) if link == "logit" else U @ V.T) + 0.05 * rng.randn(n, d)).astype(np.float32)
    Xp = np.zeros((n, 712), dtype=np.float32)          # row stride a multiple of 4 floats (TMA), d itself ragged
    Xp[:, :d] = X
    opts = {"dense_path": 1}
    if ctas:
        opts["tc_ctas"] = ctas
    if chain:
        opts["tc_chain"] = chain
    be = CudaBackend(dtype="float32", options=opts)
    Xd = be.ingest(Xp)
    from pycmf_b200.device import DenseMatrix
    Xv = DenseMatrix(Xd.t[:, :d])
    outL, outR, sq = be.resid_pass(be.to_device(U), be.to_device(V), Xv, link, want_sq=True)
    R = O.inverse(U.astype(np.float32).astype(np.float64) @ V.astype(np.float32).astype(np.float64).T, link) - X
    assert rel_fro(be.to_host(outL), R @ V) < 5e-5
    assert rel_fro(be.to_host(outR), R.T @ U) < 5e-5
    assert abs(float(be.to_host(sq)[0]) - (R ** 2).sum()) / (R ** 2).sum() < 1e-5


# ---- tcgen05 MU contractions for n_components = 64 .. 256 (tc_mu.cu) ----------------------------------
@pytest.mark.parametrize("k", [64, 128, 192, 256])
@pytest.mark.parametrize("shape", [(1000, 520), (300, 2052), (4100, 260)])
@pytest.mark.parametrize("signed", [True, False])
def test_tc_mu_wide_products_match_numpy(k, shape, signed):
    """X^T U and X V for wide factors on the tensor cores (3xTF32) vs float64 NumPy; ragged edges in both dimensions.
    Non-negative data (the MU case) has no cancellation: tighter bound."""
    n, d = shape
    rng = np.random.RandomState(n + d + k)
    X, U, V = rng.randn(n, d), rng.randn(n, k), rng.randn(d, k)
    if not signed:
        X, U, V = np.abs(X), np.abs(U), np.abs(V)
    tol = 4e-5 if signed else 1.5e-5      # tensor-memory accumulation truncates (chains of 16 tiles)
    be = _tc_backend(1)
    Xd = be.ingest(X)
    buf = be.to_host(be.mu_v_partial(Xd, be.to_device(U)))
    assert rel_fro(buf[:d], X.T @ U) < tol
    F = np.ones((n, k))
    Fd = be.to_device(F)
    be.mu_left(Fd, be.to_device(V), Xd, 0.0, 0.0)
    got = be.to_host(Fd) * (F @ (V.T @ V))
    assert rel_fro(got, X @ V) < max(tol, 1e-5)


@pytest.mark.parametrize("ctas", [1, 3, 0])
@pytest.mark.parametrize("chain", [0, 2, 5])
def test_tc_mu_wide_persistent_ranges(ctas, chain):
    """Few CTAs walk many tiles (own-tile and chain boundaries inside one CTA, several partials per own tile);
    the same products from the generic FMA kernels must agree to fp32 accuracy."""
    from pycmf_b200.device import CudaBackend, DenseMatrix
    n, d, k = 1100, 709, 128
    rng = np.random.RandomState(ctas * 10 + chain)
    U, V = np.abs(rng.randn(n, k)), np.abs(rng.randn(d, k))
    Xp = np.zeros((n, 712), dtype=np.float32)
    Xp[:, :d] = np.abs(rng.randn(n, d))
    X = Xp[:, :d].astype(np.float64)
    opts = {"dense_path": 1}
    if ctas:
        opts["tc_ctas"] = ctas
    if chain:
        opts["tc_chain"] = chain
    be = CudaBackend(dtype="float32", options=opts)
    Xv = DenseMatrix(be.ingest(Xp).t[:, :d])
    buf = be.to_host(be.mu_v_partial(Xv, be.to_device(U)))
    assert rel_fro(buf[:d], X.T @ U) < 1.5e-5
    F = np.ones((n, k))
    Fd = be.to_device(F)
    be.mu_left(Fd, be.to_device(V), Xv, 0.0, 0.0)
    assert rel_fro(be.to_host(Fd) * (F @ (V.T @ V)), X @ V) < 1.5e-5


def test_tc_mu_wide_fit_matches_oracle():
    """30 MU iterations with k = 64 on a mid-size dense problem: tensor-core path vs the float64 oracle."""
    from pycmf_b200.cmf_solvers import MUSolver
    rng = np.random.RandomState(5)
    n, d, l, k = 1500, 900, 40, 64
    X = np.abs(rng.randn(n, 24) @ rng.randn(24, d)) + 0.1 * np.abs(rng.randn(n, d))
    Y = np.abs(rng.randn(d, l))
    U0, V0, Z0 = np.abs(rng.randn(n, k)) * 0.3, np.abs(rng.randn(d, k)) * 0.3, np.abs(rng.randn(l, k)) * 0.3
    Uo, Vo, Zo = U0.copy(), V0.copy(), Z0.copy()
    for _ in range(30):
        O.mu_step(X, Y, Uo, Vo, Zo, 0.0, 0.0)
    U, V, Z = U0.copy(), V0.copy(), Z0.copy()
    s = MUSolver(tol=0, max_iter=30, dtype="float32", backend_options={"dense_path": 1}, use_cuda_graph=True)
    s.fit_iterative_update(X, Y, U, V, Z)
    assert rel_fro(U, Uo) < 1e-3 and rel_fro(V, Vo) < 1e-3 and rel_fro(Z, Zo) < 1e-3


@pytest.mark.parametrize("k", [32, 64, 128])
@pytest.mark.parametrize("unroll,lean", [(4, 1), (8, 1), (4, 0), (8, 0)])
@pytest.mark.parametrize("shape", [(3000, 2500, 0.01), (40, 70, 0.2), (5000, 300, 0.002), (257, 4000, 0.05)])
def test_spmm_nonzero_balanced_kernel(k, unroll, lean, shape):
    """The nonzero-balanced fp32 SpMM: shares that cut rows (atomics), whole rows (stores), empty rows (prologue), more
    warps than nonzeros, hot rows / columns, alpha / beta -- against scipy in float64."""
    from pycmf_b200.device import CudaBackend
    n, d, dens = shape
    rng = np.random.RandomState(5)
    S = sp.random(n, d, density=dens, random_state=rng, format="lil")
    S[min(17, n - 1), :] = rng.rand(d)
    S[:, min(33, d - 1)] = rng.rand(n, 1)
    S[5, :] = 0
    S[n - 1, :] = 0                     # trailing empty row
    S = sp.csr_matrix(S)
    S.eliminate_zeros()
    be = CudaBackend(dtype="float32", options={"spmm_path": 1, "spmm_unroll": unroll, "spmm_lean": lean})
    Sd = be.ingest(S)
    B, A, C0 = rng.randn(d, k), rng.randn(n, k), rng.randn(n, k)
    got = be.to_host(be.spmm(Sd, be.to_device(B), alpha=0.5, beta=2.0, out=be.to_device(C0)))
    assert rel_fro(got, 0.5 * (S @ B) + 2.0 * C0) < 2e-6
    got0 = be.to_host(be.spmm(Sd, be.to_device(B), out=be.to_device(np.full((n, k), np.nan))))   # beta = 0 ignores C
    assert rel_fro(got0, S @ B) < 2e-6
    assert rel_fro(be.to_host(be.spmm(Sd, be.to_device(A), transposed=True)), S.T @ A) < 2e-6


@pytest.mark.parametrize("k", [32, 64, 128])
@pytest.mark.parametrize("path", [0, 1, 2, 3])
def test_spmm_float32_kernels_agree(k, path):
    """The fp32 SpMM kernels (generic, nonzero-balanced, vector, sub-warp grouped) on a skewed matrix with hot rows /
    columns, alpha / beta handling included."""
    from pycmf_b200.device import CudaBackend
    rng = np.random.RandomState(11)
    S = sp.random(3000, 2500, density=0.01, random_state=rng, format="lil")
    S[17, :] = rng.rand(2500)
    S[:, 33] = rng.rand(3000, 1)
    S[5, :] = 0
    S = sp.csr_matrix(S)
    S.eliminate_zeros()
    be = CudaBackend(dtype="float32", options={"spmm_path": path})
    Sd = be.ingest(S)
    B, A, C0 = rng.randn(2500, k), rng.randn(3000, k), rng.randn(3000, k)
    got = be.to_host(be.spmm(Sd, be.to_device(B), alpha=0.5, beta=2.0, out=be.to_device(C0)))
    assert rel_fro(got, 0.5 * (S @ B) + 2.0 * C0) < 2e-6
    assert rel_fro(be.to_host(be.spmm(Sd, be.to_device(A), transposed=True)), S.T @ A) < 2e-6


def test_ingest_builds_the_csc_copy_on_device(be64):
    """Only the CSR arrays are uploaded; the CSC copy built on the device equals scipy's sorted transposition bit for bit
    (same order inside a column => same summation order as the host-built copy), empty rows / columns included."""
    rng = np.random.RandomState(3)
    S = sp.random(700, 430, density=0.03, random_state=rng, format="lil")
    S[11, :] = 0
    S[:, 7] = 0
    S[:, 20] = rng.rand(700, 1)
    S = sp.csr_matrix(S)
    S.eliminate_zeros()
    Sd = be64.ingest(S)
    ref = S.T.tocsr()
    ref.sort_indices()
    assert np.array_equal(be64.to_host(Sd.colptr), ref.indptr)
    assert np.array_equal(be64.to_host(Sd.rowidx), ref.indices)
    assert np.array_equal(be64.to_host(Sd.cvals), ref.data)
    sub = be64.row_slice(Sd, 100, 613)
    ref = S[100:613].T.tocsr()
    ref.sort_indices()
    assert np.array_equal(be64.to_host(sub.colptr), ref.indptr)
    assert np.array_equal(be64.to_host(sub.rowidx), ref.indices)
    assert np.array_equal(be64.to_host(sub.cvals), ref.data)


def test_dense_ingest_pads_the_row_pitch_to_128_bytes():
    """Unaligned row pitch (here 1300 x 4 B) is padded at ingest; every kernel family takes the leading dimension, so
    products and the objective must not change (compared with the unpadded layout, and against NumPy)."""
    from pycmf_b200.device import CudaBackend
    rng = np.random.RandomState(2)
    n, d, k = 700, 1300, 32
    X, U, V = rng.rand(n, d), rng.rand(n, k), rng.rand(d, k)
    outs = []
    for pad in (1, 0):
        be = CudaBackend(dtype="float32", options={"pad_pitch": pad})
        Xd = be.ingest(X)
        assert (Xd.t.stride(0) * 4) % 128 == (0 if pad else 1300 * 4 % 128) and Xd.shape == (n, d)
        buf = be.to_host(be.mu_v_partial(Xd, be.to_device(U)))
        sq = float(be.to_host(be.sqerr(be.to_device(U), be.to_device(V), Xd, "linear"))[0])
        outL, outR, _ = be.resid_pass(be.to_device(U), be.to_device(V), Xd, "linear")
        outs.append((buf, sq, be.to_host(outL), be.to_host(outR)))
    R = U @ V.T - X
    assert rel_fro(outs[0][0][:d], X.T @ U) < 2e-5 and rel_fro(outs[0][2], R @ V) < 5e-5 and rel_fro(outs[0][3], R.T @ U) < 5e-5
    assert abs(outs[0][1] - (R ** 2).sum()) / (R ** 2).sum() < 1e-5
    for a, b in zip(outs[0], outs[1]):
        assert rel_fro(np.asarray(a), np.asarray(b)) < 1e-5


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 2e-6)])
@pytest.mark.parametrize("shape", [(1000, 64), (77, 10), (300, 128), (5000, 33), (131, 20), (257, 32)])
def test_mu_fused_update_matches_the_unfused_path(dtype, tol, shape):
    """F <- F * N / (F G + l1 + l2 F) with the denominator product inside the kernel (mu_fused_kernel) against the
    separate GEMM + elementwise launches, through the MU left update (zero denominators included)."""
    from pycmf_b200.device import CudaBackend
    rows, k = shape
    rng = np.random.RandomState(12)
    F0, B, T = np.abs(rng.randn(rows, k)), np.abs(rng.randn(90, k)), np.abs(rng.randn(rows, 90))
    F0[3] = 0.0                                             # a zero row: denominator 0 -> float32 eps (cmf_solvers.py:219)
    outs = []
    for fused in (1, 0):
        be = CudaBackend(dtype=dtype, options={"mu_fused": fused, "dense_path": 0})
        F = be.to_device(F0)
        be.profile(True); be.profile_reset()
        be.mu_left(F, be.to_device(B), be.ingest(T), 0.01, 0.02)
        assert be.profile_query("mu_fused")[1] == fused
        outs.append(be.to_host(F))
    ref = F0 * ((T @ B) / np.where((d := F0 @ (B.T @ B) + 0.01 + 0.02 * F0) == 0, np.finfo(np.float32).eps, d))
    assert rel_fro(outs[0], outs[1]) < tol and rel_fro(outs[0], ref) < max(tol, 1e-6 if dtype == "float32" else 1e-12)


@pytest.mark.parametrize("link", ["linear", "logit"])
@pytest.mark.parametrize("shape", [(300, 200, 32), (1000, 130, 17), (64, 64, 128), (777, 333, 64), (129, 500, 10)])
@pytest.mark.parametrize("trans", [False, True])
def test_dmma_fused_residual_matches_numpy(be64, link, shape, trans):
    """float64 fused residual pass on the DMMA pipe (dmma_resid_kernel): R = f(A B^T) - T never leaves the SM;
    R B, R^T A and sum R^2 against NumPy, ragged tiles, odd k, transposed target, both links."""
    ra, rb, k = shape
    rng = np.random.RandomState(3)
    A, B = 0.3 * rng.randn(ra, k), 0.3 * rng.randn(rb, k)
    Tm = rng.rand(ra, rb)
    est = A @ B.T
    R = (O.expit(est) if link == "logit" else est) - Tm
    from pycmf_b200.device import DenseMatrix
    Td = DenseMatrix(be64.to_device(np.ascontiguousarray(Tm.T if trans else Tm)))
    be64.profile(True); be64.profile_reset()
    if trans:
        # target stored transposed (rows(B) x rows(A)), as the Z update reads Y (cmf_solvers.py:497)
        outL, outR, sq = be64.resid_pass(be64.to_device(A), be64.to_device(B), Td, link, want_sq=True, trans_t=True)
        wantL, wantR = R @ B, R.T @ A
    else:
        outL, outR, sq = be64.resid_pass(be64.to_device(A), be64.to_device(B), Td, link, want_sq=True)
        wantL, wantR = R @ B, R.T @ A
    ran = be64.profile_query("dmma_resid_left")[1] + be64.profile_query("dmma_resid_right")[1]
    be64.profile(False)
    assert ran == 2, "the DMMA residual kernels did not run"
    assert rel_fro(be64.to_host(outL), wantL) < 1e-12 and rel_fro(be64.to_host(outR), wantR) < 1e-12
    assert abs(float(be64.to_host(sq)[0]) - (R * R).sum()) <= 1e-12 * (R * R).sum()


@pytest.mark.parametrize("k", [64, 128])
@pytest.mark.parametrize("link", ["logit", "linear"])
@pytest.mark.parametrize("sparse", [False, True])
def test_hessian_mma_matches_fma_and_oracle(k, link, sparse):
    """Per-row weighted Grams on the tensor cores (row_grad_hess_mma_kernel, mma.sync 3xTF32) against the FMA kernel and
    the float64 oracle: one sampled Newton update of the left factor (gradient + Hessian + clamped solve per row)."""
    from pycmf_b200.device import CudaBackend
    rng = np.random.RandomState(17)
    rows, m, ns = 150, 700, 230
    F0, B = 0.2 * rng.randn(rows, k), 0.3 * rng.randn(m, k)
    T = O.expit(rng.randn(rows, m)) if link == "logit" else rng.randn(rows, m)
    if sparse:
        T = sp.csr_matrix(T * (rng.rand(rows, m) < 0.1))
    idx = np.stack([rng.permutation(m)[:ns] for _ in range(rows)]).astype(np.int32)
    ref = F0.copy()
    O._rows_newton(ref, B, T, 0.6, 0.0, 0.1, link, False, 0.2, l2_in_logit_hessian=True, idx=idx)
    outs = []
    for mma in (1, 0):
        be = CudaBackend(dtype="float32", options={"hess_mma": mma})
        F = be.to_device(F0)
        be.newton_left(F, be.to_device(B), be.ingest(T), 0.6, 0.0, 0.1, link, False, 0.2, True, idx=be.to_device(idx, np.int32))
        outs.append(be.to_host(F))
    assert rel_fro(outs[0], outs[1]) < 2e-5
    assert rel_fro(outs[0], ref) < 1e-4 and rel_fro(outs[1], ref) < 1e-4
