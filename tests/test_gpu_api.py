"""The estimator surface on the GPU: `CMF.transform()` full and partial (reference
cmf.py:726-747), `reconstruction_err_`, the up-front n_components limit and the device top-k behind
`print_topic_terms` (reference analysis.py:1-16) -- against the unmodified reference (baseline/_ref) or NumPy."""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest
import scipy.sparse as sp

from oracle.ref_loader import load_reference
from helpers import rel_fro

pytestmark = pytest.mark.gpu


def _data(seed, n=300, d=120, l=8, k=6, sparse=False):
    rng = np.random.RandomState(seed)
    Ut, Vt, Zt = np.abs(rng.randn(n, k)), np.abs(rng.randn(d, k)), np.abs(rng.randn(l, k))
    X = Ut @ Vt.T + 0.05 * np.abs(rng.randn(n, d))
    Y = Vt @ Zt.T + 0.05 * np.abs(rng.randn(d, l))
    if sparse:
        X = sp.csr_matrix(X * (rng.rand(n, d) < 0.2))
    return X, Y


@pytest.mark.parametrize("solver,kw", [("mu", {}), ("newton", dict(alpha=0.5, l2_reg=0.1, U_non_negative=False,
                                                                  V_non_negative=False, Z_non_negative=False))])
@pytest.mark.parametrize("sparse", [False, True])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-8), ("float32", 1e-3)])
def test_transform_full_and_partial_match_the_reference(solver, kw, sparse, dtype, tol):
    ref = load_reference()
    if ref is None:
        pytest.skip("unmodified reference not available (baseline/_ref)")
    from pycmf_b200 import CMF
    X, Y = _data(0, sparse=sparse)
    Xn, Yn = _data(1, sparse=sparse)
    common = dict(n_components=6, solver=solver, max_iter=15, tol=0, random_state=3, x_init="random", y_init="random", **kw)
    r = ref.CMF(**common)
    Ur, Vr, Zr = r.fit_transform(X, Y)
    m = CMF(dtype=dtype, **common)
    U, V, Z = m.fit_transform(X, Y)
    assert max(rel_fro(U, Ur), rel_fro(V, Vr), rel_fro(Z, Zr)) < tol
    assert abs(m.reconstruction_err_ - r.reconstruction_err_) <= 10 * tol * r.reconstruction_err_
    # transform from the SAME fitted state on both sides
    m.components, m.x_weights, m.y_weights = Vr.copy(), Ur.copy(), Zr.copy()
    for Xa, Ya in ((Xn, Yn), (Xn, None), (None, Yn)):
        r.components, r.x_weights, r.y_weights = Vr.copy(), Ur.copy(), Zr.copy()
        m.components, m.x_weights, m.y_weights = Vr.copy(), Ur.copy(), Zr.copy()
        Ua, Va, Za = r.transform(Xa, Ya)
        Ub, Vb, Zb = m.transform(Xa, Ya)
        assert np.array_equal(Vb, Vr)                                 # components stay fixed, bit for bit (tests/test_cmf.py:408)
        assert max(rel_fro(Ub, Ua), rel_fro(Zb, Za)) < tol


def test_n_components_above_the_backend_limit_fails_up_front():
    from pycmf_b200 import CMF
    X, Y = _data(2, n=40, d=300, l=4)
    with pytest.raises(ValueError, match="n_components"):
        CMF(max_iter=2).fit_transform(X, Y)                         # n_components=None -> max(300, 4) > 256


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_device_topk_matches_argsort(dtype):
    from pycmf_b200.device import CudaBackend
    rng = np.random.RandomState(4)
    F = rng.rand(5000, 37).astype(dtype)
    be = CudaBackend(dtype=dtype)
    for topn in (1, 10, 64):
        got = be.to_host(be.topk_per_column(be.to_device(F), topn))
        want = np.argsort(F, axis=0, kind="stable")[-topn:].T      # (columns x topn), ascending weight like analysis.py:6
        assert np.array_equal(np.take_along_axis(F.T, got, 1), np.take_along_axis(F.T, want, 1))
        assert np.array_equal(got, want)


def test_print_topic_terms_uses_the_device_topk_and_matches_the_reference_format():
    ref = load_reference()
    from pycmf_b200 import CMF, analysis
    rng = np.random.RandomState(5)
    U, Z = rng.rand(4000, 5), rng.rand(3, 5)
    words = np.array(["w%d" % i for i in range(4000)])
    buf = io.StringIO()
    with redirect_stdout(buf):
        analysis._print_topic_terms_with_importances_from_matrices(U, Z, words, topn_words=10, device=0)
    ours = buf.getvalue()
    if ref is not None:
        buf = io.StringIO()
        with redirect_stdout(buf):
            ref.analysis._print_topic_terms_with_importances_from_matrices(U, Z, words)
        assert ours == buf.getvalue()
    assert ours.count("Topic") == 5


@pytest.mark.parametrize("init", ["random", "svd", "nndsvd", "nndsvda", "nndsvdar"])
@pytest.mark.parametrize("sparse", [False, True])
def test_device_initialisation_matches_the_host_path(init, sparse):
    """`_initialize_mf` on the GPU (init_device.py: the library's GEMM / SpMM for every product with M) against the host
    restatement of reference cmf.py:41-202 (init.py, itself checked against the live reference on the CPU): same NumPy
    random draws on both sides, so the factors agree to rounding."""
    from pycmf_b200.device import CudaBackend
    from pycmf_b200.init import _initialize_mf
    from pycmf_b200.init_device import initialize_mf_device
    rng = np.random.RandomState(8)
    n, d, k = 500, 180, 7
    M = np.abs(rng.randn(n, 12)) @ np.abs(rng.randn(12, d)) * (1.0 + np.arange(d) / d)[None, :] + 0.01 * np.abs(rng.randn(n, d))
    if sparse:
        M = sp.csr_matrix(M * (rng.rand(n, d) < 0.3))
    nn = init != "svd"
    A, B = _initialize_mf(M, k, init=init, random_state=11, non_negative=nn)
    be = CudaBackend(dtype="float64")
    Ad, Bd = initialize_mf_device(be, be.ingest(M), k, init=init, random_state=11, non_negative=nn)
    assert rel_fro(be.to_host(Ad), A) < 1e-7 and rel_fro(be.to_host(Bd), B) < 1e-7
    for rows, cols in ((60, 300),):                       # rows < cols: the transposed branch of the range finder
        M2 = np.abs(rng.randn(rows, 9)) @ np.abs(rng.randn(9, cols)) + 0.01 * np.abs(rng.randn(rows, cols))
        A2, B2 = _initialize_mf(M2, 5, init=init, random_state=2, non_negative=nn)
        A2d, B2d = initialize_mf_device(be, be.ingest(M2), 5, init=init, random_state=2, non_negative=nn)
        assert rel_fro(be.to_host(A2d), A2) < 1e-7 and rel_fro(be.to_host(B2d), B2) < 1e-7


def test_fit_with_device_initialisation_matches_host_initialisation():
    from pycmf_b200 import CMF
    X, Y = _data(6, n=400, d=150, l=6, k=5)
    kw = dict(n_components=5, solver="mu", max_iter=20, tol=0, random_state=1, dtype="float64")
    Uh, Vh, Zh = CMF(**kw).fit_transform(X, Y)
    Ud, Vd, Zd = CMF(init_on_device=True, **kw).fit_transform(X, Y)
    assert max(rel_fro(Ud, Uh), rel_fro(Vd, Vh), rel_fro(Zd, Zh)) < 1e-6
