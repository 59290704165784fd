"""Drop-in check of the estimator shell on the CPU: `pycmf_b200.CMF` against the UNMODIFIED reference `pycmf.CMF`
(cmf.py:459-776) with the same arguments and random_state -- validation, initialisation (random / svd / nndsvd variants),
the (V + V_) / 2 merge, alpha = 'auto', the solver wiring, the early-stopping loop, `reconstruction_err_`, `n_iter_` and
`transform()` must agree.  The device backend is replaced by the float64 NumPy stand-in (tests/fake_backend.py), so this
tests the HOST side only; the kernels are compared with the oracle in the `-m gpu` tests.  Runs where /root/reference is
mounted."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

from oracle.ref_loader import load_reference

from helpers import rel_fro

pytestmark = pytest.mark.skipif(load_reference() is None, reason="/root/reference is not mounted here")


@pytest.fixture()
def fake_device(monkeypatch):
    import pycmf_b200.device as device
    from fake_backend import FakeBackend
    monkeypatch.setattr(device, "CudaBackend", lambda device=None, dtype=None, options=None: FakeBackend())


def _data(seed, n=36, d=24, l=5, k=4, sparse=False, unit=False):
    rng = np.random.RandomState(seed)
    Ut, Vt, Zt = np.abs(rng.randn(n, k)), np.abs(rng.randn(d, k)), rng.randn(l, k)
    X = Ut @ Vt.T + 0.05 * np.abs(rng.randn(n, d))
    if unit:
        X = 1.0 / (1.0 + np.exp(-(X - X.mean())))
    if sparse:
        X = sp.csr_matrix(X * (rng.rand(n, d) < 0.4))
    Y = 1.0 / (1.0 + np.exp(-(Vt @ Zt.T))) if unit else np.abs(Vt @ Zt.T)
    return X, Y


CONFIGS = {
    "mu_random": dict(solver="mu", x_init="random", y_init="random", max_iter=60),
    "mu_default_init": dict(solver="mu", max_iter=40),
    "mu_nndsvd": dict(solver="mu", x_init="nndsvd", y_init="nndsvda", max_iter=40),
    "mu_reg_sparse": dict(solver="mu", x_init="random", y_init="random", l1_reg=0.05, l2_reg=0.02, max_iter=40, sparse=True),
    "newton_lin_auto_alpha": dict(solver="newton", x_init="random", y_init="random", max_iter=25, l2_reg=0.1),
    "newton_logit_logit": dict(solver="newton", x_link="logit", y_link="logit", alpha=0.4, max_iter=15, l2_reg=0.1,
                               Z_non_negative=False, unit=True),
    "newton_svd_signed": dict(solver="newton", x_init="svd", y_init="svd", alpha=0.5, max_iter=15, l2_reg=0.2,
                              U_non_negative=False, V_non_negative=False, Z_non_negative=False),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_fit_transform_matches_reference(name, fake_device):
    load_reference()
    import pycmf as ref
    import pycmf_b200 as ours
    cfg = dict(CONFIGS[name])
    X, Y = _data(len(name), sparse=cfg.pop("sparse", False), unit=cfg.pop("unit", False))
    kw = dict(n_components=4, random_state=3, tol=1e-4, **cfg)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = ref.CMF(**kw)
        Ur, Vr, Zr = r.fit_transform(X, Y)
        o = ours.CMF(dtype="float64", **kw)
        Uo, Vo, Zo = o.fit_transform(X, Y)
    assert o.n_iter_ == r.n_iter_
    for got, want in ((Uo, Ur), (Vo, Vr), (Zo, Zr)):
        assert rel_fro(got, want) < 1e-8
    assert abs(o.reconstruction_err_ - r.reconstruction_err_) <= 1e-8 * abs(r.reconstruction_err_)
    assert o.n_components_ == r.n_components_ == 4
    # transform(): refit U and Z on new data with the components frozen (cmf.py:726-747)
    X2, Y2 = _data(len(name) + 100, sparse=sp.issparse(X), unit=CONFIGS[name].get("unit", False))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Ur2, Vr2, Zr2 = r.transform(X2, Y2)
        Uo2, Vo2, Zo2 = o.transform(X2, Y2)
    assert np.array_equal(Vo2, Vo) and rel_fro(Vo2, Vr2) < 1e-8          # V untouched
    assert rel_fro(Uo2, Ur2) < 1e-8 and rel_fro(Zo2, Zr2) < 1e-8


@pytest.mark.parametrize("solver", ["mu", "newton"])
@pytest.mark.parametrize("which", ["x_only", "y_only"])
def test_partial_transform_matches_reference(solver, which, fake_device):
    """transform(X, None) refits U only, transform(None, Y) refits Z only (cmf.py:726-747, update flags of the solvers
    cmf_solvers.py:252-261, :511-519)."""
    load_reference()
    import pycmf as ref
    import pycmf_b200 as ours
    X, Y = _data(5)
    kw = dict(n_components=4, random_state=1, solver=solver, x_init="random", y_init="random", max_iter=20, l2_reg=0.1)
    X2, Y2 = _data(6)
    args = (X2, None) if which == "x_only" else (None, Y2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = ref.CMF(**kw).fit(X, Y)
        o = ours.CMF(dtype="float64", **kw).fit(X, Y)
        try:
            want = r.transform(*args)
        except Exception as e:                      # the reference itself cannot run this combination
            with pytest.raises(type(e)):
                o.transform(*args)
            return
        got = o.transform(*args)
    for g, w in zip(got, want):
        assert rel_fro(g, w) < 1e-8
