"""scikit-learn style API of PyCMF on the B200 backend.

Mirrors reference pycmf/cmf.py: `collective_matrix_factorization` (:215-456) and `CMF` (:459-776)
keep their signatures, defaults, attribute names, warnings and error messages; the solver objects
they construct are the device solvers of pycmf_b200.cmf_solvers.  Extra, optional keyword arguments
select backend behaviour only: `dtype`, `device`, `sampler`, `comm`, `backend_options`.
"""
import warnings

import numpy as np
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.utils import check_array

from .analysis import _print_topic_terms_from_matrix, _print_topic_terms_with_importances_from_matrices
from .cmf_solvers import MUSolver, NewtonSolver
from .init import _init_custom


def _validated(M):
    return check_array(M, accept_sparse=('csr', 'csc'), dtype=float)


def compute_factorization_error(target, left_factor, right_factor, link, beta_loss="frobenius",
                                dtype="float64", device=None):
    """||target - f(left_factor @ right_factor)||_F evaluated on the GPU (reference cmf_solvers.py:36-42).
    `right_factor` is k x cols, as in the reference's call sites (cmf.py:697-698)."""
    if target is None:
        return 0
    from .device import CudaBackend
    be = CudaBackend(device=device, dtype=dtype)
    try:
        T = be.ingest(target)
        A = be.to_device(np.asarray(left_factor))
        B = be.to_device(np.ascontiguousarray(np.asarray(right_factor).T))
        return float(np.sqrt(max(float(be.to_host(be.sqerr(A, B, T, link))[0]), 0.0)))
    finally:
        be.close()


def collective_matrix_factorization(X, Y, U=None, V=None, Z=None,
                                    n_components=None, solver="mu", alpha=0.5,
                                    x_init=None, y_init=None, beta_loss="frobenius",
                                    tol=1e-4, l1_reg=0., l2_reg=0.,
                                    random_state=None, max_iter=200, verbose=0,
                                    U_non_negative=True, V_non_negative=True,
                                    Z_non_negative=True, update_U=True,
                                    update_V=True, update_Z=True,
                                    x_link="linear", y_link="linear",
                                    hessian_pertubation=0.2, sg_sample_ratio=1.,
                                    dtype="float32", device=None, sampler="auto", comm=None,
                                    backend_options=None, init_on_device=False, return_errors=False):
    """Compute Collective Matrix Factorization: X ~= f1(U V^T), Y ~= f2(V Z^T).

    Same contract as the reference function (cmf.py:215-456): returns (U, V, Z, n_iter); custom
    initial factors are updated in place and returned by identity.
    """
    if n_components is None:
        n_components = max(X.shape[1], Y.shape[1])
    if update_U or update_V:
        X = _validated(X)
    if update_Z or update_V:
        Y = _validated(Y)
    if update_V and X.shape[1] != Y.shape[0]:
        raise ValueError("Expected X.shape[1] == Y.shape[0], " +
                         "found X.shape = {}, Y.shape = {}".format(X.shape[1], Y.shape[0]))
    if x_link not in ["linear", "logit"]:
        raise ValueError("No such link %s for x_link" % x_link)
    if y_link not in ["linear", "logit"]:
        raise ValueError("No such link %s for y_link" % y_link)

    # ---- initial factors (cmf.py:401-430)
    backend = None
    if init_on_device and x_init != 'custom' and y_init != 'custom':
        # the same algorithms on the ingested, device-resident matrices (init_device.py); the fit then reuses them
        from .device import CudaBackend
        from .init_device import initialize_mf_device
        backend = CudaBackend(device=device, dtype=dtype, options=backend_options)
        X = backend.ingest(X) if X is not None else None
        Y = backend.ingest(Y.toarray() if hasattr(Y, "toarray") else Y) if Y is not None else None

        def _initialize_mf(M, k, init=None, random_state=None, non_negative=False):       # noqa: F811 (shadows the host one)
            A, B = initialize_mf_device(backend, M, k, init=init, random_state=random_state, non_negative=non_negative)
            return backend.to_host(A).astype(np.float64), backend.to_host(B).astype(np.float64)
    else:
        from .init import _initialize_mf
    if x_init == 'custom':
        if X is not None:
            U = _init_custom(U, X, n_components, 0, non_negative=U_non_negative, random_state=random_state)
            V = _init_custom(V, X, n_components, 1, non_negative=V_non_negative, random_state=random_state)
    else:
        x_init = "random" if x_link == "logit" else x_init
        U, V = _initialize_mf(X, n_components, init=x_init, random_state=random_state,
                              non_negative=(U_non_negative or V_non_negative))
    if y_init == 'custom':
        if Y is not None:
            V = _init_custom(V, Y, n_components, 0, non_negative=V_non_negative, random_state=random_state)
            Z = _init_custom(Z, Y, n_components, 1, non_negative=Z_non_negative, random_state=random_state)
        V_ = V
    else:
        y_init = "random" if y_link == "logit" else y_init
        V_, Z = _initialize_mf(Y, n_components, init=y_init, random_state=random_state,
                               non_negative=(Z_non_negative or V_non_negative))
    if U_non_negative == Z_non_negative:
        V = (V + V_) / 2
    elif Z_non_negative:
        V = V_

    backend_kw = dict(dtype=dtype, device=device, sampler=sampler, comm=comm, backend_options=backend_options,
                      backend=backend)
    if solver == "mu":
        if x_link != "linear" or y_link != "linear":
            warnings.warn("mu solver does not accept link functions other than linear, link arguments will be ignored")
        solver_object = MUSolver(max_iter=max_iter, tol=tol, verbose=verbose,
                                 update_U=update_U, update_V=update_V, update_Z=update_Z,
                                 l1_reg=l1_reg, l2_reg=l2_reg, beta_loss=beta_loss, random_state=random_state,
                                 **backend_kw)
    elif solver == "newton":
        if alpha == "auto":
            # weigh X and Y equally: X.shape[0] * alpha == Y.shape[1] * (1 - alpha)   (cmf.py:441-444)
            alpha = Y.shape[1] / (X.shape[0] + Y.shape[1])
        solver_object = NewtonSolver(alpha=alpha, l1_reg=l1_reg, tol=tol,
                                     l2_reg=l2_reg, max_iter=max_iter, verbose=verbose,
                                     update_U=update_U, update_V=update_V, update_Z=update_Z,
                                     U_non_negative=U_non_negative, V_non_negative=V_non_negative,
                                     Z_non_negative=Z_non_negative, x_link=x_link, y_link=y_link,
                                     hessian_pertubation=hessian_pertubation,
                                     sg_sample_ratio=sg_sample_ratio, random_state=random_state,
                                     **backend_kw)
    else:
        raise ValueError("No such solver: %s" % solver)
    # factors handed to the solver must be writable float64 C- or F-ordered arrays; keep identity for customs
    U, V, Z = (a if isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.writeable
               else np.array(a, dtype=np.float64) for a in (U, V, Z))
    if return_errors:
        solver_object.final_error_links = (x_link, y_link)
    U, V, Z, n_iter = solver_object.fit_iterative_update(X, Y, U, V, Z)
    if return_errors:
        # (||X - f1(U V^T)||_F, ||Y - f2(V Z^T)||_F) evaluated on the device-resident data right after the fit
        return U, V, Z, n_iter, solver_object.final_errors_
    return U, V, Z, n_iter


class CMF(BaseEstimator, TransformerMixin):
    """Collective Matrix Factorization estimator; drop-in for `pycmf.CMF` (reference cmf.py:459-776).

    Finds U (n x k), V (d x k), Z (l x k) with X ~= f1(U V^T), Y ~= f2(V Z^T).  Parameters and
    attributes (`components`, `x_weights`, `y_weights`, `reconstruction_err_`, `n_iter_`,
    `n_components_`) are the reference's; `dtype`, `device`, `sampler`, `backend_options` are backend knobs.
    """

    def __init__(self, n_components=None, x_init=None, y_init=None, solver='mu', alpha='auto',
                 beta_loss='frobenius', tol=1e-4, max_iter=600,
                 random_state=None, l1_reg=0., l2_reg=0., verbose=0,
                 U_non_negative=True, V_non_negative=True, Z_non_negative=True,
                 x_link="linear", y_link="linear", hessian_pertubation=0.2, sg_sample_ratio=1.,
                 dtype="float32", device=None, sampler="auto", backend_options=None, init_on_device=False):
        self.n_components = n_components
        self.x_init = x_init
        self.y_init = y_init
        self.solver = solver
        self.alpha = alpha
        self.beta_loss = beta_loss
        self.tol = tol
        self.max_iter = max_iter
        self.random_state = random_state
        self.l1_reg = l1_reg
        self.l2_reg = l2_reg
        self.verbose = verbose
        self.U_non_negative = U_non_negative
        self.V_non_negative = V_non_negative
        self.Z_non_negative = Z_non_negative
        self.x_link = x_link
        self.y_link = y_link
        self.hessian_pertubation = hessian_pertubation
        self.sg_sample_ratio = sg_sample_ratio
        self.dtype = dtype
        self.device = device
        self.sampler = sampler
        self.backend_options = backend_options
        self.init_on_device = init_on_device

    def _backend_kw(self):
        return dict(dtype=self.dtype, device=self.device, sampler=self.sampler,
                    backend_options=self.backend_options, init_on_device=self.init_on_device)

    def fit_transform(self, X, Y, U=None, V=None, Z=None):
        """Learn a CMF model for X and Y and return (U, V, Z) (reference cmf.py:645-706)."""
        X = _validated(X)
        Y = _validated(Y)
        if X.shape[1] != Y.shape[0]:
            raise ValueError("Expected X.shape[1] == Y.shape[0], " +
                             "found X.shape = {}, Y.shape = {}".format(X.shape, Y.shape))
        U, V, Z, n_iter_, errs = collective_matrix_factorization(
            X=X, Y=Y, U=U, V=V, Z=Z, n_components=self.n_components, return_errors=True,
            x_init=self.x_init, y_init=self.y_init,
            solver=self.solver, alpha=self.alpha, beta_loss=self.beta_loss,
            tol=self.tol, max_iter=self.max_iter, l1_reg=self.l1_reg,
            l2_reg=self.l2_reg, random_state=self.random_state, verbose=self.verbose,
            U_non_negative=self.U_non_negative, V_non_negative=self.V_non_negative,
            Z_non_negative=self.Z_non_negative,
            x_link=self.x_link, y_link=self.y_link,
            hessian_pertubation=self.hessian_pertubation, sg_sample_ratio=self.sg_sample_ratio,
            **self._backend_kw())
        # unweighted sum of the two Frobenius errors (cmf.py:697-698)
        if errs is not None:
            self.reconstruction_err_ = errs[0] + errs[1]           # from the state still resident in HBM
        else:                                                      # (a solver without the hook: the NumPy stand-in of the tests)
            self.reconstruction_err_ = compute_factorization_error(X, U, V.T, self.x_link, self.beta_loss,
                                                                   dtype=self.dtype, device=self.device)
            self.reconstruction_err_ += compute_factorization_error(Y, V, Z.T, self.y_link, self.beta_loss,
                                                                    dtype=self.dtype, device=self.device)
        self.n_components_ = U.shape[1]
        self.x_weights = U
        self.components = V
        self.y_weights = Z
        self.n_iter_ = n_iter_
        return U, V, Z

    def fit(self, X, Y, **params):
        """Learn a CMF model for the data X and Y; returns self (reference cmf.py:708-724)."""
        self.fit_transform(X, Y, **params)
        return self

    def transform(self, X, Y):
        """Fit U and / or Z on new X / Y while keeping the components V fixed (reference cmf.py:726-747).
        Pass None for the matrix that should not be used."""
        assert(hasattr(self, "components"))
        update_U = X is not None
        update_Z = Y is not None
        alpha = 1 if Y is None else 0 if X is None else "auto"
        U = None if update_U else self.x_weights
        Z = None if update_Z else self.y_weights
        U, V, Z, n_iter_ = collective_matrix_factorization(
            X=X, Y=Y, U=U, V=self.components, Z=Z,
            n_components=self.n_components, x_init="custom", y_init="custom",
            solver=self.solver, alpha=alpha, beta_loss=self.beta_loss,
            tol=self.tol, max_iter=self.max_iter, l1_reg=self.l1_reg,
            l2_reg=self.l2_reg, random_state=self.random_state, verbose=self.verbose,
            U_non_negative=self.U_non_negative, V_non_negative=self.V_non_negative,
            Z_non_negative=self.Z_non_negative,
            update_U=update_U, update_V=False, update_Z=update_Z,
            x_link=self.x_link, y_link=self.y_link,
            hessian_pertubation=self.hessian_pertubation, sg_sample_ratio=self.sg_sample_ratio,
            **self._backend_kw())
        return U, V, Z

    def print_topic_terms(self, vectorizer, topn_words=10, importances=True):
        """Print the topics with their heaviest words (reference cmf.py:749-776).
        `vectorizer` is a fitted CountVectorizer / TfidfVectorizer."""
        names = vectorizer.get_feature_names_out() if hasattr(vectorizer, "get_feature_names_out") \
            else vectorizer.get_feature_names()
        idx_to_word = np.array(names)
        device = self._topk_device()
        if importances:
            _print_topic_terms_with_importances_from_matrices(
                self.x_weights, self.y_weights, idx_to_word, topn_words=topn_words, device=device)
        else:
            _print_topic_terms_from_matrix(self.x_weights, idx_to_word, topn_words=topn_words, device=device)

    def _topk_device(self):
        """The GPU the estimator fits on, for the per-topic top-k (None = host argsort when no GPU is visible: printing a
        fitted model must work wherever the pickled estimator is loaded)."""
        try:
            import torch
            if torch.cuda.is_available():
                return True if self.device is None else self.device
        except Exception:  # noqa: BLE001
            pass
        return None
