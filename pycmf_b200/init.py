"""Host-side factor initialisation with the reference's semantics (pycmf/cmf.py:41-212).

Parity of the fit loop is defined "from the same initial factors", so initialisation stays on the
host and follows the reference's algorithms: 'random' (scaled |N(0,1)|), 'svd' (randomized SVD split
evenly between the factors), 'nndsvd' / 'nndsvda' / 'nndsvdar' (Boutsidis & Gallopoulos 2008) and
'custom'.  Returned as (A, B) with M ~= A B^T, A: rows x k, B: cols x k.
"""
import warnings
from math import sqrt

import numpy as np
from sklearn.utils import check_array, check_random_state
from sklearn.utils.extmath import randomized_svd
from sklearn.utils.validation import check_non_negative

NNDSVD_KINDS = ("nndsvd", "nndsvda", "nndsvdar")


def _vec_norm(v):
    return sqrt(float(np.dot(v, v)))


def _random_init(M, k, random_state, non_negative):
    scale = np.sqrt(np.abs(M.mean()) / k)                      # cmf.py:111
    rng = check_random_state(random_state)
    A = scale * rng.randn(M.shape[0], k)
    Bt = scale * rng.randn(k, M.shape[1])
    if non_negative:
        A, Bt = np.abs(A), np.abs(Bt)
    return A, Bt


def _svd_init(M, k, random_state):
    rows, cols = M.shape
    if min(rows, cols) < k:
        warnings.warn('The number of components is smaller than the rank in svd initialization.' +
                      'The input will be padded with zeros to compensate for the lack of singular values.')
    Us, s, Vt = randomized_svd(M, k, random_state=random_state)
    if k > cols:                                               # pad to the requested width (cmf.py:129-138)
        r = s.shape[0]
        Us = np.hstack([Us, np.zeros((Us.shape[0], k - Us.shape[1]))])
        Vt = np.vstack([Vt, np.zeros((k - Vt.shape[0], Vt.shape[1]))])
        s = np.concatenate([s, np.zeros(k - r)])
    root = np.sqrt(s)
    return Us * root[None, :], root[:, None] * Vt


def _nndsvd_init(M, k, kind, eps, random_state, non_negative):
    if not non_negative:
        warnings.warn('%s results in non-negative constrained factors,' % kind +
                      'so SVD initialization should provide better initial estimate')
    Us, s, Vt = randomized_svd(M, k, random_state=random_state)
    A, Bt = np.zeros(Us.shape), np.zeros(Vt.shape)
    # leading triplet is sign-definite
    A[:, 0] = np.sqrt(s[0]) * np.abs(Us[:, 0])
    Bt[0, :] = np.sqrt(s[0]) * np.abs(Vt[0, :])
    for j in range(1, k):
        x, y = Us[:, j], Vt[j, :]
        xp, yp = np.maximum(x, 0), np.maximum(y, 0)
        xn, yn = np.abs(np.minimum(x, 0)), np.abs(np.minimum(y, 0))
        xp_n, yp_n, xn_n, yn_n = _vec_norm(xp), _vec_norm(yp), _vec_norm(xn), _vec_norm(yn)
        pos, neg = xp_n * yp_n, xn_n * yn_n
        if pos > neg:
            u, v, sigma = xp / xp_n, yp / yp_n, pos
        else:
            u, v, sigma = xn / xn_n, yn / yn_n, neg
        scale = np.sqrt(s[j] * sigma)
        A[:, j] = scale * u
        Bt[j, :] = scale * v
    A[A < eps] = 0
    Bt[Bt < eps] = 0
    if kind == "nndsvda":
        avg = M.mean()
        A[A == 0] = avg
        Bt[Bt == 0] = avg
    elif kind == "nndsvdar":
        rng = check_random_state(random_state)
        avg = M.mean()
        A[A == 0] = abs(avg * rng.randn(len(A[A == 0])) / 100)
        Bt[Bt == 0] = abs(avg * rng.randn(len(Bt[Bt == 0])) / 100)
    return A, Bt


def _initialize_mf(M, n_components, init=None, eps=1e-6, random_state=None, non_negative=False):
    """Initial guess M ~= A B^T; same defaults and errors as the reference (cmf.py:41-202)."""
    if non_negative:
        check_non_negative(M, "MF initialization")
    cols = M.shape[1]
    if init is None:
        if n_components < cols:
            init = 'nndsvdar' if non_negative else 'svd'
        else:
            init = 'random'
    if init == 'random':
        A, Bt = _random_init(M, n_components, random_state, non_negative)
    elif init == 'svd':
        if non_negative:
            raise ValueError('SVD initialization incompatible with NMF (use nndsvd instead)')
        A, Bt = _svd_init(M, n_components, random_state)
    elif init in NNDSVD_KINDS:
        A, Bt = _nndsvd_init(M, n_components, init, eps, random_state, non_negative)
    else:
        raise ValueError("Invalid init argument")
    return A, Bt.T


def _check_init(A, shape, whom, non_negative):
    A = check_array(A)
    if np.shape(A) != shape:
        raise ValueError('Array with wrong shape passed to %s. Expected %s, '
                         'but got %s ' % (whom, shape, np.shape(A)))
    if non_negative:
        check_non_negative(A, whom)
        if np.max(A) == 0:
            raise ValueError('Array passed to %s is full of zeros.' % whom)


def _init_custom(A, M, n_components, idx, non_negative=False, random_state=None):
    """cmf.py:205-212: a supplied array is validated and returned BY IDENTITY, else random init."""
    if A is not None:
        _check_init(A, (M.shape[idx], n_components), "CMF (input {})".format(idx), non_negative)
        return A
    return _initialize_mf(M, n_components, init="random", random_state=random_state,
                          non_negative=non_negative)[idx]
