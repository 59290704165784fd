"""Import the UNMODIFIED reference (smn-ailab/PyCMF) from /root/reference -- TEST INFRASTRUCTURE.

The reference does ``from sklearn.decomposition.nmf import _beta_divergence`` (cmf_solvers.py:7);
that private module was renamed ``_nmf`` in scikit-learn 0.22, so we alias it before import.
Nothing is copied or edited.  /root/reference exists only in the build container, never on the
GPU box: callers must handle ``load_reference() is None``.
"""
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("PYCMF_REFERENCE_ROOT", "/root/reference")


def load_reference():
    """Return the reference ``pycmf`` package, or None when /root/reference is absent."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "pycmf")):
        return None
    import sklearn.decomposition._nmf as _nmf
    sys.modules.setdefault("sklearn.decomposition.nmf", _nmf)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pycmf
        import pycmf.cmf_solvers  # noqa: F401
    return pycmf
