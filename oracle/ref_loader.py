"""Import the UNMODIFIED reference (smn-ailab/PyCMF) -- TEST INFRASTRUCTURE.

Search order: $PYCMF_REFERENCE_ROOT, then `baseline/_ref/` (the reference as installed by
`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`, done by
`__graft_entry__.build()` in the build container: the reference's own .py files plus its compiled Cython module
`pycmf.cmf_newton_solver`; git-ignored, but it travels to the GPU box), then `/root/reference` (build container only).

The reference does ``from sklearn.decomposition.nmf import _beta_divergence`` (cmf_solvers.py:7);
that private module was renamed ``_nmf`` in scikit-learn 0.22, so we alias it before import.
Nothing is copied into the repository's history and nothing is edited.
"""
import glob
import os
import sys
import warnings

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED = os.path.join(REPO, "baseline", "_ref")
SOURCE = "/root/reference"


def reference_root():
    """Directory that holds the reference's `pycmf` package, or None."""
    for root in (os.environ.get("PYCMF_REFERENCE_ROOT"), INSTALLED, SOURCE):
        if root and os.path.isfile(os.path.join(root, "pycmf", "cmf_solvers.py")):
            return root
    return None


REFERENCE_ROOT = reference_root()


def load_reference():
    """Return the reference ``pycmf`` package, or None when it is not available."""
    root = reference_root()
    if root is None:
        return None
    import sklearn.decomposition._nmf as _nmf
    sys.modules.setdefault("sklearn.decomposition.nmf", _nmf)
    if root not in sys.path:
        sys.path.insert(0, root)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pycmf
        import pycmf.cmf_solvers  # noqa: F401
    return pycmf


def load_reference_cython():
    """The reference's compiled Cython twin ``pycmf.cmf_newton_solver`` (cmf_newton_solver.pyx; dead code at the
    reference's HEAD, `USE_CYTHON = False`, cmf_solvers.py:11), or None when no compiled copy exists."""
    root = reference_root()
    if root is None or not glob.glob(os.path.join(root, "pycmf", "cmf_newton_solver*.so")):
        return None
    if load_reference() is None:
        return None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pycmf.cmf_newton_solver as m
    return m
