"""pycmf_b200: B200-native backend for PyCMF's fit loop (drop-in for `pycmf.CMF`)."""
from .cmf import CMF, collective_matrix_factorization, compute_factorization_error  # noqa: F401
from . import analysis  # noqa: F401

__all__ = ["CMF", "collective_matrix_factorization", "compute_factorization_error", "analysis"]
