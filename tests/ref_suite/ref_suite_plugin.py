"""pytest plugin used by tests/test_reference_suite.py -- TEST INFRASTRUCTURE.

Lets the reference's OWN test file (/root/reference/tests/test_cmf.py, executed where it lies, never copied) run against
pycmf_b200: `import pycmf` resolves to pycmf_b200, the device backend is the float64 NumPy stand-in (tests/fake_backend.py,
so this exercises the host side: estimator, validation, initialisation, solver orchestration), and the scikit-learn modules
the 2018-era test file imports but scikit-learn has since removed (`sklearn.utils.testing`, `sklearn.decomposition.nmf`) are
provided as thin aliases of their modern equivalents."""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import sklearn.decomposition._nmf as _nmf  # noqa: E402

sys.modules.setdefault("sklearn.decomposition.nmf", _nmf)

shim = types.ModuleType("sklearn.utils.testing")


def _assert_raise_message(exc, message, fn, *args, **kwargs):
    try:
        fn(*args, **kwargs)
    except exc as e:
        assert message in str(e), "expected %r in %r" % (message, str(e))
        return
    raise AssertionError("%s not raised" % (exc,))


def _assert_no_warnings(fn, *args, **kwargs):
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = fn(*args, **kwargs)
    w = [x for x in w if not issubclass(x.category, (DeprecationWarning, FutureWarning, PendingDeprecationWarning))]
    assert not w, [str(x.message) for x in w]
    return out


class _ignore_warnings:
    def __init__(self, fn=None, category=Warning):
        self.fn, self.category = fn, category

    def __call__(self, fn):
        import functools

        @functools.wraps(fn)
        def wrapper(*a, **k):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", self.category)
                return fn(*a, **k)
        return wrapper

    def __enter__(self):
        self._cm = warnings.catch_warnings()
        self._cm.__enter__()
        warnings.simplefilter("ignore", self.category)

    def __exit__(self, *exc):
        return self._cm.__exit__(*exc)


def ignore_warnings(obj=None, category=Warning):
    if callable(obj):
        return _ignore_warnings(category=category)(obj)
    return _ignore_warnings(category=category)


shim.assert_true = lambda x, msg=None: (_ for _ in ()).throw(AssertionError(msg)) if not x else None
shim.assert_false = lambda x, msg=None: (_ for _ in ()).throw(AssertionError(msg)) if x else None
shim.assert_raise_message = _assert_raise_message
shim.assert_no_warnings = _assert_no_warnings
shim.assert_array_equal = np.testing.assert_array_equal
shim.assert_array_almost_equal = np.testing.assert_array_almost_equal
shim.assert_almost_equal = np.testing.assert_almost_equal
shim.assert_less = lambda a, b, msg=None: np.testing.assert_array_less(a, b)
shim.assert_greater = lambda a, b, msg=None: np.testing.assert_array_less(b, a)
shim.ignore_warnings = ignore_warnings
sys.modules["sklearn.utils.testing"] = shim

import pycmf_b200  # noqa: E402
import pycmf_b200.analysis  # noqa: E402
import pycmf_b200.device as _device  # noqa: E402
from fake_backend import FakeBackend  # noqa: E402

_device.CudaBackend = lambda device=None, dtype=None, options=None: FakeBackend()
sys.modules["pycmf"] = pycmf_b200
sys.modules["pycmf.analysis"] = pycmf_b200.analysis
