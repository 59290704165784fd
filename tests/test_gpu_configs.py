"""Parity on the BASELINE.json configurations themselves, as bench.py's generator makes them:
C1 and C2 at full size, C3 / C4 / C5 row-scaled to what the CPU oracle finishes in seconds (oracle/parity.py says the
scale).  Bars are north_star's: per-iteration objective within 1e-4 relative, final U, V, Z within 1e-3 relative
Frobenius in fp32; 1e-9 on the fp64 path.  Every call goes through the C ABI."""
import pytest

pytestmark = pytest.mark.gpu


def _run(name, dtype, shape=None):
    import torch
    from bench import make_solver
    from oracle.parity import run_parity
    from pycmf_b200.device import CudaBackend
    from pycmf_b200.sharding import Comm
    be = CudaBackend(device=0, dtype=dtype)
    try:
        return run_parity(name, be, Comm(), make_solver, dtype=dtype, shape=shape)
    finally:
        be.close()
        torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5"])
def test_baseline_config_fp32(name):
    res = _run(name, "float32")
    assert res["pass"], res


@pytest.mark.parametrize("name,shape", [("c1", None), ("c2", (0.1, 1.0, 3)), ("c3", (0.002, 1.0, 3)),
                                        ("c5", (0.005, 1.0, 2))])
def test_baseline_config_fp64(name, shape):
    res = _run(name, "float64", shape)
    assert res["pass"], res
