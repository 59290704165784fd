"""Whole fits on degenerate shapes (one row, one label column, rank one, an all-zero CSR row, two columns, sample sets of one
and zero indices, single-factor updates) through the C ABI against the oracle, float64 <= 1e-9 and float32 within the
fp32 bars.  The same cases are pinned to the live reference on the CPU (tests/test_oracle_vs_reference.py).

Written after the round's GPU budget was spent, so this file has NOT been run on a GPU by the builder: it runs in a
subprocess (a faulting kernel cannot poison the CUDA context of the other GPU tests; the file sorts last) and is marked
xfail(strict=False) -- an XPASS in the driver's log is the evidence that these shapes work, an XFAIL names the shape that
does not."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np
from helpers import run_oracle, draw_masks_for_case, rel_fro
from oracle.cases import EDGE, make_edge_case
from pycmf_b200.cmf_solvers import MUSolver, NewtonSolver

bad = 0
for name in sorted(EDGE):
    case = make_edge_case(name)
    masks = draw_masks_for_case(case)
    hist_o, Uo, Vo, Zo = run_oracle(case, masks)
    for dtype, obj_tol, fac_tol in (("float64", 1e-9, 1e-9), ("float32", 1e-4, 1e-3)):
        try:
            p = dict(case["params"]); solver = p.pop("solver")
            s = (MUSolver if solver == "mu" else NewtonSolver)(max_iter=case["iters"], tol=0, random_state=case["rng_seed"],
                                                               dtype=dtype, **p)
            s.history, s.masks_per_iter = [], masks
            U, V, Z = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
            s.fit_iterative_update(case["X"], case["Y"], U, V, Z)
            eo = np.abs(np.asarray(s.history) - hist_o[1:]).max() / max(np.abs(hist_o).max(), 1e-300)
            ef = max(rel_fro(a, b) for a, b in ((U, Uo), (V, Vo), (Z, Zo)))
            ok = eo <= obj_tol and ef <= fac_tol
            print("%s %-26s %-8s objective %.2e factors %.2e" % ("OK  " if ok else "FAIL", name, dtype, eo, ef), flush=True)
        except Exception as e:
            ok = False
            print("FAIL %-26s %-8s %s: %s" % (name, dtype, type(e).__name__, str(e)[:200]), flush=True)
        bad += not ok
print("EDGE_DONE bad=%d" % bad)
'''


@pytest.mark.xfail(strict=False, reason="added after the round's GPU budget was spent: not yet run on a GPU by the builder")
def test_fits_on_degenerate_shapes_match_the_oracle(tmp_path):
    script = tmp_path / "edge_worker.py"
    script.write_text(_WORKER)
    out = subprocess.run([sys.executable, str(script), ROOT], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=150).stdout
    print(out[-4000:])
    assert "EDGE_DONE bad=0" in out, "\n".join(ln for ln in out.splitlines() if ln.startswith("FAIL"))[:3000] or out[-2000:]
