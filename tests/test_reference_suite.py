"""Runs the reference's OWN test-suite (/root/reference/tests/test_cmf.py, in place) against pycmf_b200 with the NumPy
stand-in backend: a user of `pycmf.CMF` who switches the import must see the same behaviour.  Where /root/reference is
mounted only (the build container)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(os.environ.get("PYCMF_REFERENCE_ROOT", "/root/reference"), "tests", "test_cmf.py")


@pytest.mark.skipif(not os.path.exists(REF_TESTS), reason="/root/reference is not mounted here")
def test_reference_test_suite_passes_against_pycmf_b200(tmp_path):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "ref_suite"), ROOT]))
    cmd = [sys.executable, "-m", "pytest", REF_TESTS, "-q", "-p", "ref_suite_plugin", "-p", "no:cacheprovider",
           "--rootdir", str(tmp_path), "-W", "ignore"]
    out = subprocess.run(cmd, cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=900).stdout
    m = re.search(r"(\d+) passed", out)
    failed = re.search(r"(\d+) failed", out)
    assert m is not None and failed is None, out[-4000:]
    assert int(m.group(1)) >= 30, out[-2000:]
