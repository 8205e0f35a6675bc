"""Runs tests/lab (the experimental kernels) against librfa_b200_lab.so in a subprocess: the product library
librfa_b200.so does not carry those kernels, and one process binds one build of the library."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lab_kernels_against_the_lab_build():
    from rfanalyzer_b200 import _lib
    if not os.path.exists(_lib.LAB_LIB_PATH):
        pytest.skip("librfa_b200_lab.so is not built (make -C rfanalyzer_b200/csrc lab)")
    env = dict(os.environ, RFA_B200_LIB=_lib.LAB_LIB_PATH)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "lab"), "-x", "-q", "-m", "gpu",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1800)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    assert r.returncode == 0, "lab suite failed:\n" + tail + "\n" + r.stderr[-2000:]
    assert " passed" in tail


def test_product_library_refuses_lab_options(gpu_ctx):
    import rfanalyzer_b200 as rfa
    for name in ("kernel", "fs_fused", "fourstep"):
        with pytest.raises(rfa.RfaError) as e:
            gpu_ctx.set_option(name, 1)
        assert e.value.code == rfa._lib.ERR_UNSUPPORTED
    assert gpu_ctx.get_option("staged") == 1
    with gpu_ctx.options(staged=0):
        assert gpu_ctx.get_option("staged") == 0
    assert gpu_ctx.get_option("staged") == 1
