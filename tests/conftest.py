import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# tests/lab exercises the experimental kernels of librfa_b200_lab.so; it is collected only when the package is
# pointed at that build (tests/test_lab_gpu.py does so in a subprocess)
collect_ignore_glob = [] if os.environ.get("RFA_B200_LIB", "").endswith("_lab.so") else ["lab/*"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (C restatement; built on demand with gcc)."""
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def emu():
    """CPU emulation of the fused spectrum kernel's per-thread phases (tests/emu)."""
    import ctypes as C
    src = os.path.join(ROOT, "tests", "emu", "emu_spectrum.cpp")
    so = os.path.join(ROOT, "tests", "emu", "libemu_spectrum.so")
    deps = [src] + [os.path.join(ROOT, "rfanalyzer_b200", "csrc", f)
                    for f in ("rfa_fft_core.cuh", "spectrum_kernel.cuh", "spectrum2_kernel.cuh", "spectrum64_kernel.cuh", "fourstep_kernel.cuh", "fourstep_cluster.cuh", "rfa_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    lib = C.CDLL(so)
    lib.emu_spectrum_avg.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                                     C.c_void_p, C.c_int]
    lib.emu_spectrum_avg.restype = C.c_int
    lib.emu_spectrum2_avg.argtypes = [C.c_int] * 3 + [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                                      C.c_int, C.c_longlong]
    lib.emu_spectrum2_avg.restype = C.c_int
    lib.emu_spectrum64.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
    lib.emu_spectrum64.restype = C.c_int
    lib.emu_fourstep_spectrum.argtypes = [C.c_int] * 3 + [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong]
    lib.emu_fourstep_spectrum.restype = C.c_int
    lib.emu_cluster_spectrum.argtypes = lib.emu_fourstep_spectrum.argtypes
    lib.emu_cluster_spectrum.restype = C.c_int
    return lib


@pytest.fixture(scope="session")
def gpu_ctx():
    """A librfa_b200 context on cuda:0, ordered on a private torch stream."""
    import torch
    import rfanalyzer_b200 as rfa
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    stream = torch.cuda.Stream()
    ctx = rfa.Context(0, stream)
    ctx.torch_stream = stream
    return ctx
