"""include/rfa_b200.hpp -- the C++ host mirror of the reference's classes -- compiles, links against
librfa_b200.so and runs the reference's FIR known-answer test (GPU) / fails loudly (no GPU)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def build():
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
    libdir = os.path.join(ROOT, "rfanalyzer_b200", "lib")
    deps = [src, os.path.join(ROOT, "include", "rfa_b200.hpp"), os.path.join(ROOT, "include", "rfa_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-o", EXE, src, "-L" + libdir, "-lrfa_b200",
                        "-Wl,-rpath," + libdir], check=True)


def test_host_mirror_compiles_and_fails_loudly_without_gpu():
    import torch
    build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_mirror_on_gpu():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0 and "host mirror ok" in r.stdout, r.stdout + r.stderr
