"""libnativedsp.so -- the drop-in for the reference's JNI library -- exports the reference's two
mangled symbols; driven through a fake JNIEnv (no JVM in this image)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "rfanalyzer_b200", "lib", "libnativedsp.so")
SYMS = ["Java_com_mantz_1it_nativedsp_NativeDsp_performFFT", "Java_com_mantz_1it_nativedsp_NativeDsp_performFFTAndLogMag"]


@pytest.fixture(scope="module")
def fake():
    src = os.path.join(ROOT, "tests", "jni", "fake_jni.cpp")
    so = os.path.join(ROOT, "tests", "jni", "libfake_jni.so")
    if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    lib = C.CDLL(so)
    lib.fake_jni_call.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return lib


def test_shim_exports_the_reference_symbols():
    """nativedsp.cpp:19-21, :44-46 -- names as javah mangles NativeDsp.performFFT*."""
    lib = C.CDLL(SHIM)
    for s in SYMS:
        assert hasattr(lib, s)


def test_jni_function_table_slots(fake):
    """JNI spec function-table slots: GetArrayLength 171, GetFloatArrayRegion 205, SetFloatArrayRegion 213."""
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    total = fake.fake_jni_slot_offsets(C.byref(a), C.byref(b), C.byref(c))
    assert (a.value, b.value, c.value) == (171, 205, 213) and total == 235


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1024, 16384])
def test_jni_entry_points_on_gpu(fake, oracle, n):
    shim = C.CDLL(SHIM)
    x = np.random.default_rng(n).standard_normal(2 * n).astype(np.float32)
    out = np.zeros(2 * n, np.float32)
    fn = C.cast(getattr(shim, SYMS[0]), C.c_void_p)
    fake.fake_jni_call(fn, x.ctypes.data, 2 * n, out.ctypes.data, 2 * n)
    ref = np.fft.fft(x[0::2].astype(np.float64) + 1j * x[1::2].astype(np.float64))
    assert np.abs((out[0::2] + 1j * out[1::2]) - ref).max() / np.abs(ref).max() < 1e-6
    mag = np.zeros(n, np.float32)
    fn = C.cast(getattr(shim, SYMS[1]), C.c_void_p)
    fake.fake_jni_call(fn, x.ctypes.data, 2 * n, mag.ctypes.data, n)
    want = np.empty(n, np.float32)
    oracle.lib().orc_fft_logmag(x, want, n)
    assert np.abs(mag - want).max() < 0.01
