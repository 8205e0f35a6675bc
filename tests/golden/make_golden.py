"""Writes tests/golden/*.npz.

 * fir_kat.npz  -- the reference's own known-answer vectors, transcribed from
   app/src/androidTest/java/com/mantz_it/rfanalyzer/ApplicationTest.kt:55-121 (testFirFilter)
   and :165-170 (testFirFilter2).  They pin createLowPassTaps + Blackman + FirFilter.filter.
 * spectrum_ref.npz -- outputs of the reference's own native code (pffft.c + nativedsp.cpp,
   compiled in place into oracle/_ref) on the synthetic generator's frames, so the GPU box can
   check against the real reference numbers even without /root/reference.
Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TEST = "/root/reference/app/src/androidTest/java/com/mantz_it/rfanalyzer/ApplicationTest.kt"


def float_arrays(text):
    out = []
    for m in re.finditer(r"floatArrayOf\((.*?)\)", text, re.S):
        vals = [float(v.strip().rstrip("f")) for v in m.group(1).replace("\n", " ").split(",") if v.strip()]
        out.append(np.array(vals, dtype=np.float32))
    return out


def main():
    text = open(REF_TEST).read()
    arrs = [a for a in float_arrays(text) if len(a) in (32, 63)]
    assert [len(a) for a in arrs[:4]] == [32, 32, 63, 63], [len(a) for a in arrs]
    np.savez(os.path.join(HERE, "fir_kat.npz"), re1=arrs[0], im1=arrs[1], re2=arrs[2], im2=arrs[3])

    out = {}
    for fmt, name in ((O.FMT_S8, "s8"), (O.FMT_U8, "u8"), (O.FMT_S16LE, "s16")):
        for n in (1024, 4096, 65536):
            frames = 2 if n <= 4096 else 1
            iq = O.synth_iq(fmt, n * frames, first=12345)
            rows, peaks, avg = O.ref_spectrum_run(fmt, iq, n, 1, nthreads=1)
            out[f"{name}_{n}_rows"] = rows
            out[f"{name}_{n}_peaks"] = peaks
            out[f"{name}_{n}_avg"] = avg
    np.savez_compressed(os.path.join(HERE, "spectrum_ref.npz"), **out)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
