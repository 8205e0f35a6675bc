// rfa_tables.h -- host-side constant tables of the spectrum path (window, twiddles).
#pragma once
#include <math.h>

#include <vector>

#include "rfa_fft_core.cuh"

namespace rfa {

enum : int { WIN_BLACKMAN_REF = 0, WIN_HANN = 1, WIN_RECT = 2 };

// WIN_BLACKMAN_REF is NativeDsp.makeWindow (nativedsp/.../NativeDsp.kt:14-21): evaluated in
// double, cast once.  WIN_HANN is the north-star variant with the same (N-1) symmetry.
inline void make_window(int kind, int N, float *w) {
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < N; i++) {
        if (kind == WIN_BLACKMAN_REF)
            w[i] = (float)(0.42 - 0.5 * cos(2 * PI * i / (N - 1)) + 0.08 * cos(4 * PI * i / (N - 1)));
        else if (kind == WIN_HANN)
            w[i] = (float)(0.5 - 0.5 * cos(2 * PI * i / (N - 1)));
        else
            w[i] = 1.0f;
    }
}

// tw[t] = exp(-2*pi*i*t/N) in double, cast once (pffft computes its twiddles in float,
// pffft.c:1134-1170; double-then-cast is at least as accurate).
inline void make_twiddles(int N, cf *tw) {
    const double PI = 3.14159265358979323846;
    for (int t = 0; t < N; t++) {
        double a = -2.0 * PI * (double)t / (double)N;
        tw[t] = cf{(float)cos(a), (float)sin(a)};
    }
}

// Per-pass tables for the Stockham passes of an NL-point transform (layout in
// rfa_fft_core.cuh: pass_tw_offset + (r-1)*P + k  holds  exp(-2*pi*i*k*r/(P*R))).
inline std::vector<cf> make_pass_twiddles(int NL) {
    const double PI = 3.14159265358979323846;
    int lg = 0;
    while ((1 << lg) < NL) lg++;
    const int passes = plan_passes(lg);
    std::vector<cf> out;
    int P = 1;
    for (int pass = 0; pass < passes; pass++) {
        const int R = plan_radix(lg, pass);
        if (pass >= 1)
            for (int r = 1; r < R; r++)
                for (int k = 0; k < P; k++) {
                    const double a = -2.0 * PI * (double)k * (double)r / ((double)P * (double)R);
                    out.push_back(cf{(float)cos(a), (float)sin(a)});
                }
        P *= R;
    }
    if (out.empty()) out.push_back(cf{1.f, 0.f});
    return out;
}

}  // namespace rfa
