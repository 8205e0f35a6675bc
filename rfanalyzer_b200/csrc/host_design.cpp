// host_design.cpp -- see host_design.h.  Compiled with -ffp-contract=off: the JVM never
// fuses a multiply into an add, and these tables must match it bit for bit.
#include "host_design.h"

#include <cmath>

namespace rfa {
namespace design {

static const double kPi = 3.14159265358979323846;
static const float kPiF = (float)3.14159265358979323846;

int java_d2i(double v) {
    if (std::isnan(v)) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return -2147483647 - 1;
    return (int)v;
}

int optimal_cosine_length(int sample_rate, int cosine_frequency) {
    const double cycle = sample_rate / std::fabs((double)cosine_frequency);
    int best = java_d2i(cycle);
    double best_err = std::fabs(best - cycle);
    for (int i = 1; i * cycle < 500; i++) {
        const double x = i * cycle;
        const int xi = java_d2i(x);
        if (std::fabs(x - xi) < best_err) {
            best = xi;
            best_err = std::fabs(best - x);
        }
    }
    return best;
}

void nco_tables(int fmt, int sample_rate, int mix, int *effective, std::vector<float> *cos_t,
                std::vector<float> *sin_t) {
    const int amix = mix < 0 ? -mix : mix;
    if (mix == 0 || sample_rate / amix > 500) mix += sample_rate;
    if (effective) *effective = mix;
    int len = optimal_cosine_length(sample_rate, mix);
    if (len < 0) len = 0;
    cos_t->assign((size_t)len, 0.f);
    sin_t->assign((size_t)len, 0.f);
    if (fmt == 2) {  // 16-bit: angle = ((2*pi*f)/fs) * t
        const double w = (2.0 * kPi * mix) / (double)sample_rate;
        for (int t = 0; t < len; t++) {
            const double a = w * t;
            (*cos_t)[t] = (float)std::cos(a);
            (*sin_t)[t] = (float)std::sin(a);
        }
    } else {  // 8-bit: angle = 2*pi*f*t / (float)fs, evaluated left to right
        const double fs = (double)(float)sample_rate;
        for (int t = 0; t < len; t++) {
            const double a = 2 * kPi * mix * t / fs;
            (*cos_t)[t] = (float)std::cos(a);
            (*sin_t)[t] = (float)std::sin(a);
        }
    }
}

static double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double half = x / 2.0;
    for (int k = 1;; k++) {
        const double q = half / k;
        term *= q * q;
        sum += term;
        if (term < 1e-12) break;
    }
    return sum;
}

float tap_window(int kind, double beta, int n, int N) {
    if (kind == 0) {  // Blackman: each cosine is cast to float before the float arithmetic
        const float c1 = (float)std::cos(2.0 * kPi * n / (N - 1));
        const float c2 = (float)std::cos(4.0 * kPi * n / (N - 1));
        const float a = 0.5f * c1, b = 0.08f * c2;
        const float t = 0.42f - a;
        return t + b;
    }
    if (kind == 1) {  // Hamming
        const float c1 = (float)std::cos(2.0 * kPi * n / (N - 1));
        const float a = 0.46f * c1;
        return 0.54f - a;
    }
    const double inv = 1.0 / bessel_i0(beta);  // Kaiser
    if (n == 0 || n == N - 1) return (float)inv;
    const double t = 2.0 * n * (1.0 / (double)(N - 1)) - 1.0;
    return (float)(bessel_i0(beta * std::sqrt(1.0 - t * t)) * inv);
}

// truncated-sinc prototype shared by the low-pass and band-pass designs
static std::vector<float> sinc_prototype(int ntaps, float gain, float fs, float cutoff, int window_kind,
                                         double beta) {
    std::vector<float> taps((size_t)ntaps);
    const int M = (ntaps - 1) / 2;
    const float two_pi = 2 * kPiF;
    const float scaled = two_pi * cutoff;
    const float w0 = scaled / fs;
    for (int n = -M; n <= M; n++) {
        const float win = tap_window(window_kind, beta, n + M, ntaps);
        float ideal;
        if (n == 0) {
            ideal = w0 / kPiF;
        } else {
            const float arg = (float)n * w0;
            const float s = (float)std::sin((double)arg);
            const float den = (float)n * kPiF;
            ideal = s / den;
        }
        taps[(size_t)(n + M)] = ideal * win;
    }
    float dc = taps[(size_t)M];
    for (int n = 1; n <= M; n++) {
        const float twice = 2 * taps[(size_t)(n + M)];
        dc += twice;
    }
    const float norm = gain / dc;
    for (auto &t : taps) t *= norm;
    return taps;
}

std::vector<float> lowpass_taps(float gain, float fs, float cutoff, float tw, float att, int window_kind,
                                double beta, int max_taps) {
    if (fs <= 0.0 || cutoff <= 0.0 || cutoff > fs / 2 || tw <= 0) return {};
    const float att_fs = att * fs;
    int ntaps = java_d2i((double)att_fs / (22.0 * (double)tw));
    if (max_taps > 0 && ntaps > max_taps) ntaps = max_taps;
    if (ntaps < 0) return {};
    if ((ntaps & 1) == 0) ntaps++;
    return sinc_prototype(ntaps, gain, fs, cutoff, window_kind, beta);
}

bool bandpass_taps(float gain, float fs, float lo, float hi, float tw, float att, std::vector<float> *re,
                   std::vector<float> *im) {
    if (fs <= 0.0) return false;
    if ((double)lo < (double)fs * -0.5 || (double)hi > (double)fs * 0.5) return false;
    if (lo >= hi || tw <= 0) return false;
    const float att_fs = att * fs;
    int ntaps = java_d2i((double)att_fs / (22.0 * (double)tw));
    if (ntaps < 0) return false;
    if ((ntaps & 1) == 0) ntaps++;
    const float lp_cut = (hi - lo) / 2.0f;
    std::vector<float> lp = sinc_prototype(ntaps, gain, fs, lp_cut, 0, 0.0);
    re->assign((size_t)ntaps, 0.f);
    im->assign((size_t)ntaps, 0.f);
    const float band_sum = hi + lo;
    const float scaled = kPiF * band_sum;
    const float step = scaled / fs;
    float phase = -step * (float)(ntaps / 2);
    for (int i = 0; i < ntaps; i++) {
        (*re)[(size_t)i] = lp[(size_t)i] * (float)std::cos((double)phase);
        (*im)[(size_t)i] = lp[(size_t)i] * (float)std::sin((double)phase);
        phase += step;
    }
    return true;
}

int gcd(int a, int b) {
    int x = a < 0 ? -a : a, y = b < 0 ? -b : b;
    while (y != 0) {
        const int t = y;
        y = x % y;
        x = t;
    }
    return x;
}

void limit_denominator(int num, int den, int max_den, int *out_num, int *out_den) {
    const double target = (double)num / (double)den;
    const int g = gcd(num, den);
    if (den / g <= max_den) {
        *out_num = num / g;
        *out_den = den / g;
        return;
    }
    int ln = 0, ld = 1, un = 1, ud = 0;  // Farey mediants between 0/1 and 1/0
    for (;;) {
        const int mn = ln + un, md = ld + ud;
        if (md > max_den) break;
        if ((double)mn / md < target) {
            ln = mn;
            ld = md;
        } else {
            un = mn;
            ud = md;
        }
    }
    const double le = std::fabs(target - (double)ln / ld), ue = std::fabs(target - (double)un / ud);
    if (le < ue) {
        *out_num = ln;
        *out_den = ld;
    } else {
        *out_num = un;
        *out_den = ud;
    }
}

std::vector<float> resampler_taps(int interpolation, int decimation, float fractional_bw, int max_taps) {
    const double halfband = 0.5;
    const float rate = (float)interpolation / (float)decimation;
    float trans, mid;
    if (rate >= 1.0f) {
        trans = (float)(halfband - (double)fractional_bw);
        mid = (float)(halfband - (double)trans / 2.0);
    } else {
        trans = (float)((double)rate * (halfband - (double)fractional_bw));
        mid = (float)((double)rate * halfband - (double)trans / 2.0);
    }
    return lowpass_taps((float)interpolation, (float)interpolation, mid, trans, 72.22087f, 2, 7.0,
                        max_taps * interpolation);
}

void channel_bins(int n, long long frequency, int sample_rate, long long chan_start, long long chan_end, int *b0,
                  int *b1) {
    const float per_hz = n / (float)sample_rate;
    const long long f0 = frequency - sample_rate / 2;
    auto clamp = [n](int v) { return v < 0 ? 0 : (v > n ? n : v); };
    *b0 = clamp(java_d2i((double)((float)(chan_start - f0) * per_hz)));
    *b1 = clamp(java_d2i((double)((float)(chan_end - f0) * per_hz)));
}

}  // namespace design
}  // namespace rfa
