// detect.cu -- signal detectors on waterfall rows.
//
// Reference (app/src/main/java/com/mantz_it/rfanalyzer/ui/MainViewModel.kt): every detector reads the newest
// waterfall row and reduces a window of bins to (peak, average):
//   getAverageSignalLevel        :1392-1414   whole row, average
//   detectSignal                 :1416-1461   whole row, peak + average, detection mode
//   detectSignalsInFFT           :1463-1550   +-2 bins around every step of the scan grid
//   detectIEMChannelsInFFT       :861-935     +-100 kHz around each channel (at least +-5 bins)
//   detectAirCommSignal[AtFrequency] :1151-1250  +-12.5 kHz (at least +-3 bins), peak only
//   groupSignals / finalizeGroup :1552-1607   neighbouring detections merged
// `windowData.maxOrNull()` is a NaN-propagating float max, `windowData.average().toFloat()` a double sum in index
// order divided by the count.  Here one warp reduces one window (rows stay in HBM, thousands of windows per
// launch); the double partial sums are added in tree order, which equals the sequential sum whenever that sum is
// exact -- it is for dB rows, whose 24-bit values span far fewer than 53 bits -- and is otherwise within one
// float ulp after the final rounding.
// The scalar host functions keep the JVM's arithmetic: Long -> Float conversions, float division, truncating
// saturating toInt().
#include <math.h>

#include <algorithm>
#include <vector>

#include "capi_core.h"

using namespace rfa;

namespace {

__device__ __forceinline__ float max_nan(float a, float b) {  // Math.max: NaN if either is NaN
    return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b);
}

__global__ void __launch_bounds__(256) detect_windows_kernel(const float *rows, long long row_stride, int n,
                                                              const rfa_detect_window *win, int nwin, float *peak,
                                                              float *avg) {
    const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= nwin) return;
    const rfa_detect_window q = win[w];
    const int lo = q.start < 0 ? 0 : q.start, hi = q.end > n - 1 ? n - 1 : q.end;
    const float *p = rows + q.row * row_stride;
    double sum = 0.0;
    float mx = -INFINITY;
    bool any = false;
    for (int i = lo + lane; i <= hi; i += 32) {
        const float v = p[i];
        sum += (double)v;
        mx = any ? max_nan(mx, v) : v;
        any = true;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double s2 = __shfl_xor_sync(0xffffffffu, sum, o);
        const float m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        const bool a2 = __shfl_xor_sync(0xffffffffu, any ? 1 : 0, o) != 0;
        sum += s2;
        if (a2) mx = any ? max_nan(mx, m2) : m2;
        any = any || a2;
    }
    if (lane == 0) {
        const int cnt = hi - lo + 1;
        // empty window: maxOrNull() == null -> the callers skip it; average() of nothing is NaN
        peak[w] = cnt > 0 ? mx : __int_as_float(0x7fc00000);
        avg[w] = cnt > 0 ? (float)(sum / (double)cnt) : __int_as_float(0x7fc00000);
    }
}

// Kotlin/Java float -> int: truncation toward zero, saturating, NaN -> 0
int jvm_to_int(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}

float resolution(long long sample_rate, int n) { return (float)sample_rate / (float)n; }

}  // namespace

extern "C" {

int rfa_detect_windows(rfa_ctx *c, const float *rows, long long row_stride, int n, const rfa_detect_window *win,
                       int nwin, float *peak, float *avg, int mem_win, int mem_out) {
    RFA_REQUIRE(c && rows && peak && avg, "rfa_detect_windows: NULL argument");
    RFA_REQUIRE(n > 0 && row_stride >= n, "bad row geometry");
    RFA_REQUIRE(nwin >= 0, "negative window count");
    if (nwin == 0) return RFA_OK;
    RFA_REQUIRE(win != nullptr, "rfa_detect_windows: NULL windows");
    if (int rc = c->use()) return rc;
    const rfa_detect_window *dwin = win;
    if (mem_win == RFA_MEM_HOST) {
        for (int i = 0; i < nwin; i++) RFA_REQUIRE(win[i].row >= 0, "window %d: negative row", i);
        if (int rc = c->stage[5].ensure((size_t)nwin * sizeof(rfa_detect_window))) return rc;
        RFA_CK(cudaMemcpyAsync(c->stage[5].p, win, (size_t)nwin * sizeof(rfa_detect_window), cudaMemcpyHostToDevice, c->stream));
        dwin = c->stage[5].as<rfa_detect_window>();
    }
    float *dpeak = peak, *davg = avg;
    if (mem_out == RFA_MEM_HOST) {
        if (int rc = c->stage[6].ensure((size_t)nwin * 2 * sizeof(float))) return rc;
        dpeak = c->stage[6].as<float>();
        davg = dpeak + nwin;
    }
    const unsigned blocks = (unsigned)(((long long)nwin * 32 + 255) / 256);
    detect_windows_kernel<<<blocks, 256, 0, c->stream>>>(rows, row_stride, n, dwin, nwin, dpeak, davg);
    RFA_CK(cudaGetLastError());
    c->launches++;
    if (mem_out == RFA_MEM_HOST) {
        RFA_CK(cudaMemcpyAsync(peak, dpeak, (size_t)nwin * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        RFA_CK(cudaMemcpyAsync(avg, davg, (size_t)nwin * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    if (mem_out == RFA_MEM_HOST || mem_win == RFA_MEM_HOST) RFA_CK(cudaStreamSynchronize(c->stream));  // staging is reused
    return RFA_OK;
}

int rfa_detect_bin(long long center_freq, long long sample_rate, int n, long long freq) {
    const long long start = center_freq - sample_rate / 2;
    return jvm_to_int((float)(freq - start) / resolution(sample_rate, n));
}

int rfa_detect_half_width(long long sample_rate, int n, int half_width_hz, int min_half) {
    const int h = jvm_to_int((float)half_width_hz / resolution(sample_rate, n));
    return h < min_half ? min_half : h;
}

int rfa_detect_window_at(long long center_freq, long long sample_rate, int n, long long freq, int half_width_hz,
                         int min_half, int *bin, int *start, int *end) {
    RFA_REQUIRE(n > 0 && sample_rate > 0 && bin && start && end, "rfa_detect_window_at: bad argument");
    const int b = rfa_detect_bin(center_freq, sample_rate, n, freq);
    const int h = half_width_hz < 0 ? min_half : rfa_detect_half_width(sample_rate, n, half_width_hz, min_half);
    *bin = b;
    *start = (long long)b - h < 0 ? 0 : b - h;
    *end = (long long)b + h > n - 1 ? n - 1 : b + h;
    return (b >= 0 && b < n) ? 1 : 0;  // `binIndex in currentFFT.indices`
}

int rfa_detect_decide(float peak, float avg, float threshold, float noise_floor, float margin, int mode) {
    const float nf = noise_floor + margin;
    const float eff = threshold > nf ? threshold : nf;  // maxOf(threshold, noiseFloor + noiseFloorMargin)
    switch (mode) {
        case RFA_DETECT_PEAK_ONLY: return peak > eff;
        case RFA_DETECT_AVERAGE_ONLY: return avg > eff;
        case RFA_DETECT_PEAK_OR_AVERAGE: return peak > eff || avg > eff;
    }
    set_error("unknown detection mode %d", mode);
    return -1;
}

long long rfa_scan_grid(long long center_freq, long long sample_rate, long long usable_bandwidth, long long step,
                        long long scan_start, long long scan_end, int n, long long row, long long *freqs,
                        rfa_detect_window *win, long long cap) {
    if (n <= 0 || sample_rate <= 0 || step <= 0) {
        set_error("rfa_scan_grid: bad argument");
        return -1;
    }
    const long long start = center_freq - sample_rate / 2;
    const long long usable_lo = (sample_rate - usable_bandwidth) / 2, usable_hi = usable_lo + usable_bandwidth;
    long long f = std::max(scan_start, start + usable_lo);
    const long long f_end = std::min(scan_end, start + usable_hi);
    const float res = resolution(sample_rate, n);
    long long count = 0;
    for (; f <= f_end; f += step) {
        const int b = jvm_to_int((float)(f - start) / res);
        if (b < 0 || b >= n) continue;
        if (count < cap) {
            if (freqs) freqs[count] = f;
            if (win) win[count] = rfa_detect_window{row, std::max(0, b - 2), std::min(n - 1, b + 2)};
        }
        count++;
    }
    return count;
}

long long rfa_group_signals(const rfa_signal *in, long long n, long long step, int minimum_gap, rfa_signal *out) {
    if (n <= 0) return 0;
    if (!in || !out) {
        set_error("rfa_group_signals: NULL argument");
        return -1;
    }
    std::vector<rfa_signal> s(in, in + n);
    std::stable_sort(s.begin(), s.end(), [](const rfa_signal &a, const rfa_signal &b) { return a.frequency < b.frequency; });
    const long long gap_threshold = step * (long long)minimum_gap;
    long long nout = 0;
    auto finalize = [&](long long lo, long long hi) {  // [lo, hi)
        if (hi - lo == 1) {
            out[nout++] = s[lo];
            return;
        }
        rfa_signal g{};
        long long mn = s[lo].frequency, mx = s[lo].frequency;
        float pk = s[lo].peak;
        double sum = 0.0;
        for (long long i = lo; i < hi; i++) {
            mn = std::min(mn, s[i].frequency);
            mx = std::max(mx, s[i].frequency);
            pk = (pk != pk || s[i].peak != s[i].peak) ? NAN : std::max(pk, s[i].peak);
            sum += (double)s[i].average;
        }
        g.frequency = (mn + mx) / 2;
        g.bandwidth = mx - mn;
        g.peak = pk;
        g.average = (float)(sum / (double)(hi - lo));
        g.grouped = 1;
        out[nout++] = g;
    };
    long long lo = 0;
    for (long long i = 1; i < n; i++)
        if (s[i].frequency - s[i - 1].frequency > gap_threshold) {
            finalize(lo, i);
            lo = i;
        }
    finalize(lo, n);
    return nout;
}

}  // extern "C"
