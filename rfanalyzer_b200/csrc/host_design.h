// host_design.h -- host-side design functions of the hot path (tables and taps that the
// reference computes once per configuration on the JVM).  Float/double usage follows the
// reference expression by expression so the resulting tables are bit-identical.
#pragma once
#include <vector>

namespace rfa {
namespace design {

// IQConverter.calcOptimalCosineLength (A/source/IQConverter.java:64-76)
int optimal_cosine_length(int sample_rate, int cosine_frequency);
// generateMixerLookupTable's frequency rule + oscillator tables
// (Signed8BitIQConverter.java:54-77, Unsigned8BitIQConverter.java:54-77, Signed16BitIQConverter.kt:59-87)
void nco_tables(int fmt, int sample_rate, int mix_frequency, int *effective_frequency,
                std::vector<float> *cos_t, std::vector<float> *sin_t);

// A/dsp/WindowFunctions.kt:44-100
float tap_window(int kind, double beta, int n, int N);
// FirFilter.createLowPassTaps (A/dsp/FirFilter.kt:182-241); empty vector = firdes check failed
std::vector<float> lowpass_taps(float gain, float fs, float cutoff, float tw, float att, int window_kind,
                                double beta, int max_taps);
// ComplexFirFilter.createBandPass (A/dsp/ComplexFirFilter.java:186-262)
bool bandpass_taps(float gain, float fs, float lo, float hi, float tw, float att, std::vector<float> *re,
                   std::vector<float> *im);
// RationalResampler.gcd / limitDenominator / designResamplerTaps (A/dsp/RationalResampler.kt:161-255)
int gcd(int a, int b);
void limit_denominator(int num, int den, int max_den, int *out_num, int *out_den);
std::vector<float> resampler_taps(int interpolation, int decimation, float fractional_bw, int max_taps);
// FftProcessor.kt:143-149 channel bin range
void channel_bins(int n, long long frequency, int sample_rate, long long chan_start, long long chan_end,
                  int *b0, int *b1);

int java_d2i(double v);  // (int) cast of the JVM: truncating, saturating, NaN -> 0

}  // namespace design
}  // namespace rfa
