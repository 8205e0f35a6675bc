// rfa_b200.hpp -- C++ host-side mirror of the reference's JVM classes on the hot path, header-only
// over the C ABI (rfa_b200.h).  The reference is Kotlin/Java and no JVM toolchain exists in this
// image, so the host side above the C ABI is C++: same class and method names, argument meaning,
// return values and error behaviour as the reference, so a Kotlin/Java caller maps 1:1 (see
// INTEGRATION.md for the JNI glue a maintainer adds per class).
//
//   rfa::SamplePacket            A/source/SamplePacket.java:28-137
//   rfa::IQConverter + 3         A/source/IQConverter.java:31-87 and subclasses
//   rfa::NativeDsp               nativedsp/src/main/java/com/mantz_it/nativedsp/NativeDsp.kt:10-63
//   rfa::FirFilter               A/dsp/FirFilter.kt:34-263
//   rfa::ComplexFirFilter        A/dsp/ComplexFirFilter.java:33-279
//   rfa::RationalResampler       A/dsp/RationalResampler.kt:36-257
// (A/ = app/src/main/java/com/mantz_it/rfanalyzer/).  Packets are host memory, like float[].
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "rfa_b200.h"

namespace rfa {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc != RFA_OK) throw Error(rc, rfa_last_error());
}

class Context {
public:
    explicit Context(int device = 0, void *stream = nullptr) { check(rfa_ctx_create(device, stream, &ctx_)); }
    ~Context() { rfa_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    rfa_ctx *get() const { return ctx_; }
    void sync() { check(rfa_ctx_sync(ctx_)); }

private:
    rfa_ctx *ctx_ = nullptr;
};

class SamplePacket {
public:
    explicit SamplePacket(int size) : re_(size, 0.f), im_(size, 0.f) {}
    SamplePacket(std::vector<float> re, std::vector<float> im, long long frequency, int sampleRate)
        : re_(std::move(re)), im_(std::move(im)), frequency_(frequency), sampleRate_(sampleRate) {
        if (re_.size() != im_.size()) throw std::invalid_argument("Arrays must be of the same length");
        size_ = (int)re_.size();
    }
    float *re() { return re_.data(); }
    float *im() { return im_.data(); }
    const float *re() const { return re_.data(); }
    const float *im() const { return im_.data(); }
    float re(int i) const { return re_[i]; }
    float im(int i) const { return im_[i]; }
    int capacity() const { return (int)re_.size(); }
    int size() const { return size_; }
    void setSize(int size) { size_ = size < capacity() ? size : capacity(); }
    long long getFrequency() const { return frequency_; }
    int getSampleRate() const { return sampleRate_; }
    void setFrequency(long long f) { frequency_ = f; }
    void setSampleRate(int r) { sampleRate_ = r; }

private:
    std::vector<float> re_, im_;
    long long frequency_ = 0;
    int sampleRate_ = 0;
    int size_ = 0;
};

class IQConverter {
public:
    virtual ~IQConverter() = default;
    long long getFrequency() const { return frequency_; }
    void setFrequency(long long f) { frequency_ = f; }
    int getSampleRate() const { return sampleRate_; }
    void setSampleRate(int r) {
        if (sampleRate_ != r) {
            sampleRate_ = r;
            cosineFrequency_ = -1;  // forces a new mixer table (IQConverter.java:57-62)
        }
    }
    // returns the number of samples appended to the packet
    int fillPacketIntoSamplePacket(const uint8_t *packet, int packetLength, SamplePacket &sp) {
        const int start = sp.size();
        const int count = room(packetLength, sp);
        if (count == 0) return 0;
        check(rfa_convert(ctx_.get(), format(), packet, count, sp.re() + start, sp.im() + start, RFA_MEM_HOST));
        sp.setSize(start + count);
        sp.setSampleRate(sampleRate_);
        sp.setFrequency(frequency_);
        return count;
    }
    int mixPacketIntoSamplePacket(const uint8_t *packet, int packetLength, SamplePacket &sp, long long channelFrequency) {
        generateMixerLookupTable((int)(frequency_ - channelFrequency));
        const int start = sp.size();
        const int count = room(packetLength, sp);
        if (count == 0 || cos_.empty()) return 0;
        if (cosineIndex_ >= (int)cos_.size()) cosineIndex_ = 0;
        check(rfa_mix(ctx_.get(), format(), packet, count, cos_.data(), sin_.data(), (int)cos_.size(), cosineIndex_,
                      sp.re() + start, sp.im() + start, RFA_MEM_HOST));
        cosineIndex_ = (int)((cosineIndex_ + (long long)count) % (long long)cos_.size());
        sp.setSize(start + count);
        sp.setSampleRate(sampleRate_);
        sp.setFrequency(channelFrequency);
        return count;
    }

protected:
    explicit IQConverter(Context &ctx) : ctx_(ctx) {}
    virtual int format() const = 0;
    void generateMixerLookupTable(int mixFrequency) {
        const int amix = mixFrequency < 0 ? -mixFrequency : mixFrequency;
        if (mixFrequency == 0 || sampleRate_ / amix > 500) mixFrequency += sampleRate_;
        if (!cos_.empty() && mixFrequency == cosineFrequency_) return;
        std::vector<float> c(500), s(500);
        int eff = 0, len = 0;
        check(rfa_nco_design(format(), sampleRate_, mixFrequency, &eff, &len, c.data(), s.data()));
        c.resize(len);
        s.resize(len);
        cos_ = std::move(c);
        sin_ = std::move(s);
        cosineFrequency_ = mixFrequency;
        cosineIndex_ = 0;
    }
    int room(int packetLength, const SamplePacket &sp) const {
        if (sp.size() >= sp.capacity()) return 0;
        const int avail = packetLength / (format() == RFA_FMT_S16LE ? 4 : 2);
        const int free_ = sp.capacity() - sp.size();
        return avail < free_ ? avail : free_;
    }
    Context &ctx_;
    long long frequency_ = 0;
    int sampleRate_ = 0;
    int cosineFrequency_ = 0, cosineIndex_ = 0;
    std::vector<float> cos_, sin_;
};
class Signed8BitIQConverter : public IQConverter {
public:
    explicit Signed8BitIQConverter(Context &c) : IQConverter(c) {}
    int format() const override { return RFA_FMT_S8; }
};
class Unsigned8BitIQConverter : public IQConverter {
public:
    explicit Unsigned8BitIQConverter(Context &c) : IQConverter(c) {}
    int format() const override { return RFA_FMT_U8; }
};
class Signed16BitIQConverter : public IQConverter {
public:
    explicit Signed16BitIQConverter(Context &c) : IQConverter(c) {}
    int format() const override { return RFA_FMT_S16LE; }
};

class NativeDsp {
public:
    explicit NativeDsp(Context &c) : ctx_(c) {}
    // false when the array sizes do not match (NativeDsp.kt:44-46); callers may ignore it
    bool performWindowedFftAndReturnMag(const std::vector<float> &re, const std::vector<float> &im,
                                        std::vector<float> &magOut) {
        const size_t n = re.size();
        if (im.size() != n || magOut.size() != n) return false;
        check(rfa_windowed_fft_logmag(ctx_.get(), re.data(), im.data(), magOut.data(), (int)n, 1, RFA_WIN_BLACKMAN_REF,
                                      RFA_MEM_HOST));
        return true;
    }

private:
    Context &ctx_;
};

class FirFilter {
public:
    FirFilter(Context &c, std::vector<float> taps, int decimation, float gain = 1.f, float sampleRate = 1.f,
              float cutOffFrequency = 0.f, float transitionWidth = 0.f, float attenuation = 0.f,
              int flags = RFA_SUM_EXACT)
        : taps(std::move(taps)), decimation(decimation), gain(gain), sampleRate(sampleRate),
          cutOffFrequency(cutOffFrequency), transitionWidth(transitionWidth), attenuation(attenuation) {
        check(rfa_fir_create(c.get(), this->taps.data(), nullptr, (int)this->taps.size(), decimation, flags, &f_));
    }
    ~FirFilter() { rfa_fir_destroy(f_); }
    FirFilter(const FirFilter &) = delete;
    int numberOfTaps() const { return (int)taps.size(); }
    // empty result = firdes check failed (the reference returns null)
    static std::vector<float> createLowPassTaps(int, float gain, float sampleRate, float cutoffFrequency,
                                                float transitionWidth, float attenuationInDecibels,
                                                int window = RFA_TAPWIN_BLACKMAN, double beta = 0.0, int maxTaps = 0) {
        int n = 0;
        if (rfa_design_lowpass(gain, sampleRate, cutoffFrequency, transitionWidth, attenuationInDecibels, window, beta,
                               maxTaps, nullptr, 0, &n) != RFA_OK)
            return {};
        std::vector<float> t(n);
        check(rfa_design_lowpass(gain, sampleRate, cutoffFrequency, transitionWidth, attenuationInDecibels, window, beta,
                                 maxTaps, t.data(), n, &n));
        return t;
    }
    static std::unique_ptr<FirFilter> createLowPass(Context &c, int decimation, float gain, float sampleRate,
                                                    float cutoffFrequency, float transitionWidth,
                                                    float attenuationInDecibels) {
        auto t = createLowPassTaps(decimation, gain, sampleRate, cutoffFrequency, transitionWidth, attenuationInDecibels);
        if (t.empty()) return nullptr;
        return std::make_unique<FirFilter>(c, std::move(t), decimation, gain, sampleRate, cutoffFrequency,
                                           transitionWidth, attenuationInDecibels);
    }
    // returns the number of samples consumed from the input packet
    int filter(const SamplePacket &in, SamplePacket &out, int offset, int length) { return run(in, out, offset, length, false); }
    int filterReal(const SamplePacket &in, SamplePacket &out, int offset, int length) { return run(in, out, offset, length, true); }

    const std::vector<float> taps;
    const int decimation;
    const float gain, sampleRate, cutOffFrequency, transitionWidth, attenuation;

private:
    int run(const SamplePacket &in, SamplePacket &out, int offset, int length, bool real) {
        const int start = out.size();
        long long nout = 0, consumed = 0;
        check(rfa_fir_process(f_, in.re() + offset, real ? nullptr : in.im() + offset, length, out.re() + start,
                              real ? nullptr : out.im() + start, out.capacity() - start, &nout, &consumed, RFA_MEM_HOST));
        out.setSize(start + (int)nout);
        out.setSampleRate(in.getSampleRate() / decimation);
        return (int)consumed;
    }
    rfa_fir *f_ = nullptr;
};

class ComplexFirFilter {
public:
    ~ComplexFirFilter() { rfa_fir_destroy(f_); }
    static std::unique_ptr<ComplexFirFilter> createBandPass(Context &c, int decimation, float gain, float sampling_freq,
                                                            float low_cutoff_freq, float high_cutoff_freq,
                                                            float transition_width, float attenuation_dB,
                                                            int flags = RFA_SUM_EXACT) {
        int n = 0;
        if (rfa_design_bandpass(gain, sampling_freq, low_cutoff_freq, high_cutoff_freq, transition_width, attenuation_dB,
                                nullptr, nullptr, 0, &n) != RFA_OK)
            return nullptr;
        std::unique_ptr<ComplexFirFilter> f(new ComplexFirFilter());
        f->tapsReal.resize(n);
        f->tapsImag.resize(n);
        check(rfa_design_bandpass(gain, sampling_freq, low_cutoff_freq, high_cutoff_freq, transition_width, attenuation_dB,
                                  f->tapsReal.data(), f->tapsImag.data(), n, &n));
        f->decimation = decimation;
        f->lowCutOffFrequency = low_cutoff_freq;
        f->highCutOffFrequency = high_cutoff_freq;
        check(rfa_fir_create(c.get(), f->tapsReal.data(), f->tapsImag.data(), n, decimation, flags, &f->f_));
        return f;
    }
    int getNumberOfTaps() const { return (int)tapsReal.size(); }
    int getDecimation() const { return decimation; }
    float getLowCutOffFrequency() const { return lowCutOffFrequency; }
    float getHighCutOffFrequency() const { return highCutOffFrequency; }
    int filter(const SamplePacket &in, SamplePacket &out, int offset, int length) {
        const int start = out.size();
        long long nout = 0, consumed = 0;
        check(rfa_fir_process(f_, in.re() + offset, in.im() + offset, length, out.re() + start, out.im() + start,
                              out.capacity() - start, &nout, &consumed, RFA_MEM_HOST));
        out.setSize(start + (int)nout);
        out.setSampleRate(in.getSampleRate() / decimation);
        return (int)consumed;
    }

private:
    ComplexFirFilter() = default;
    std::vector<float> tapsReal, tapsImag;
    int decimation = 1;
    float lowCutOffFrequency = 0, highCutOffFrequency = 0;
    rfa_fir *f_ = nullptr;
};

class RationalResampler {
public:
    RationalResampler(Context &c, int interpolation, int decimation, const std::vector<float> *taps = nullptr,
                      float fractionalBw = 0.4f, int maxTaps = 0, int flags = RFA_SUM_EXACT) {
        if (interpolation <= 0) throw std::invalid_argument("Interpolation must be > 0");
        if (decimation <= 0) throw std::invalid_argument("Decimation must be > 0");
        check(rfa_resampler_create(c.get(), interpolation, decimation, taps ? taps->data() : nullptr,
                                   taps ? (int)taps->size() : 0, fractionalBw, maxTaps, flags, &r_));
        check(rfa_resampler_info(r_, &interpolation_, &decimation_, &tapsPerPhase_));
    }
    ~RationalResampler() { rfa_resampler_destroy(r_); }
    RationalResampler(const RationalResampler &) = delete;
    int getInterpolation() const { return interpolation_; }
    int getDecimation() const { return decimation_; }
    static std::pair<int, int> limitDenominator(int numerator, int denominator, int maxDenominator = 10000) {
        int a = 0, b = 0;
        check(rfa_limit_denominator(numerator, denominator, maxDenominator, &a, &b));
        return {a, b};
    }
    static std::vector<float> designResamplerTaps(int interpolation, int decimation, float fractionalBw, int maxTaps) {
        int n = 0;
        check(rfa_design_resampler_taps(interpolation, decimation, fractionalBw, maxTaps, nullptr, 0, &n));
        std::vector<float> t(n);
        check(rfa_design_resampler_taps(interpolation, decimation, fractionalBw, maxTaps, t.data(), n, &n));
        return t;
    }
    // returns the number of consumed input samples
    int resample(const SamplePacket &in, SamplePacket &out, int offset, int length) {
        const int start = out.size();
        long long nout = 0, consumed = 0;
        check(rfa_resampler_process(r_, in.re() + offset, in.im() + offset, length, out.re() + start, out.im() + start,
                                    out.capacity() - start, &nout, &consumed, RFA_MEM_HOST));
        out.setSize(start + (int)nout);
        out.setSampleRate((int)((long long)in.getSampleRate() * interpolation_ / decimation_));
        out.setFrequency(in.getFrequency());
        return (int)consumed;
    }

private:
    rfa_resampler *r_ = nullptr;
    int interpolation_ = 0, decimation_ = 0, tapsPerPhase_ = 0;
};

}  // namespace rfa
