// host_mirror_test.cpp -- the C++ host mirror (include/rfa_b200.hpp) running the reference's own
// FIR known-answer test (ApplicationTest.kt:20-127) and a converter/NativeDsp round through the C ABI.
// Exit code 0 = pass; 77 = no GPU (the no-GPU run still proves the mirror compiles, links and fails loudly).
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/rfa_b200.hpp"

int main() {
    // host-only: design functions work without a device
    auto taps = rfa::FirFilter::createLowPassTaps(4, 1.f, 1000.f, 100.f, 50.f, 60.f);
    if (taps.size() != 55) return 1;
    if (!rfa::FirFilter::createLowPassTaps(1, 1.f, 1000.f, 600.f, 50.f, 60.f).empty()) return 2;  // firdes check
    auto id = rfa::RationalResampler::limitDenominator(384000, 2400000);
    if (id.first != 4 || id.second != 25) return 3;
    try {
        rfa::Context ctx(0);
        const int n = 128, sr = 1000;
        std::vector<float> re(n), im(n);
        for (int i = 0; i < n; i++) {
            re[i] = (float)std::cos(2 * M_PI * 50 * i / (float)sr) + (float)std::cos(2 * M_PI * 200 * i / (float)sr);
            im[i] = (float)std::sin(2 * M_PI * 50 * i / (float)sr) + (float)std::sin(2 * M_PI * 200 * i / (float)sr);
        }
        rfa::SamplePacket in(re, im, 0, sr), out(n / 4);
        auto f = rfa::FirFilter::createLowPass(ctx, 4, 1.f, (float)sr, 100.f, 50.f, 60.f);
        if (f->filter(in, out, 0, in.size()) != n || out.size() != 32) return 4;
        const float want_re[4] = {1.7833801E-4f, 5.6521903E-4f, -0.008516869f, 0.028878199f};  // ApplicationTest.kt:55-58
        for (int i = 0; i < 4; i++)
            if (out.re(i) != want_re[i]) return 5;
        if (std::fabs(out.re(31) - 0.99987644f) > 1e-9 || std::fabs(out.im(31) - 1.675597E-8f) > 1e-9) return 6;
        // converter + NativeDsp
        rfa::Signed8BitIQConverter conv(ctx);
        conv.setSampleRate(20000000);
        conv.setFrequency(100000000);
        std::vector<uint8_t> bytes(2 * 1024);
        for (size_t i = 0; i < bytes.size(); i++) bytes[i] = (uint8_t)(i * 37);
        rfa::SamplePacket sp(1024);
        if (conv.fillPacketIntoSamplePacket(bytes.data(), (int)bytes.size(), sp) != 1024) return 7;
        if (sp.re(1) != (float)((int8_t)bytes[2]) / 128.0f) return 8;
        rfa::NativeDsp dsp(ctx);
        std::vector<float> r(sp.re(), sp.re() + 1024), q(sp.im(), sp.im() + 1024), mag(1024), bad(1000);
        if (!dsp.performWindowedFftAndReturnMag(r, q, mag)) return 9;
        if (dsp.performWindowedFftAndReturnMag(r, q, bad)) return 10;  // size mismatch -> false (NativeDsp.kt:44-46)
        printf("host mirror ok\n");
        return 0;
    } catch (const rfa::Error &e) {
        if (e.code == RFA_ERR_CUDA && std::strstr(e.what(), "no CPU fallback")) {
            printf("no GPU: %s\n", e.what());
            return 77;
        }
        fprintf(stderr, "unexpected: %s\n", e.what());
        return 20;
    }
}
