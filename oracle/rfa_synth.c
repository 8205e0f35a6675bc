/*
 * rfa_synth.c -- deterministic synthetic IQ generator (TEST INFRASTRUCTURE ONLY).
 * The reference ships no input fixtures beyond inline sinusoids
 * (ApplicationTest.kt:33-38); this is the builder's generator from SURVEY.md 8(d),
 * restated with integer tables so the device-side generator matches it bit for bit.
 */
#include "rfa_oracle.h"

#include <math.h>
#include <stdlib.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static int16_t g_tab[4096];
static int g_tab_ready = 0;

static void make_tab(void) {
    if (g_tab_ready) return;
    for (int j = 0; j < 4096; j++) g_tab[j] = (int16_t)lround(16384.0 * cos(2.0 * M_PI * j / 4096.0));
    g_tab_ready = 1;
}

static uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

uint32_t orc_synth_step(double cyclesPerSample) {
    double f = cyclesPerSample - floor(cyclesPerSample);
    return (uint32_t)llround(f * 4294967296.0);
}

int orc_synth_default_comps(int fmt, orc_synth_comp *c) {
    int mul = (fmt == ORC_FMT_S16LE) ? 256 : 1;
    c[0] = (orc_synth_comp){orc_synth_step(0.1234), 48 * mul, 0, 0};
    c[1] = (orc_synth_comp){orc_synth_step(-0.3071), 24 * mul, 0, 0};
    c[2] = (orc_synth_comp){orc_synth_step(0.0127), 12 * mul, 0, 0};
    return 3;
}

void orc_synth_iq(int fmt, uint32_t seed, const orc_synth_comp *comps, int ncomp, int noiseShift,
                  long long firstSample, long long nsamples, uint8_t *out) {
    make_tab();
    for (long long k = 0; k < nsamples; k++) {
        uint64_t n = (uint64_t)(firstSample + k);
        uint32_t h = fmix32(seed ^ (uint32_t)n ^ ((uint32_t)(n >> 32) * 0x9E3779B9u));
        int32_t vi, vq;
        if (fmt == ORC_FMT_S16LE) {
            vi = ((int32_t)(int16_t)(h & 0xFFFF)) >> noiseShift;
            vq = ((int32_t)(int16_t)(h >> 16)) >> noiseShift;
        } else {
            vi = ((int32_t)(int8_t)(h & 0xFF)) >> noiseShift;
            vq = ((int32_t)(int8_t)((h >> 8) & 0xFF)) >> noiseShift;
        }
        for (int c = 0; c < ncomp; c++) {
            uint32_t ph = (uint32_t)(n * comps[c].step);
            if (comps[c].modK != 0) {
                uint32_t mph = (uint32_t)(n * comps[c].modStep);
                /* sine of the modulating phase */
                int32_t m = g_tab[((mph - 0x40000000u) >> 20) & 4095];
                ph += (uint32_t)((int64_t)comps[c].modK * (int64_t)m);
            }
            int32_t ci = g_tab[(ph >> 20) & 4095];
            int32_t si = g_tab[((ph - 0x40000000u) >> 20) & 4095];
            vi += (comps[c].amp * ci + 8192) >> 14;
            vq += (comps[c].amp * si + 8192) >> 14;
        }
        if (fmt == ORC_FMT_S16LE) {
            if (vi > 32767) vi = 32767;
            if (vi < -32768) vi = -32768;
            if (vq > 32767) vq = 32767;
            if (vq < -32768) vq = -32768;
            uint16_t ui = (uint16_t)(int16_t)vi, uq = (uint16_t)(int16_t)vq;
            out[4 * k + 0] = (uint8_t)(ui & 0xFF);
            out[4 * k + 1] = (uint8_t)(ui >> 8);
            out[4 * k + 2] = (uint8_t)(uq & 0xFF);
            out[4 * k + 3] = (uint8_t)(uq >> 8);
        } else {
            if (vi > 127) vi = 127;
            if (vi < -128) vi = -128;
            if (vq > 127) vq = 127;
            if (vq < -128) vq = -128;
            int bias = (fmt == ORC_FMT_U8) ? 128 : 0;
            out[2 * k + 0] = (uint8_t)(vi + bias);
            out[2 * k + 1] = (uint8_t)(vq + bias);
        }
    }
}
