/*
 * rfa_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see rfa_oracle.h).
 *
 * Plain-C restatement of the reference's JVM DSP.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared
 * so every float product/sum is rounded exactly once, like the JVM does.
 * "A/" below = /root/reference/app/src/main/java/com/mantz_it/rfanalyzer/.
 */
#include "rfa_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* Java's (int) / Kotlin's toInt() on a double: truncation toward zero, saturating, NaN -> 0 */
static int j2i(double v) {
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (-2147483647 - 1);
    return (int)v;
}

/* ======================================================================== */
/* SamplePacket  (A/source/SamplePacket.java:28-137)                         */
/* ======================================================================== */
orc_packet *orc_packet_new(int capacity) {
    orc_packet *p = (orc_packet *)calloc(1, sizeof(*p));
    p->re = (float *)calloc((size_t)(capacity > 0 ? capacity : 1), sizeof(float));
    p->im = (float *)calloc((size_t)(capacity > 0 ? capacity : 1), sizeof(float));
    p->capacity = capacity;
    return p;
}
void orc_packet_free(orc_packet *p) {
    if (!p) return;
    free(p->re);
    free(p->im);
    free(p);
}
void orc_packet_set_size(orc_packet *p, int size) { p->size = size < p->capacity ? size : p->capacity; }

/* ======================================================================== */
/* IQConverter family                                                        */
/*   A/source/IQConverter.java:31-87                                         */
/*   A/source/Signed8BitIQConverter.java:37-131                              */
/*   A/source/Unsigned8BitIQConverter.java:37-131                            */
/*   A/source/Signed16BitIQConverter.kt:46-181                               */
/* ======================================================================== */
struct orc_converter {
    int fmt;
    long long frequency;
    int sampleRate;
    float *lut;
    int lutLen;
    int haveMixer;       /* cosine tables non-null */
    int cosineFrequency; /* Java field default 0; setSampleRate() invalidates with -1 */
    int cosineIndex;
    int cosLen;
    float *cosT, *sinT; /* per-t oscillator values (8-bit: the cosineAtT / sineAtT scalars) */
    float *cosRe2d, *cosIm2d; /* 8-bit only: [t][256] product tables */
};

#define MAX_COSINE_LENGTH 500 /* IQConverter.java:42 */

static void conv_make_lut(orc_converter *c) {
    if (c->fmt == ORC_FMT_S8) { /* Signed8BitIQConverter.java:48-50 */
        c->lutLen = 256;
        c->lut = (float *)malloc(256 * sizeof(float));
        for (int i = 0; i < 256; i++) c->lut[i] = (i - 128) / 128.0f;
    } else if (c->fmt == ORC_FMT_U8) { /* Unsigned8BitIQConverter.java:48-50 */
        c->lutLen = 256;
        c->lut = (float *)malloc(256 * sizeof(float));
        for (int i = 0; i < 256; i++) c->lut[i] = (i - 127.4f) / 128.0f;
    } else { /* Signed16BitIQConverter.kt:46-57 */
        c->lutLen = 65536;
        c->lut = (float *)malloc(65536 * sizeof(float));
        for (int u = 0; u < 65536; u++) {
            int s = (int)(int16_t)(uint16_t)u;
            c->lut[u] = s / 32768.0f;
        }
    }
}

orc_converter *orc_converter_new(int fmt) {
    orc_converter *c = (orc_converter *)calloc(1, sizeof(*c));
    c->fmt = fmt;
    conv_make_lut(c);
    return c;
}
void orc_converter_free(orc_converter *c) {
    if (!c) return;
    free(c->lut);
    free(c->cosT);
    free(c->sinT);
    free(c->cosRe2d);
    free(c->cosIm2d);
    free(c);
}
void orc_converter_set_frequency(orc_converter *c, long long f) { c->frequency = f; }
void orc_converter_set_sample_rate(orc_converter *c, int fs) { /* IQConverter.java:57-62 */
    if (c->sampleRate != fs) {
        c->sampleRate = fs;
        c->cosineFrequency = -1;
    }
}
const float *orc_converter_lut(const orc_converter *c, int *n) {
    if (n) *n = c->lutLen;
    return c->lut;
}

/* IQConverter.java:64-76 */
int orc_calc_optimal_cosine_length(int sampleRate, int cosineFrequency) {
    double cycleLength = sampleRate / fabs((double)cosineFrequency);
    int bestLength = j2i(cycleLength);
    double bestLengthError = fabs(bestLength - cycleLength);
    for (int i = 1; i * cycleLength < MAX_COSINE_LENGTH; i++) {
        if (fabs(i * cycleLength - j2i(i * cycleLength)) < bestLengthError) {
            bestLength = j2i(i * cycleLength);
            bestLengthError = fabs(bestLength - (i * cycleLength));
        }
    }
    return bestLength;
}

/* Signed8BitIQConverter.java:54-77, Unsigned8BitIQConverter.java:54-77,
 * Signed16BitIQConverter.kt:59-87 */
static void conv_make_mixer(orc_converter *c, int mixFrequency) {
    int absMix = mixFrequency < 0 ? -mixFrequency : mixFrequency;
    if (mixFrequency == 0 || (c->sampleRate / absMix > MAX_COSINE_LENGTH)) mixFrequency += c->sampleRate;
    if (c->haveMixer && mixFrequency == c->cosineFrequency) return;

    c->cosineFrequency = mixFrequency;
    int bestLength = orc_calc_optimal_cosine_length(c->sampleRate, c->cosineFrequency);
    if (bestLength < 0) bestLength = 0;
    free(c->cosT);
    free(c->sinT);
    free(c->cosRe2d);
    free(c->cosIm2d);
    c->cosRe2d = c->cosIm2d = NULL;
    c->cosLen = bestLength;
    c->cosT = (float *)malloc(sizeof(float) * (size_t)(bestLength + 1));
    c->sinT = (float *)malloc(sizeof(float) * (size_t)(bestLength + 1));
    if (c->fmt == ORC_FMT_S16LE) {
        /* angle = ((2*pi*f)/fs) * t, all double (Signed16BitIQConverter.kt:73-79) */
        double twoPiFOverFs = (2.0 * M_PI * c->cosineFrequency) / (double)c->sampleRate;
        for (int t = 0; t < bestLength; t++) {
            double angle = twoPiFOverFs * t;
            c->cosT[t] = (float)cos(angle);
            c->sinT[t] = (float)sin(angle);
        }
    } else {
        /* angle = 2*pi*f*t / (float)fs, left-to-right in double
         * (Signed8BitIQConverter.java:68-69) */
        c->cosRe2d = (float *)malloc(sizeof(float) * 256 * (size_t)(bestLength + 1));
        c->cosIm2d = (float *)malloc(sizeof(float) * 256 * (size_t)(bestLength + 1));
        for (int t = 0; t < bestLength; t++) {
            double fsd = (double)(float)c->sampleRate;
            float cosineAtT = (float)cos(2 * M_PI * c->cosineFrequency * t / fsd);
            float sineAtT = (float)sin(2 * M_PI * c->cosineFrequency * t / fsd);
            c->cosT[t] = cosineAtT;
            c->sinT[t] = sineAtT;
            for (int i = 0; i < 256; i++) {
                /* value expression repeated, not the LUT (…java:71-72) */
                float v = (c->fmt == ORC_FMT_S8) ? (i - 128) / 128.0f : (i - 127.4f) / 128.0f;
                c->cosRe2d[t * 256 + i] = v * cosineAtT;
                c->cosIm2d[t * 256 + i] = v * sineAtT;
            }
        }
    }
    c->haveMixer = 1;
    c->cosineIndex = 0;
}

int orc_converter_nco_len(const orc_converter *c) { return c->haveMixer ? c->cosLen : 0; }
int orc_converter_nco_index(const orc_converter *c) { return c->cosineIndex; }
int orc_converter_nco_freq(const orc_converter *c) { return c->cosineFrequency; }
void orc_converter_nco_table(const orc_converter *c, float *cosT, float *sinT) {
    for (int t = 0; t < c->cosLen; t++) {
        cosT[t] = c->cosT[t];
        sinT[t] = c->sinT[t];
    }
}

int orc_converter_fill(orc_converter *c, const uint8_t *packet, int nbytes, orc_packet *sp) {
    int capacity = sp->capacity;
    int startIndex = sp->size;
    int count = 0;
    if (startIndex >= capacity) return 0;
    if (c->fmt == ORC_FMT_S16LE) { /* Signed16BitIQConverter.kt:89-124 */
        int i = 0, outIdx = startIndex;
        while (i + 3 < nbytes && outIdx < capacity) {
            int iU = (packet[i] & 0xFF) | ((int)(int8_t)packet[i + 1] * 256);
            int qU = (packet[i + 2] & 0xFF) | ((int)(int8_t)packet[i + 3] * 256);
            sp->re[outIdx] = c->lut[iU & 0xFFFF];
            sp->im[outIdx] = c->lut[qU & 0xFFFF];
            i += 4;
            outIdx++;
            count++;
        }
        if (count == 0) return 0;
        orc_packet_set_size(sp, startIndex + count);
    } else { /* Signed8BitIQConverter.java:80-99 / Unsigned8BitIQConverter.java:80-99 */
        for (int i = 0; i < nbytes; i += 2) {
            int a, b;
            if (c->fmt == ORC_FMT_S8) {
                a = (int)(int8_t)packet[i] + 128;
                b = (int)(int8_t)packet[i + 1] + 128;
            } else {
                a = packet[i] & 0xff;
                b = packet[i + 1] & 0xff;
            }
            sp->re[startIndex + count] = c->lut[a];
            sp->im[startIndex + count] = c->lut[b];
            count++;
            if (startIndex + count >= capacity) break;
        }
        orc_packet_set_size(sp, sp->size + count);
    }
    sp->sampleRate = c->sampleRate;
    sp->frequency = c->frequency;
    return count;
}

int orc_converter_mix(orc_converter *c, const uint8_t *packet, int nbytes, orc_packet *sp,
                      long long channelFrequency) {
    int mixFrequency = (int)(c->frequency - channelFrequency); /* Java (int) of long: low 32 bits */
    conv_make_mixer(c, mixFrequency);
    int capacity = sp->capacity;
    int startIndex = sp->size;
    int count = 0;
    if (startIndex >= capacity) return 0;
    if (c->cosLen == 0) return 0;
    if (c->fmt == ORC_FMT_S16LE) { /* Signed16BitIQConverter.kt:126-181 */
        int cIdx = c->cosineIndex;
        int cLen = c->cosLen;
        if (cIdx >= cLen) cIdx = 0;
        int i = 0, outIdx = startIndex;
        while (i + 3 < nbytes && outIdx < capacity) {
            int iU = (packet[i] & 0xFF) | ((int)(int8_t)packet[i + 1] * 256);
            int qU = (packet[i + 2] & 0xFF) | ((int)(int8_t)packet[i + 3] * 256);
            float iF = c->lut[iU & 0xFFFF];
            float qF = c->lut[qU & 0xFFFF];
            float cs = c->cosT[cIdx];
            float sn = c->sinT[cIdx];
            float a = iF * cs, b = qF * sn, d = qF * cs, e = iF * sn;
            sp->re[outIdx] = a - b;
            sp->im[outIdx] = d + e;
            cIdx++;
            if (cIdx == cLen) cIdx = 0;
            i += 4;
            outIdx++;
            count++;
        }
        if (count == 0) return 0;
        c->cosineIndex = cIdx;
        orc_packet_set_size(sp, startIndex + count);
    } else { /* Signed8BitIQConverter.java:102-131 / Unsigned8BitIQConverter.java:102-131 */
        if (c->cosineIndex >= c->cosLen) c->cosineIndex = 0;
        for (int i = 0; i < nbytes; i += 2) {
            int a, b;
            if (c->fmt == ORC_FMT_S8) {
                a = (int)(int8_t)packet[i] + 128;
                b = (int)(int8_t)packet[i + 1] + 128;
            } else {
                a = packet[i] & 0xff;
                b = packet[i + 1] & 0xff;
            }
            const float *cr = c->cosRe2d + (size_t)c->cosineIndex * 256;
            const float *ci = c->cosIm2d + (size_t)c->cosineIndex * 256;
            sp->re[startIndex + count] = cr[a] - ci[b];
            sp->im[startIndex + count] = cr[b] + ci[a];
            c->cosineIndex = (c->cosineIndex + 1) % c->cosLen;
            count++;
            if (startIndex + count >= capacity) break;
        }
        orc_packet_set_size(sp, sp->size + count);
    }
    sp->sampleRate = c->sampleRate;
    sp->frequency = channelFrequency;
    return count;
}

/* ======================================================================== */
/* NativeDsp window, FFT, log magnitude                                      */
/*   nativedsp/src/main/java/com/mantz_it/nativedsp/NativeDsp.kt:14-21,43-62 */
/*   nativedsp/src/main/cpp/nativedsp.cpp:19-81                              */
/* ======================================================================== */
void orc_nativedsp_window(int N, float *w) { /* NativeDsp.kt:14-21 */
    for (int i = 0; i < N; i++)
        w[i] = (float)(0.42 - 0.5 * cos(2 * M_PI * i / (N - 1)) + 0.08 * cos(4 * M_PI * i / (N - 1)));
}

static int ilog2(int n) {
    int l = 0;
    while ((1 << l) < n) l++;
    return l;
}

/* The reference calls pffft_transform_ordered (pffft.c:1904), a float32 mixed-radix
 * FFT.  pffft's exact rounding sequence is not restated; this is a plain float32
 * radix-2 decimation-in-time transform with double-computed twiddles, whose error
 * vs the exact DFT is of the same order as pffft's (both ~1e-7 * sqrt(log N) of the
 * frame's rms spectrum).  oracle/_ref (the compiled reference) pins it. */
void orc_fft_c2c_f32(const float *in, float *out, int N) {
    int lg = ilog2(N);
    for (int i = 0; i < N; i++) {
        unsigned r = 0;
        for (int b = 0; b < lg; b++)
            if (i & (1 << b)) r |= 1u << (lg - 1 - b);
        out[2 * r] = in[2 * i];
        out[2 * r + 1] = in[2 * i + 1];
    }
    float *twr = (float *)malloc(sizeof(float) * (size_t)(N / 2 + 1));
    float *twi = (float *)malloc(sizeof(float) * (size_t)(N / 2 + 1));
    for (int k = 0; k < N / 2; k++) {
        twr[k] = (float)cos(-2.0 * M_PI * k / N);
        twi[k] = (float)sin(-2.0 * M_PI * k / N);
    }
    for (int len = 2; len <= N; len <<= 1) {
        int half = len >> 1, step = N / len;
        for (int base = 0; base < N; base += len)
            for (int j = 0; j < half; j++) {
                float wr = twr[j * step], wi = twi[j * step];
                float *a = out + 2 * (base + j), *b = out + 2 * (base + j + half);
                float tr = b[0] * wr - b[1] * wi;
                float ti = b[0] * wi + b[1] * wr;
                b[0] = a[0] - tr;
                b[1] = a[1] - ti;
                a[0] = a[0] + tr;
                a[1] = a[1] + ti;
            }
    }
    free(twr);
    free(twi);
}

void orc_fft_c2c_f64(const float *in, double *out, int N) {
    int lg = ilog2(N);
    for (int i = 0; i < N; i++) {
        unsigned r = 0;
        for (int b = 0; b < lg; b++)
            if (i & (1 << b)) r |= 1u << (lg - 1 - b);
        out[2 * r] = in[2 * i];
        out[2 * r + 1] = in[2 * i + 1];
    }
    for (int len = 2; len <= N; len <<= 1) {
        int half = len >> 1;
        for (int base = 0; base < N; base += len)
            for (int j = 0; j < half; j++) {
                double ang = -2.0 * M_PI * j / len;
                double wr = cos(ang), wi = sin(ang);
                double *a = out + 2 * (base + j), *b = out + 2 * (base + j + half);
                double tr = b[0] * wr - b[1] * wi;
                double ti = b[0] * wi + b[1] * wr;
                b[0] = a[0] - tr;
                b[1] = a[1] - ti;
                a[0] += tr;
                a[1] += ti;
            }
    }
}

/* nativedsp.cpp:72-79.  <string> pulls in <cmath>, so sqrt/log10 on float resolve to
 * the float overloads and `10 * float` stays float. */
static void logmag_shift(const float *spec, float *mag, int N) {
    for (int i = 0; i < N; i++) {
        float realPower = spec[2 * i] / (float)N;
        realPower *= realPower;
        float imagPower = spec[2 * i + 1] / (float)N;
        imagPower *= imagPower;
        int targetIndex = (i + N / 2) % N;
        mag[targetIndex] = 10 * log10f(sqrtf(realPower + imagPower));
    }
}

void orc_fft_logmag(const float *interleaved, float *mag, int N) {
    float *spec = (float *)malloc(sizeof(float) * 2 * (size_t)N);
    orc_fft_c2c_f32(interleaved, spec, N);
    logmag_shift(spec, mag, N);
    free(spec);
}

int orc_windowed_fft_logmag(const float *re, const float *im, int N, int imLen, int magLen, float *mag) {
    if (imLen != N || magLen != N) return 0; /* NativeDsp.kt:44-46 */
    float *w = (float *)malloc(sizeof(float) * (size_t)N);
    float *buf = (float *)malloc(sizeof(float) * 2 * (size_t)N);
    orc_nativedsp_window(N, w);
    for (int i = 0; i < N; i++) { /* NativeDsp.kt:55-58 */
        buf[2 * i] = re[i] * w[i];
        buf[2 * i + 1] = im[i] * w[i];
    }
    orc_fft_logmag(buf, mag, N);
    free(w);
    free(buf);
    return 1;
}

/* ======================================================================== */
/* FftProcessor: waterfall ring, peak hold, signal strength                   */
/*   A/analyzer/FftProcessor.kt:143-157, 163-245                             */
/* ======================================================================== */
struct orc_fftproc {
    int ringRows, peakHold;
    int N;
    float *ring; /* [ringRows][N] */
    float *peaks;
    int havePeaks;
    int writeIndex, readIndex;
    int haveLast;
    long long lastFrequency, lastSampleRate;
    int frequencyOrSampleRateChanged;
};

orc_fftproc *orc_fftproc_new(int ringRows, int peakHold) {
    orc_fftproc *p = (orc_fftproc *)calloc(1, sizeof(*p));
    p->ringRows = ringRows;
    p->peakHold = peakHold;
    p->frequencyOrSampleRateChanged = 1;
    return p;
}
void orc_fftproc_free(orc_fftproc *p) {
    if (!p) return;
    free(p->ring);
    free(p->peaks);
    free(p);
}
const float *orc_fftproc_row(const orc_fftproc *p, int idx) { return p->ring + (size_t)idx * p->N; }
const float *orc_fftproc_peaks(const orc_fftproc *p) { return p->havePeaks ? p->peaks : NULL; }
int orc_fftproc_read_index(const orc_fftproc *p) { return p->readIndex; }
int orc_fftproc_write_index(const orc_fftproc *p) { return p->writeIndex; }
int orc_fftproc_rows(const orc_fftproc *p) { return p->ringRows; }

int orc_fftproc_push(orc_fftproc *p, const float *mag, int N, long long frequency, int sampleRate) {
    float samplesPerHz = N / (float)sampleRate; /* FftProcessor.kt:143 */
    int frequencyChanged = !p->haveLast || frequency != p->lastFrequency;
    int sampleRateChanged = !p->haveLast || (long long)sampleRate != p->lastSampleRate;
    p->frequencyOrSampleRateChanged = frequencyChanged || sampleRateChanged; /* :170-172 */
    long long frequencyDiff = p->haveLast ? p->lastFrequency - frequency : 0; /* :174 */
    p->lastFrequency = frequency;
    p->lastSampleRate = sampleRate;
    p->haveLast = 1;

    if (p->ring == NULL || p->N != N) { /* :180-184 */
        free(p->ring);
        p->N = N;
        p->ring = (float *)malloc(sizeof(float) * (size_t)p->ringRows * (size_t)N);
        for (size_t i = 0; i < (size_t)p->ringRows * (size_t)N; i++) p->ring[i] = -9999.0f;
        p->writeIndex = 0;
    }
    if (frequencyDiff != 0) { /* :199-217: shift history after a retune */
        int shiftOffset = j2i((double)((float)frequencyDiff * samplesPerHz));
        int shiftLeft = shiftOffset < 0;
        if ((shiftLeft && shiftOffset * -1 < N) || (!shiftLeft && shiftOffset < N)) {
            for (int r = 0; r < p->ringRows; r++) {
                float *it = p->ring + (size_t)r * N;
                if (shiftLeft) {
                    memmove(it, it + (-shiftOffset), sizeof(float) * (size_t)(N + shiftOffset));
                    for (int i = N + shiftOffset; i < N; i++) it[i] = -9999.0f;
                } else {
                    memmove(it + shiftOffset, it, sizeof(float) * (size_t)(N - shiftOffset));
                    for (int i = 0; i < shiftOffset; i++) it[i] = -9999.0f;
                }
            }
        } else {
            for (size_t i = 0; i < (size_t)p->ringRows * (size_t)N; i++) p->ring[i] = -9999.0f;
        }
    } else if (sampleRateChanged) { /* :218-222 */
        for (size_t i = 0; i < (size_t)p->ringRows * (size_t)N; i++) p->ring[i] = -9999.0f;
    }
    memcpy(p->ring + (size_t)p->writeIndex * N, mag, sizeof(float) * (size_t)N); /* :224 */
    p->readIndex = p->writeIndex;                                               /* :228 */
    p->writeIndex = (p->writeIndex == 0) ? p->ringRows - 1 : p->writeIndex - 1; /* :229 */

    if (p->peakHold) { /* :232-245 */
        if (!p->havePeaks || p->peaks == NULL) {
            free(p->peaks);
            p->peaks = (float *)malloc(sizeof(float) * (size_t)N);
            for (int i = 0; i < N; i++) p->peaks[i] = -999999.0f;
            p->havePeaks = 1;
        }
        if (p->frequencyOrSampleRateChanged)
            for (int i = 0; i < N; i++) p->peaks[i] = -999999.0f;
        const float *row = p->ring + (size_t)p->readIndex * N;
        for (int i = 0; i < N; i++) p->peaks[i] = p->peaks[i] > row[i] ? p->peaks[i] : row[i];
    } else {
        p->havePeaks = 0;
    }
    return p->readIndex;
}

static int coerce_in(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

int orc_signal_strength(const float *mag, int N, long long frequency, int sampleRate,
                        long long chanStart, long long chanEnd, float *out) {
    float samplesPerHz = N / (float)sampleRate;                /* FftProcessor.kt:143 */
    long long frequencyAtIndexZero = frequency - sampleRate / 2; /* :144 */
    int s = coerce_in(j2i((double)((float)(chanStart - frequencyAtIndexZero) * samplesPerHz)), 0, N);
    int e = coerce_in(j2i((double)((float)(chanEnd - frequencyAtIndexZero) * samplesPerHz)), 0, N);
    if (e > s) {
        float sum = 0.0f;
        for (int i = s; i < e; i++) sum += mag[i];
        *out = sum / (e - s);
        return 1;
    }
    return 0;
}

void orc_time_average(const orc_fftproc *p, int L, float *avg) {
    /* AnalyzerSurface.kt:683-684,710-714 with one bin per pixel: rows newest->oldest */
    int N = p->N;
    for (int i = 0; i < N; i++) avg[i] = 0.0f;
    for (int rowNumber = 0; rowNumber <= L && rowNumber < p->ringRows; rowNumber++) {
        int bufferIndex = (p->readIndex + rowNumber) % p->ringRows;
        const float *row = p->ring + (size_t)bufferIndex * N;
        for (int i = 0; i < N; i++) avg[i] += row[i];
    }
    for (int i = 0; i < N; i++) avg[i] = avg[i] / (L + 1);
}

void orc_ema_rows(const float *rows, long long frames, int N, float alpha, const float *init, float *avg) {
    /* not in the reference (AnalyzerSurface.kt:710-714 is a box-car mean): the sequential definition of the
     * north-star's exponential-averaging option, compiled without FMA contraction like everything here */
    for (int i = 0; i < N; i++) {
        long long f = 0;
        float a;
        if (init) {
            a = init[i];
        } else {
            a = rows[i];
            f = 1;
        }
        for (; f < frames; f++) {
            float d = rows[(size_t)f * N + i] - a;
            float t = alpha * d;
            a = a + t;
        }
        avg[i] = a;
    }
}

void orc_draw_preprocess(const orc_fftproc *p, int width, int fftHeight,
                         long long viewportFrequency, long long viewportSampleRate,
                         float minDB, float maxDB, int L, int colorMapSize,
                         float *timeAverage, int *colorIndex, float *peaksY) {
    /* AnalyzerSurface.kt:646-734 (arithmetic only; every row treated as dirty) */
    int fftSize = p->N;
    int rows = p->ringRows;
    long long frequency = p->lastFrequency, sampleRate = p->lastSampleRate;
    float samplesPerHz = (float)fftSize / (float)sampleRate;
    long long frequencyDiff = viewportFrequency - frequency;
    long long sampleRateDiff = viewportSampleRate - sampleRate;
    int start = j2i(((double)frequencyDiff - sampleRateDiff / 2.0) * samplesPerHz);
    int end = fftSize + j2i(((double)frequencyDiff + sampleRateDiff / 2.0) * samplesPerHz);
    float samplesPerPx = (float)(end - start) / (float)width;
    float dbDiff = maxDB - minDB;
    float dbWidth = fftHeight / dbDiff;
    float scale = colorMapSize / dbDiff;
    int firstPixel = start >= 0 ? 0 : j2i((double)((start * -1) / samplesPerPx));
    int lastPixel = end >= fftSize ? j2i((double)((fftSize - start) / samplesPerPx))
                                   : j2i((double)((end - start) / samplesPerPx));
    float *sum = (float *)calloc((size_t)width, sizeof(float));
    int calcPeaks = p->havePeaks && peaksY != NULL;
    for (int i = 0; i < width; i++) timeAverage[i] = NAN;
    for (int rowNumber = 0; rowNumber < rows; rowNumber++) {
        int bufferIndex = (p->readIndex + rowNumber) % rows;
        const float *fftRow = p->ring + (size_t)bufferIndex * fftSize;
        for (int i = 0; i < width; i++) {
            if (i >= firstPixel + 1 && i < lastPixel - 1) {
                float avg = 0.0f, peakAvg = 0.0f;
                int counter = 0;
                int j = j2i((double)(i * samplesPerPx));
                while (j < (i + 1) * samplesPerPx && (j + start) < fftSize) {
                    avg += fftRow[j + start];
                    if (rowNumber == 0 && calcPeaks) peakAvg += p->peaks[j + start];
                    counter++;
                    j++;
                }
                avg /= counter;
                if (rowNumber == 0 && calcPeaks) peaksY[i] = fftHeight - (peakAvg / counter - minDB) * dbWidth;
                if (rowNumber <= L) sum[i] += avg;
                if (rowNumber == L) timeAverage[i] = sum[i] / (L + 1);
                int idx = j2i((double)((avg - minDB) * scale));
                colorIndex[(size_t)bufferIndex * width + i] = idx < 0 ? 0 : (idx >= colorMapSize ? colorMapSize - 1 : idx);
            } else {
                colorIndex[(size_t)bufferIndex * width + i] = -1;
                if (calcPeaks) peaksY[i] = -1.0f;
            }
        }
    }
    free(sum);
}

/* ======================================================================== */
/* Window functions  (A/dsp/WindowFunctions.kt:44-100)                       */
/* ======================================================================== */
static double kaiser_izero(double x) { /* :86-99 */
    double sum = 1.0, term = 1.0, halfX = x / 2.0;
    for (int k = 1;; k++) {
        double tmp = halfX / k;
        term *= tmp * tmp;
        sum += term;
        if (term < 1e-12) break;
    }
    return sum;
}

float orc_window_value(int kind, double beta, int n, int N) {
    if (kind == ORC_WIN_BLACKMAN) { /* :45-49: float arithmetic on separately cast cosines */
        float c1 = (float)cos(2.0 * M_PI * n / (N - 1));
        float c2 = (float)cos(4.0 * M_PI * n / (N - 1));
        float a = 0.5f * c1;
        float b = 0.08f * c2;
        float t = 0.42f - a;
        return t + b;
    }
    if (kind == ORC_WIN_HAMMING) { /* :56-58 */
        float c1 = (float)cos(2.0 * M_PI * n / (N - 1));
        float a = 0.46f * c1;
        return 0.54f - a;
    }
    /* Kaiser :65-82 */
    double iBeta = 1.0 / kaiser_izero(beta);
    if (n == 0 || n == N - 1) return (float)iBeta;
    double inm1 = 1.0 / (double)(N - 1);
    double temp = 2.0 * n * inm1 - 1.0;
    double valN = kaiser_izero(beta * sqrt(1.0 - temp * temp)) * iBeta;
    return (float)valN;
}

/* ======================================================================== */
/* FirFilter  (A/dsp/FirFilter.kt:34-263)                                    */
/* ======================================================================== */
int orc_lowpass_taps(float gain, float sampleRate, float cutoff, float transitionWidth,
                     float attenuation, int windowKind, double beta, int maxTaps, float **tapsOut) {
    *tapsOut = NULL;
    if (sampleRate <= 0.0) return 0;                          /* :194-197 */
    if (cutoff <= 0.0 || cutoff > sampleRate / 2) return 0;   /* :199-202 */
    if (transitionWidth <= 0) return 0;                       /* :204-207 */
    /* :212: Float*Float in float, divided by Double(22.0*tw) */
    float attFs = attenuation * sampleRate;
    int ntaps = j2i((double)attFs / (22.0 * (double)transitionWidth));
    if (maxTaps > 0 && ntaps > maxTaps) ntaps = maxTaps;
    if (ntaps < 0) return 0; /* JVM: NegativeArraySizeException */
    if ((ntaps & 1) == 0) ntaps++;
    float *taps = (float *)malloc(sizeof(float) * (size_t)ntaps);
    int M = (ntaps - 1) / 2;
    const float PI_F = (float)M_PI;
    float twoPi = 2 * PI_F;              /* Int * Float */
    float t0 = twoPi * cutoff;
    float fwT0 = t0 / sampleRate;        /* :221 */
    for (int n = -M; n <= M; n++) {
        float w = orc_window_value(windowKind, beta, n + M, ntaps);
        if (n == 0) {
            float q = fwT0 / PI_F;
            taps[n + M] = q * w;
        } else {
            float nf = (float)n * fwT0;  /* Int*Float -> Float (n exact in float for |n|<2^24) */
            float s = (float)sin((double)nf);
            float d = (float)n * PI_F;
            float q = s / d;
            taps[n + M] = q * w;
        }
    }
    float fmax = taps[0 + M]; /* :233-234 */
    for (int n = 1; n <= M; n++) {
        float two = 2 * taps[n + M];
        fmax += two;
    }
    float actualGain = gain / fmax;
    for (int i = 0; i < ntaps; i++) taps[i] *= actualGain;
    *tapsOut = taps;
    return ntaps;
}

struct orc_fir {
    float *taps;
    int ntaps, decimation;
    int tapCounter, decimationCounter;
    float *delaysReal, *delaysImag;
};

orc_fir *orc_fir_new(const float *taps, int ntaps, int decimation) {
    orc_fir *f = (orc_fir *)calloc(1, sizeof(*f));
    f->taps = (float *)malloc(sizeof(float) * (size_t)ntaps);
    memcpy(f->taps, taps, sizeof(float) * (size_t)ntaps);
    f->ntaps = ntaps;
    f->decimation = decimation;
    f->decimationCounter = 1; /* FirFilter.kt:46 */
    f->delaysReal = (float *)calloc((size_t)ntaps, sizeof(float));
    f->delaysImag = (float *)calloc((size_t)ntaps, sizeof(float));
    return f;
}
orc_fir *orc_fir_lowpass(int decimation, float gain, float fs, float cutoff, float tw, float att) {
    float *taps;
    int n = orc_lowpass_taps(gain, fs, cutoff, tw, att, ORC_WIN_BLACKMAN, 0.0, 0, &taps);
    if (n == 0) return NULL;
    orc_fir *f = orc_fir_new(taps, n, decimation);
    free(taps);
    return f;
}
void orc_fir_free(orc_fir *f) {
    if (!f) return;
    free(f->taps);
    free(f->delaysReal);
    free(f->delaysImag);
    free(f);
}
int orc_fir_ntaps(const orc_fir *f) { return f->ntaps; }
const float *orc_fir_taps(const orc_fir *f) { return f->taps; }

static int fir_run(orc_fir *f, const orc_packet *in, orc_packet *out, int offset, int length, int complexIn) {
    /* FirFilter.kt:63-110 (filter) and :121-163 (filterReal) */
    int indexOut = out->size;
    int outputCapacity = out->capacity;
    for (int i = 0; i < length; i++) {
        f->delaysReal[f->tapCounter] = in->re[offset + i];
        if (complexIn) f->delaysImag[f->tapCounter] = in->im[offset + i];
        if (f->decimationCounter == 0) {
            if (indexOut == outputCapacity) {
                orc_packet_set_size(out, indexOut);
                out->sampleRate = in->sampleRate / f->decimation;
                return i;
            }
            float accRe = 0.0f, accIm = 0.0f;
            int index = f->tapCounter;
            for (int t = 0; t < f->ntaps; t++) {
                float pr = f->taps[t] * f->delaysReal[index];
                accRe = accRe + pr;
                if (complexIn) {
                    float pi_ = f->taps[t] * f->delaysImag[index];
                    accIm = accIm + pi_;
                }
                index--;
                if (index < 0) index = f->ntaps - 1;
            }
            out->re[indexOut] = accRe;
            if (complexIn) out->im[indexOut] = accIm;
            indexOut++;
        }
        f->decimationCounter++;
        if (f->decimationCounter >= f->decimation) f->decimationCounter = 0;
        f->tapCounter++;
        if (f->tapCounter >= f->ntaps) f->tapCounter = 0;
    }
    orc_packet_set_size(out, indexOut);
    out->sampleRate = in->sampleRate / f->decimation;
    return length;
}
int orc_fir_filter(orc_fir *f, const orc_packet *in, orc_packet *out, int offset, int length) {
    return fir_run(f, in, out, offset, length, 1);
}
int orc_fir_filter_real(orc_fir *f, const orc_packet *in, orc_packet *out, int offset, int length) {
    return fir_run(f, in, out, offset, length, 0);
}

/* ======================================================================== */
/* ComplexFirFilter  (A/dsp/ComplexFirFilter.java:33-279)                    */
/* ======================================================================== */
int orc_bandpass_taps(float gain, float fs, float lo, float hi, float tw, float att,
                      float **tapsReOut, float **tapsImOut) {
    *tapsReOut = *tapsImOut = NULL;
    if (fs <= 0.0) return 0;                                              /* :193-196 */
    if ((double)lo < (double)fs * -0.5 || (double)hi > (double)fs * 0.5) return 0; /* :198-201 */
    if (lo >= hi) return 0;                                               /* :203-206 */
    if (tw <= 0) return 0;                                                /* :208-211 */
    float attFs = att * fs;
    int ntaps = j2i((double)attFs / (22.0 * (double)tw)); /* :216 */
    if (ntaps < 0) return 0;
    if ((ntaps & 1) == 0) ntaps++;
    float lowPassCutOff = (hi - lo) / 2.0f; /* :223 */
    float *lp = (float *)malloc(sizeof(float) * (size_t)ntaps);
    int M = (ntaps - 1) / 2;
    const float PI_F = (float)M_PI;
    float twoPi = 2 * PI_F;
    float t0 = twoPi * lowPassCutOff;
    float fwT0 = t0 / fs; /* :228 */
    for (int n = -M; n <= M; n++) {
        int i = n + M;
        /* makeWindow :270-278 (same float expression as WindowFunctions' Blackman) */
        float w = orc_window_value(ORC_WIN_BLACKMAN, 0.0, i, ntaps);
        if (n == 0) {
            float q = fwT0 / PI_F;
            lp[i] = q * w;
        } else {
            float nf = (float)n * fwT0;
            float s = (float)sin((double)nf);
            float d = (float)n * PI_F;
            float q = s / d;
            lp[i] = q * w;
        }
    }
    float fmax = lp[M]; /* :240-245 */
    for (int n = 1; n <= M; n++) {
        float two = 2 * lp[n + M];
        fmax += two;
    }
    float actualGain = gain / fmax;
    for (int i = 0; i < ntaps; i++) lp[i] *= actualGain;

    float *tr = (float *)malloc(sizeof(float) * (size_t)ntaps);
    float *ti = (float *)malloc(sizeof(float) * (size_t)ntaps);
    float sumf = hi + lo; /* :250: (float)PI * (hi+lo) / fs, left to right */
    float fq0 = PI_F * sumf;
    float freq = fq0 / fs;
    float phase = -freq * (float)(ntaps / 2); /* :251 */
    for (int i = 0; i < ntaps; i++) {
        tr[i] = lp[i] * (float)cos((double)phase);
        ti[i] = lp[i] * (float)sin((double)phase);
        phase += freq;
    }
    free(lp);
    *tapsReOut = tr;
    *tapsImOut = ti;
    return ntaps;
}

struct orc_cfir {
    float *tapsReal, *tapsImag;
    int ntaps, decimation;
    int tapCounter, decimationCounter;
    float *delaysReal, *delaysImag;
    float lowCut, highCut;
};

orc_cfir *orc_cfir_bandpass(int decimation, float gain, float fs, float lo, float hi, float tw, float att) {
    float *tr, *ti;
    int n = orc_bandpass_taps(gain, fs, lo, hi, tw, att, &tr, &ti);
    if (n == 0) return NULL;
    orc_cfir *f = (orc_cfir *)calloc(1, sizeof(*f));
    f->tapsReal = tr;
    f->tapsImag = ti;
    f->ntaps = n;
    f->decimation = decimation;
    f->decimationCounter = 1; /* ComplexFirFilter.java:40 */
    f->delaysReal = (float *)calloc((size_t)n, sizeof(float));
    f->delaysImag = (float *)calloc((size_t)n, sizeof(float));
    f->lowCut = lo;
    f->highCut = hi;
    return f;
}
void orc_cfir_free(orc_cfir *f) {
    if (!f) return;
    free(f->tapsReal);
    free(f->tapsImag);
    free(f->delaysReal);
    free(f->delaysImag);
    free(f);
}
int orc_cfir_ntaps(const orc_cfir *f) { return f->ntaps; }
const float *orc_cfir_taps_re(const orc_cfir *f) { return f->tapsReal; }
const float *orc_cfir_taps_im(const orc_cfir *f) { return f->tapsImag; }

int orc_cfir_filter(orc_cfir *f, const orc_packet *in, orc_packet *out, int offset, int length) {
    /* ComplexFirFilter.java:123-170 */
    int indexOut = out->size;
    int outputCapacity = out->capacity;
    for (int i = 0; i < length; i++) {
        f->delaysReal[f->tapCounter] = in->re[offset + i];
        f->delaysImag[f->tapCounter] = in->im[offset + i];
        if (f->decimationCounter == 0) {
            if (indexOut == outputCapacity) {
                orc_packet_set_size(out, indexOut);
                out->sampleRate = in->sampleRate / f->decimation;
                return i;
            }
            float accRe = 0.0f, accIm = 0.0f;
            int index = f->tapCounter;
            for (int j = 0; j < f->ntaps; j++) {
                float a = f->tapsReal[j] * f->delaysReal[index];
                float b = f->tapsImag[j] * f->delaysImag[index];
                float c = f->tapsImag[j] * f->delaysReal[index];
                float d = f->tapsReal[j] * f->delaysImag[index];
                float e = a - b;
                float g = c + d;
                accRe = accRe + e;
                accIm = accIm + g;
                index--;
                if (index < 0) index = f->ntaps - 1;
            }
            out->re[indexOut] = accRe;
            out->im[indexOut] = accIm;
            indexOut++;
        }
        f->decimationCounter++;
        if (f->decimationCounter >= f->decimation) f->decimationCounter = 0;
        f->tapCounter++;
        if (f->tapCounter >= f->ntaps) f->tapCounter = 0;
    }
    orc_packet_set_size(out, indexOut);
    out->sampleRate = in->sampleRate / f->decimation;
    return length;
}

/* ======================================================================== */
/* RationalResampler  (A/dsp/RationalResampler.kt:36-257)                    */
/* ======================================================================== */
int orc_gcd(int a, int b) { /* :161-170 */
    int x = a < 0 ? -a : a, y = b < 0 ? -b : b;
    while (y != 0) {
        int t = y;
        y = x % y;
        x = t;
    }
    return x;
}

void orc_limit_denominator(int numerator, int denominator, int maxDenominator, int *outNum, int *outDen) {
    /* :183-223 */
    double target = (double)numerator / (double)denominator;
    int g0 = orc_gcd(numerator, denominator);
    int simpleNum = numerator / g0, simpleDen = denominator / g0;
    if (simpleDen <= maxDenominator) {
        *outNum = simpleNum;
        *outDen = simpleDen;
        return;
    }
    int lowerNum = 0, lowerDen = 1, upperNum = 1, upperDen = 0;
    for (;;) {
        int mediantNum = lowerNum + upperNum, mediantDen = lowerDen + upperDen;
        if (mediantDen > maxDenominator) break;
        if ((double)mediantNum / mediantDen < target) {
            lowerNum = mediantNum;
            lowerDen = mediantDen;
        } else {
            upperNum = mediantNum;
            upperDen = mediantDen;
        }
    }
    double lowerError = fabs(target - (double)lowerNum / lowerDen);
    double upperError = fabs(target - (double)upperNum / upperDen);
    if (lowerError < upperError) {
        *outNum = lowerNum;
        *outDen = lowerDen;
    } else {
        *outNum = upperNum;
        *outDen = upperDen;
    }
}

int orc_design_resampler_taps(int interpolation, int decimation, float fractionalBw, int maxTaps, float **taps) {
    /* :230-255 */
    double beta = 7.0, halfband = 0.5;
    float rate = (float)interpolation / (float)decimation;
    float transWidth, midTransitionBand;
    if (rate >= 1.0f) {
        transWidth = (float)(halfband - (double)fractionalBw);
        midTransitionBand = (float)(halfband - (double)transWidth / 2.0);
    } else {
        transWidth = (float)((double)rate * (halfband - (double)fractionalBw));
        midTransitionBand = (float)((double)rate * halfband - (double)transWidth / 2.0);
    }
    int n = orc_lowpass_taps((float)interpolation, (float)interpolation, midTransitionBand, transWidth,
                             72.22087f, ORC_WIN_KAISER, beta, maxTaps * interpolation, taps);
    return n;
}

struct orc_resampler {
    int interpolation, decimation;
    int nt; /* taps per phase */
    float *bank; /* [interpolation][nt] */
    float *delayReal, *delayImag;
    int delayIndex, ctr;
};

orc_resampler *orc_resampler_new(int interpolation, int decimation, const float *tapsIn, int ntapsIn,
                                 float fractionalBw, int maxTaps) {
    /* :53-84 */
    if (interpolation <= 0 || decimation <= 0) return NULL;
    if (fractionalBw <= 0 || fractionalBw >= 0.5f) fractionalBw = 0.4f;
    int d = orc_gcd(interpolation, decimation);
    interpolation /= d;
    decimation /= d;
    float *staps = NULL;
    int ns;
    if (tapsIn) {
        ns = ntapsIn;
        staps = (float *)malloc(sizeof(float) * (size_t)(ns > 0 ? ns : 1));
        memcpy(staps, tapsIn, sizeof(float) * (size_t)ns);
    } else {
        ns = orc_design_resampler_taps(interpolation, decimation, fractionalBw, maxTaps, &staps);
    }
    int padded = ns;
    int rem = ns % interpolation;
    if (rem > 0) padded += interpolation - rem;
    int nt = padded / interpolation;
    orc_resampler *r = (orc_resampler *)calloc(1, sizeof(*r));
    r->interpolation = interpolation;
    r->decimation = decimation;
    r->nt = nt;
    r->bank = (float *)calloc((size_t)interpolation * (size_t)(nt > 0 ? nt : 1), sizeof(float));
    for (int phase = 0; phase < interpolation; phase++)
        for (int i = 0; i < nt; i++) {
            int src = i * interpolation + phase;
            r->bank[(size_t)phase * nt + i] = src < ns ? staps[src] : 0.0f;
        }
    r->delayReal = (float *)calloc((size_t)(nt > 0 ? nt : 1), sizeof(float));
    r->delayImag = (float *)calloc((size_t)(nt > 0 ? nt : 1), sizeof(float));
    free(staps);
    return r;
}
void orc_resampler_free(orc_resampler *r) {
    if (!r) return;
    free(r->bank);
    free(r->delayReal);
    free(r->delayImag);
    free(r);
}
int orc_resampler_interp(const orc_resampler *r) { return r->interpolation; }
int orc_resampler_decim(const orc_resampler *r) { return r->decimation; }
int orc_resampler_taps_per_phase(const orc_resampler *r) { return r->nt; }
void orc_resampler_bank(const orc_resampler *r, float *bank) {
    memcpy(bank, r->bank, sizeof(float) * (size_t)r->interpolation * (size_t)r->nt);
}

int orc_resampler_resample(orc_resampler *r, const orc_packet *in, orc_packet *out, int offset, int length) {
    /* :90-156 */
    int outputCapacity = out->capacity;
    int indexOut = out->size;
    int consumed = 0;
    int inIdx = offset;
    int nt = r->nt;
    r->delayReal[r->delayIndex] = in->re[inIdx];
    r->delayImag[r->delayIndex] = in->im[inIdx];
    while (r->ctr >= r->interpolation) { /* :104-117 */
        r->ctr -= r->interpolation;
        inIdx++;
        if (++r->delayIndex >= nt) r->delayIndex = 0;
        consumed++;
        if (consumed >= length) break;
        r->delayReal[r->delayIndex] = in->re[inIdx];
        r->delayImag[r->delayIndex] = in->im[inIdx];
    }
    while (consumed < length && indexOut < outputCapacity) { /* :119-150 */
        float reSum = 0.0f, imSum = 0.0f;
        const float *taps = r->bank + (size_t)r->ctr * nt;
        int di = r->delayIndex;
        for (int t = 0; t < nt; t++) {
            float pr = taps[t] * r->delayReal[di];
            float pi_ = taps[t] * r->delayImag[di];
            reSum = reSum + pr;
            imSum = imSum + pi_;
            if (--di < 0) di = nt - 1;
        }
        out->re[indexOut] = reSum;
        out->im[indexOut] = imSum;
        indexOut++;
        r->ctr += r->decimation;
        while (r->ctr >= r->interpolation) {
            r->ctr -= r->interpolation;
            inIdx++;
            if (++r->delayIndex >= nt) r->delayIndex = 0;
            consumed++;
            if (consumed >= length) break;
            r->delayReal[r->delayIndex] = in->re[inIdx];
            r->delayImag[r->delayIndex] = in->im[inIdx];
        }
    }
    orc_packet_set_size(out, indexOut);
    out->sampleRate = (int)((long long)in->sampleRate * r->interpolation / r->decimation);
    out->frequency = in->frequency;
    return consumed;
}

/* ======================================================================== */
/* Demodulator  (A/analyzer/Demodulator.kt:40-404)                           */
/* ======================================================================== */
static const int MODE_MIN_CW[] = {0, 3000, 3000, 30000, 1500, 1500, 150};     /* DemodulationTab.kt:91-97 */
static const int MODE_MAX_CW[] = {50000, 15000, 15000, 150000, 5000, 5000, 800};
static const int MODE_DEF_CW[] = {0, 8000, 10000, 100000, 2800, 2800, 300};
#define AUDIO_RATE 48000
#define BAND_PASS_ATTENUATION 40
#define USER_FILTER_ATTENUATION 60
#define CW_OFFSET_FREQUENCY 750

int orc_mode_quadrature_rate(int mode) { /* Demodulator.kt:53-62 */
    switch (mode) {
        case ORC_MODE_WFM: return 8 * AUDIO_RATE;
        case ORC_MODE_CW: return 1 * AUDIO_RATE;
        default: return 2 * AUDIO_RATE;
    }
}

struct orc_demod {
    int mode, channelWidth;
    float volume;
    orc_fir *userFilter;
    float userFilterCutoff;
    orc_packet *quadratureSamples;
    float carryRe, carryIm, lastMax;
    orc_cfir *bandPass;
};

orc_demod *orc_demod_new(int packetSize) {
    orc_demod *d = (orc_demod *)calloc(1, sizeof(*d));
    d->mode = ORC_MODE_OFF;
    d->volume = 1.0f;
    d->quadratureSamples = orc_packet_new(packetSize);
    return d;
}
void orc_demod_free(orc_demod *d) {
    if (!d) return;
    orc_fir_free(d->userFilter);
    orc_cfir_free(d->bandPass);
    orc_packet_free(d->quadratureSamples);
    free(d);
}
void orc_demod_set_channel_width(orc_demod *d, int w) { /* :75-76 */
    d->channelWidth = coerce_in(w, MODE_MIN_CW[d->mode], MODE_MAX_CW[d->mode]);
}
void orc_demod_set_mode(orc_demod *d, int mode) { /* :97-101 */
    d->mode = mode;
    orc_demod_set_channel_width(d, MODE_DEF_CW[mode]);
}
int orc_demod_channel_width(const orc_demod *d) { return d->channelWidth; }
void orc_demod_set_volume(orc_demod *d, float v) { d->volume = v; }

void orc_demod_user_filter(orc_demod *d, const orc_packet *input, orc_packet *output) {
    /* :215-240 */
    if (d->userFilter == NULL || j2i((double)d->userFilterCutoff) != d->channelWidth) {
        orc_fir_free(d->userFilter);
        d->userFilterCutoff = (float)d->channelWidth;
        d->userFilter = orc_fir_lowpass(1, 1.0f, (float)input->sampleRate, (float)d->channelWidth,
                                        input->sampleRate * 0.10f, (float)USER_FILTER_ATTENUATION);
        if (d->userFilter == NULL) return;
    }
    output->size = 0;
    orc_fir_filter(d->userFilter, input, output, 0, input->size);
}

void orc_demod_fm(orc_demod *d, const orc_packet *input, orc_packet *output, float maxDeviation) {
    /* :251-275 */
    int inputSize = input->size;
    float quadratureGain = orc_mode_quadrature_rate(d->mode) / (float)(2 * M_PI * (double)maxDeviation);
    if (inputSize == 0) return;
    const float *reIn = input->re, *imIn = input->im;
    {
        float a = reIn[0] * d->carryRe, b = imIn[0] * d->carryIm;
        float c = imIn[0] * d->carryRe, e = reIn[0] * d->carryIm;
        float re = a + b, im = c - e;
        output->im[0] = im;
        output->re[0] = quadratureGain * (float)atan2((double)im, (double)re);
    }
    for (int i = 1; i < inputSize; i++) {
        float a = reIn[i] * reIn[i - 1], b = imIn[i] * imIn[i - 1];
        float c = imIn[i] * reIn[i - 1], e = reIn[i] * imIn[i - 1];
        float re = a + b, im = c - e;
        output->im[i] = im;
        output->re[i] = quadratureGain * (float)atan2((double)im, (double)re);
    }
    d->carryRe = reIn[inputSize - 1];
    d->carryIm = imIn[inputSize - 1];
    orc_packet_set_size(output, inputSize);
    output->sampleRate = orc_mode_quadrature_rate(d->mode);
}

void orc_demod_am(orc_demod *d, const orc_packet *input, orc_packet *output) {
    /* :285-306 */
    float avg = 0.0f;
    d->lastMax *= (float)0.95;
    for (int i = 0; i < input->size; i++) {
        float a = input->re[i] * input->re[i], b = input->im[i] * input->im[i];
        output->re[i] = a + b;
        avg += output->re[i];
        if (output->re[i] > d->lastMax) d->lastMax = output->re[i];
    }
    avg /= input->size;
    float gain = 0.75f / d->lastMax;
    for (int i = 0; i < input->size; i++) output->re[i] = (output->re[i] - avg) * gain;
    orc_packet_set_size(output, input->size);
    output->sampleRate = orc_mode_quadrature_rate(d->mode);
}

static void agc_normalise(orc_demod *d, orc_packet *output) { /* :347-355, :394-402 */
    d->lastMax *= (float)0.95;
    for (int i = 0; i < output->size; i++)
        if (output->re[i] > d->lastMax) d->lastMax = output->re[i];
    float gain = 0.75f / d->lastMax;
    for (int i = 0; i < output->size; i++) output->re[i] *= gain;
}

void orc_demod_ssb(orc_demod *d, const orc_packet *input, orc_packet *output, int upperBand) {
    /* :317-356 */
    if (d->bandPass == NULL || (upperBand && (j2i((double)d->bandPass->highCut) != d->channelWidth)) ||
        (!upperBand && (j2i((double)d->bandPass->lowCut) != -d->channelWidth))) {
        orc_cfir_free(d->bandPass);
        d->bandPass = orc_cfir_bandpass(2, 1.0f, (float)input->sampleRate,
                                        upperBand ? 200.0f : -(float)d->channelWidth,
                                        upperBand ? (float)d->channelWidth : -200.0f,
                                        input->sampleRate * 0.01f, (float)BAND_PASS_ATTENUATION);
        if (d->bandPass == NULL) return;
    }
    output->size = 0;
    orc_cfir_filter(d->bandPass, input, output, 0, input->size);
    agc_normalise(d, output);
}

void orc_demod_cw(orc_demod *d, const orc_packet *input, orc_packet *output) {
    /* :366-403 */
    if (d->bandPass == NULL || j2i((double)d->bandPass->highCut) != CW_OFFSET_FREQUENCY + d->channelWidth / 2) {
        orc_cfir_free(d->bandPass);
        d->bandPass = orc_cfir_bandpass(1, 1.0f, (float)input->sampleRate,
                                        CW_OFFSET_FREQUENCY - d->channelWidth / 2.0f,
                                        CW_OFFSET_FREQUENCY + d->channelWidth / 2.0f,
                                        input->sampleRate * 0.01f, (float)BAND_PASS_ATTENUATION);
        if (d->bandPass == NULL) return;
    }
    output->size = 0;
    orc_cfir_filter(d->bandPass, input, output, 0, input->size);
    agc_normalise(d, output);
}

void orc_demod_process(orc_demod *d, const orc_packet *resampled, orc_packet *audio) {
    /* Demodulator.run :147-187 */
    orc_demod_user_filter(d, resampled, d->quadratureSamples);
    audio->size = 0;
    switch (d->mode) {
        case ORC_MODE_AM: orc_demod_am(d, d->quadratureSamples, audio); break;
        case ORC_MODE_NFM: orc_demod_fm(d, d->quadratureSamples, audio, d->channelWidth * 0.75f); break;
        case ORC_MODE_WFM: orc_demod_fm(d, d->quadratureSamples, audio, d->channelWidth * 0.85f); break;
        case ORC_MODE_LSB: orc_demod_ssb(d, d->quadratureSamples, audio, 0); break;
        case ORC_MODE_USB: orc_demod_ssb(d, d->quadratureSamples, audio, 1); break;
        case ORC_MODE_CW: orc_demod_cw(d, d->quadratureSamples, audio); break;
        default: break;
    }
    for (int i = 0; i < audio->size; i++) audio->re[i] = audio->re[i] * d->volume;
}

/* ======================================================================== */
/* AudioSink filters  (A/analyzer/AudioSink.java:94-96, 215-237)             */
/* ======================================================================== */
struct orc_audiosink {
    int sampleRate;
    orc_fir *f1, *f2;
    orc_packet *tmp;
};
orc_audiosink *orc_audiosink_new(int packetSize, int sampleRate) {
    orc_audiosink *a = (orc_audiosink *)calloc(1, sizeof(*a));
    a->sampleRate = sampleRate;
    a->f1 = orc_fir_lowpass(2, 1, 1, 0.1f, 0.15f, 30);
    a->f2 = orc_fir_lowpass(4, 1, 1, 0.1f, 0.1f, 30);
    a->tmp = orc_packet_new(packetSize);
    return a;
}
void orc_audiosink_free(orc_audiosink *a) {
    if (!a) return;
    orc_fir_free(a->f1);
    orc_fir_free(a->f2);
    orc_packet_free(a->tmp);
    free(a);
}
int orc_audiosink_filter(orc_audiosink *a, const orc_packet *input, orc_packet *output) {
    if (input->sampleRate / a->sampleRate == 8) {
        a->tmp->size = 0;
        orc_fir_filter_real(a->f1, input, a->tmp, 0, input->size);
        output->size = 0;
        orc_fir_filter_real(a->f2, a->tmp, output, 0, a->tmp->size);
        return 1;
    } else if (input->sampleRate / a->sampleRate == 2) {
        output->size = 0;
        orc_fir_filter_real(a->f1, input, output, 0, input->size);
        return 1;
    }
    return 0;
}

/* ======================================================================== */
/* Whole chains                                                              */
/* ======================================================================== */
static int fmt_bytes_per_sample(int fmt) { return fmt == ORC_FMT_S16LE ? 4 : 2; }

long long orc_spectrum_run(int fmt, const uint8_t *iq, long long nsamples, int N, int L,
                           float *rows, float *peaks, float *avg) {
    /* Scheduler.kt:254-272 (fill) -> FftProcessor.kt:135 (FFT) -> :224-245 (rows, peaks)
     * -> AnalyzerSurface.kt:710-714 (time average), loss-free and contiguous. */
    long long F = nsamples / N;
    int bps = fmt_bytes_per_sample(fmt);
    orc_converter *c = orc_converter_new(fmt);
    orc_converter_set_sample_rate(c, 1000000);
    orc_packet *sp = orc_packet_new(N);
    float *mag = (float *)malloc(sizeof(float) * (size_t)N);
    if (peaks)
        for (int i = 0; i < N; i++) peaks[i] = -999999.0f;
    for (long long f = 0; f < F; f++) {
        sp->size = 0;
        orc_converter_fill(c, iq + (size_t)f * N * bps, N * bps, sp);
        orc_windowed_fft_logmag(sp->re, sp->im, N, N, N, mag);
        if (rows) memcpy(rows + (size_t)f * N, mag, sizeof(float) * (size_t)N);
        if (peaks)
            for (int i = 0; i < N; i++) peaks[i] = peaks[i] > mag[i] ? peaks[i] : mag[i];
    }
    if (avg && rows && F > 0) {
        for (int i = 0; i < N; i++) avg[i] = 0.0f;
        for (int rowNumber = 0; rowNumber <= L; rowNumber++) {
            long long f = F - 1 - rowNumber;
            for (int i = 0; i < N; i++) avg[i] += (f >= 0) ? rows[(size_t)f * N + i] : -9999.0f;
        }
        for (int i = 0; i < N; i++) avg[i] = avg[i] / (L + 1);
    }
    free(mag);
    orc_packet_free(sp);
    orc_converter_free(c);
    return F;
}

long long orc_chain_run(int fmt, const uint8_t *iq, long long nsamples, int sampleRate,
                        long long srcFrequency, long long channelFrequency, int mode,
                        int channelWidth, int packetSamples, float volume,
                        float *audio, long long audioCapacity) {
    /* Scheduler.kt:237-244 -> Resampler.kt:95-113 -> Demodulator.kt:147-187 ->
     * AudioSink.java:182-187; every packet delivered (no drops). */
    int bps = fmt_bytes_per_sample(fmt);
    orc_converter *c = orc_converter_new(fmt);
    orc_converter_set_sample_rate(c, sampleRate);
    orc_converter_set_frequency(c, srcFrequency);
    orc_demod *d = orc_demod_new(packetSamples);
    orc_demod_set_mode(d, mode);
    if (channelWidth > 0) orc_demod_set_channel_width(d, channelWidth);
    orc_demod_set_volume(d, volume);
    orc_audiosink *sink = orc_audiosink_new(packetSamples, AUDIO_RATE);
    int outRate = orc_mode_quadrature_rate(mode);
    int I, D;
    orc_limit_denominator(outRate, sampleRate, 10000, &I, &D); /* Resampler.kt:99 */
    orc_resampler *rs = orc_resampler_new(I, D, NULL, 0, 0.4f, 500); /* :102 */
    orc_packet *demodBuffer = orc_packet_new(packetSamples);
    orc_packet *resampled = orc_packet_new(packetSamples);
    orc_packet *audioBuf = orc_packet_new(packetSamples);
    orc_packet *filtered = orc_packet_new(packetSamples);
    long long nAudio = 0;
    for (long long pos = 0; pos < nsamples; pos += packetSamples) {
        long long n = nsamples - pos < packetSamples ? nsamples - pos : packetSamples;
        demodBuffer->size = 0;
        orc_converter_mix(c, iq + (size_t)pos * bps, (int)(n * bps), demodBuffer, channelFrequency);
        resampled->size = 0;
        orc_resampler_resample(rs, demodBuffer, resampled, 0, demodBuffer->size);
        resampled->sampleRate = outRate; /* Resampler.kt:113 */
        orc_demod_process(d, resampled, audioBuf);
        const orc_packet *fin = audioBuf;
        if (audioBuf->sampleRate > AUDIO_RATE) { /* AudioSink.java:182-187 */
            if (orc_audiosink_filter(sink, audioBuf, filtered)) fin = filtered;
        }
        for (int i = 0; i < fin->size && nAudio < audioCapacity; i++) audio[nAudio++] = fin->re[i];
    }
    orc_packet_free(demodBuffer);
    orc_packet_free(resampled);
    orc_packet_free(audioBuf);
    orc_packet_free(filtered);
    orc_resampler_free(rs);
    orc_audiosink_free(sink);
    orc_demod_free(d);
    orc_converter_free(c);
    return nAudio;
}


/* ------------------------------------------------------------------------------------------------
 * Airspy real -> IQ converter, restated from iqconverter_int16.c (line references in rfa_oracle.h).
 * ------------------------------------------------------------------------------------------------ */
struct orc_iqconv {
    int len;            /* cnv->len = len/2 + 1 (:60) */
    int32_t *kernel;    /* hb_kernel[2*i] (:72-75) */
    int32_t *queue;     /* the newest cnv->len even samples, newest first */
    int16_t *delay;     /* cnv->len >> 1 odd samples */
    int delay_index;
    int16_t old_x, old_y;
    int32_t old_e;
};

orc_iqconv *orc_iqconv_new(const int16_t *hb_kernel, int len) {
    orc_iqconv *c = (orc_iqconv *)calloc(1, sizeof(*c));
    c->len = len / 2 + 1;
    c->kernel = (int32_t *)calloc((size_t)c->len, sizeof(int32_t));
    c->queue = (int32_t *)calloc((size_t)c->len, sizeof(int32_t));
    c->delay = (int16_t *)calloc((size_t)c->len, sizeof(int16_t));
    for (int i = 0; i < c->len; i++) c->kernel[i] = hb_kernel[i * 2];
    return c;
}
void orc_iqconv_free(orc_iqconv *c) {
    if (!c) return;
    free(c->kernel);
    free(c->queue);
    free(c->delay);
    free(c);
}
void orc_iqconv_reset(orc_iqconv *c) {
    c->delay_index = 0;
    c->old_x = c->old_y = 0;
    c->old_e = 0;
    memset(c->queue, 0, sizeof(int32_t) * (size_t)c->len);
    memset(c->delay, 0, sizeof(int16_t) * (size_t)c->len);
}
void orc_iqconv_process(orc_iqconv *c, int16_t *samples, long long len) {
    /* remove_dc (:160-186) */
    int16_t old_x = c->old_x, old_y = c->old_y;
    int32_t old_e = c->old_e;
    for (long long i = 0; i < len; i++) {
        int16_t x = samples[i];
        int16_t w = (int16_t)(x - old_x);
        int32_t u = old_e + (int32_t)old_y * 32100;
        int16_t s = (int16_t)(u >> 15);
        int16_t y = (int16_t)(w + s);
        old_e = u - (int32_t)((uint32_t)(int32_t)s << 15);
        old_x = x;
        old_y = y;
        samples[i] = y;
    }
    c->old_x = old_x;
    c->old_y = old_y;
    c->old_e = old_e;
    /* translate_fs_4 (:188-198) */
    for (long long i = 0; i + 3 < len; i += 4) {
        samples[i + 0] = (int16_t)(-samples[i + 0]);
        samples[i + 1] = (int16_t)(-samples[i + 1] >> 1);
        samples[i + 3] = (int16_t)(samples[i + 3] >> 1);
    }
    /* fir_interleaved (:97-134): queue[j] = the even sample j steps ago */
    for (long long i = 0; i < len; i += 2) {
        memmove(c->queue + 1, c->queue, sizeof(int32_t) * (size_t)(c->len - 1));
        c->queue[0] = samples[i];
        uint32_t acc = 0; /* int32 accumulation, wrap-around made explicit */
        for (int j = 0; j < c->len; j++) acc += (uint32_t)(c->kernel[j] * c->queue[j]);
        samples[i] = (int16_t)((int32_t)acc >> 15);
    }
    /* delay_interleaved on samples + 1 (:136-158) */
    int half = c->len >> 1, index = c->delay_index;
    for (long long i = 1; i < len + 1 && i - 1 < len; i += 2) {
        if (i >= len) break;
        int16_t res = c->delay[index];
        c->delay[index] = samples[i];
        samples[i] = res;
        if (++index >= half) index = 0;
    }
    c->delay_index = index;
}
void orc_airspy_convert_samples(const uint16_t *src, int16_t *dst, long long count) {
    for (long long i = 0; i < count; i++) dst[i] = (int16_t)(((int)src[i] - 2048) << 4);
}
