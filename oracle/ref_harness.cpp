/*
 * ref_harness.cpp -- drives the REAL reference native code (TEST INFRASTRUCTURE ONLY).
 *
 * oracle/Makefile compiles /root/reference/nativedsp/src/main/cpp/{pffft.c,nativedsp.cpp}
 * where they lie (never copied), with the reference's own flags for pffft
 * (-O3 -ffast-math, CMakeLists.txt:15), and links them with this file and the
 * oracle restatement into oracle/_ref/librfa_ref.so.
 *
 *  ref_perform_fft / ref_perform_fft_logmag : call the two JNI entry points
 *      (nativedsp.cpp:19-42, :44-81) through the stub JNIEnv.
 *  ref_spectrum_run : the reference CPU spectrum path end to end
 *      (restated JVM stages + reference pffft); the CPU baseline of bench.py.
 */
#include <jni.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "pffft.h"
#include "rfa_oracle.h"

extern "C" {
void Java_com_mantz_1it_nativedsp_NativeDsp_performFFT(JNIEnv *, jobject, jfloatArray, jfloatArray);
void Java_com_mantz_1it_nativedsp_NativeDsp_performFFTAndLogMag(JNIEnv *, jobject, jfloatArray, jfloatArray);
}

/* nativedsp.cpp:12 keeps one process-global `fftSize` for both entry points, and only
 * performFFTAndLogMag allocates `outputMag` (:63): calling it after performFFT with the
 * same length would dereference NULL.  The app never mixes the two; the harness
 * invalidates the cached size whenever the entry point changes. */
extern int fftSize;
static int g_last_kind = 0;
static void switch_kind(int kind) {
    if (g_last_kind != kind) fftSize = -1;
    g_last_kind = kind;
}

extern "C" __attribute__((visibility("default"))) void ref_perform_fft(const float *in, float *out, int len) {
    switch_kind(1);
    JNIEnv env;
    _jfloatArray a{len, const_cast<float *>(in)}, b{len, out};
    Java_com_mantz_1it_nativedsp_NativeDsp_performFFT(&env, nullptr, &a, &b);
}

extern "C" __attribute__((visibility("default"))) void ref_perform_fft_logmag(const float *in, float *out, int len) {
    switch_kind(2);
    JNIEnv env;
    _jfloatArray a{len, const_cast<float *>(in)}, b{len / 2, out};
    Java_com_mantz_1it_nativedsp_NativeDsp_performFFTAndLogMag(&env, nullptr, &a, &b);
}

extern "C" __attribute__((visibility("default"))) int ref_pffft_simd_size(void) { return pffft_simd_size(); }

namespace {
struct Job {
    int fmt, N, tid, nthreads;
    const uint8_t *iq;
    long long F;
    float *rows, *peaks; /* peaks: per-thread partial [N] */
    int useJni;
};

void *worker(void *arg) {
    Job *j = static_cast<Job *>(arg);
    const int N = j->N;
    const int bps = j->fmt == ORC_FMT_S16LE ? 4 : 2;
    long long f0 = j->F * j->tid / j->nthreads, f1 = j->F * (j->tid + 1) / j->nthreads;
    orc_converter *c = orc_converter_new(j->fmt);
    orc_converter_set_sample_rate(c, 1000000);
    orc_packet *sp = orc_packet_new(N);
    float *w = static_cast<float *>(malloc(sizeof(float) * N));
    orc_nativedsp_window(N, w);
    PFFFT_Setup *setup = nullptr;
    float *in = static_cast<float *>(pffft_aligned_malloc(sizeof(float) * 2 * N));
    float *out = static_cast<float *>(pffft_aligned_malloc(sizeof(float) * 2 * N));
    float *scratch = static_cast<float *>(pffft_aligned_malloc(sizeof(float) * 2 * N));
    float *mag = static_cast<float *>(malloc(sizeof(float) * N));
    if (!j->useJni) setup = pffft_new_setup(N, PFFFT_COMPLEX);
    for (int i = 0; i < N; i++) j->peaks[i] = -999999.0f;
    for (long long f = f0; f < f1; f++) {
        sp->size = 0;
        orc_converter_fill(c, j->iq + (size_t)f * N * bps, N * bps, sp); /* Scheduler.kt:266 */
        for (int i = 0; i < N; i++) {                                      /* NativeDsp.kt:55-58 */
            in[2 * i] = sp->re[i] * w[i];
            in[2 * i + 1] = sp->im[i] * w[i];
        }
        float *dst = j->rows ? j->rows + (size_t)f * N : mag;
        if (j->useJni) {
            ref_perform_fft_logmag(in, dst, 2 * N); /* the reference's own loop */
        } else {
            pffft_transform_ordered(setup, in, out, scratch, PFFFT_FORWARD);
            for (int i = 0; i < N; i++) { /* nativedsp.cpp:72-79, per-thread restatement */
                float realPower = out[2 * i] / (float)N;
                realPower *= realPower;
                float imagPower = out[2 * i + 1] / (float)N;
                imagPower *= imagPower;
                dst[(i + N / 2) % N] = 10 * log10f(sqrtf(realPower + imagPower));
            }
        }
        for (int i = 0; i < N; i++) j->peaks[i] = j->peaks[i] > dst[i] ? j->peaks[i] : dst[i]; /* FftProcessor.kt:244 */
    }
    if (setup) pffft_destroy_setup(setup);
    pffft_aligned_free(in);
    pffft_aligned_free(out);
    pffft_aligned_free(scratch);
    free(mag);
    free(w);
    orc_packet_free(sp);
    orc_converter_free(c);
    return nullptr;
}
}  // namespace

/* nthreads == 1: every frame goes through the real JNI function (process-global
 * state, as in the app).  nthreads > 1: frames are split across threads, each with its
 * own pffft setup (the JNI function's globals are not thread safe, NativeDsp.kt:23-26). */
extern "C" __attribute__((visibility("default"))) long long ref_spectrum_run(
    int fmt, const uint8_t *iq, long long nsamples, int N, int L, float *rows, float *peaks, float *avg,
    int nthreads) {
    long long F = nsamples / N;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    Job jobs[256];
    pthread_t th[256];
    float *partial = static_cast<float *>(malloc(sizeof(float) * (size_t)N * nthreads));
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = Job{fmt, N, t, nthreads, iq, F, rows, partial + (size_t)t * N, nthreads == 1};
        if (nthreads == 1)
            worker(&jobs[t]);
        else
            pthread_create(&th[t], nullptr, worker, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], nullptr);
    if (peaks) {
        for (int i = 0; i < N; i++) peaks[i] = -999999.0f;
        for (int t = 0; t < nthreads; t++)
            for (int i = 0; i < N; i++) peaks[i] = peaks[i] > partial[(size_t)t * N + i] ? peaks[i] : partial[(size_t)t * N + i];
    }
    if (avg && rows && F > 0) { /* AnalyzerSurface.kt:710-714 */
        for (int i = 0; i < N; i++) avg[i] = 0.0f;
        for (int r = 0; r <= L; r++) {
            long long f = F - 1 - r;
            for (int i = 0; i < N; i++) avg[i] += (f >= 0) ? rows[(size_t)f * N + i] : -9999.0f;
        }
        for (int i = 0; i < N; i++) avg[i] = avg[i] / (L + 1);
    }
    free(partial);
    return F;
}


/* ---- the reference's Airspy real -> IQ converter, compiled in place (libairspy/.../iqconverter_int16.c) ----------
 * ref_iqconv_new zeroes the whole delay line: the reference's reset clears only half of it (iqconverter_int16.c:94,
 * `cnv->len * sizeof(int16_t) / 4` bytes of a `cnv->len * sizeof(int32_t) / 4`-byte block), so its first len/4 odd
 * outputs would otherwise be whatever malloc returned. */
extern "C" {
#include "iqconverter_int16.h"
}
#include "filters.h"
extern "C" __attribute__((visibility("default"))) void *ref_iqconv_new(const int16_t *hb_kernel, int len) {
    iqconverter_int16_t *c = iqconverter_int16_create(hb_kernel, len);
    memset(c->delay_line, 0, (size_t)c->len * sizeof(int32_t) / 4);
    return c;
}
extern "C" __attribute__((visibility("default"))) void ref_iqconv_free(void *c) { iqconverter_int16_free((iqconverter_int16_t *)c); }
extern "C" __attribute__((visibility("default"))) void ref_iqconv_process(void *c, int16_t *samples, int len) {
    iqconverter_int16_process((iqconverter_int16_t *)c, samples, len);
}
extern "C" __attribute__((visibility("default"))) int ref_airspy_hb_kernel(int16_t *out, int capacity) {
    if (out && capacity >= HB_KERNEL_INT16_LEN) memcpy(out, HB_KERNEL_INT16, sizeof(HB_KERNEL_INT16));
    return HB_KERNEL_INT16_LEN;
}
