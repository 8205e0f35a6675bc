/*
 * rfa_oracle.h -- CPU oracle for the RF Analyzer IQ->spectrum / IQ->audio hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rfanalyzer_b200/ (the product) may
 * include, link, import or execute anything in oracle/.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker (or as the timed CPU reference arm).
 *
 * It is a plain-C restatement of the reference's JVM DSP (Kotlin/Java) with
 * float32 where the reference says Float/float, double where it says
 * Double/Math.*, no FMA contraction (-ffp-contract=off) and strictly sequential
 * sums, exactly in the reference's order.  Each function cites the reference
 * file:line it follows (paths relative to /root/reference, with
 * A = app/src/main/java/com/mantz_it/rfanalyzer).
 *
 * Parity pinning: see oracle/README.md.  FirFilter/createLowPassTaps are pinned
 * bit-exactly by the reference's own golden vectors (ApplicationTest.kt:55-121,
 * :165-170); the FFT+log-magnitude is pinned against the reference's own
 * pffft.c + nativedsp.cpp compiled in place into oracle/_ref/; the remaining
 * JVM-only stages have no reference test ("parity unpinned by reference tests")
 * and are anchored by the restatement rules only.
 */
#ifndef RFA_ORACLE_H
#define RFA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* sample formats of the three IQConverter subclasses */
enum { ORC_FMT_S8 = 0, ORC_FMT_U8 = 1, ORC_FMT_S16LE = 2 };
/* Demodulator modes, DemodulationTab.kt:90-99 ordinal order */
enum { ORC_MODE_OFF = 0, ORC_MODE_AM, ORC_MODE_NFM, ORC_MODE_WFM, ORC_MODE_LSB, ORC_MODE_USB, ORC_MODE_CW };
/* WindowFunctions.kt */
enum { ORC_WIN_BLACKMAN = 0, ORC_WIN_HAMMING = 1, ORC_WIN_KAISER = 2 };

/* ---- SamplePacket (A/source/SamplePacket.java:28-137) ------------------- */
typedef struct {
    float *re, *im;
    int capacity;
    int size;
    int sampleRate;
    long long frequency;
} orc_packet;

orc_packet *orc_packet_new(int capacity);
void orc_packet_free(orc_packet *p);
void orc_packet_set_size(orc_packet *p, int size); /* clamps to capacity */

/* ---- IQConverter family ------------------------------------------------ */
typedef struct orc_converter orc_converter;
orc_converter *orc_converter_new(int fmt);
void orc_converter_free(orc_converter *c);
void orc_converter_set_frequency(orc_converter *c, long long f);
void orc_converter_set_sample_rate(orc_converter *c, int fs);
const float *orc_converter_lut(const orc_converter *c, int *n);
int orc_converter_fill(orc_converter *c, const uint8_t *packet, int nbytes, orc_packet *sp);
int orc_converter_mix(orc_converter *c, const uint8_t *packet, int nbytes, orc_packet *sp,
                      long long channelFrequency);
/* inspect NCO state: table length, current index, effective mix frequency; cos/sin at t */
int orc_converter_nco_len(const orc_converter *c);
int orc_converter_nco_index(const orc_converter *c);
int orc_converter_nco_freq(const orc_converter *c);
void orc_converter_nco_table(const orc_converter *c, float *cosT, float *sinT);
int orc_calc_optimal_cosine_length(int sampleRate, int cosineFrequency);

/* ---- NativeDsp window + FFT + log magnitude ---------------------------- */
void orc_nativedsp_window(int N, float *w);
/* forward ordered C2C FFT, interleaved float32, unnormalised (restated, radix-2) */
void orc_fft_c2c_f32(const float *in, float *out, int N);
/* double-precision DFT truth (O(N log N), double twiddles) for error budgets */
void orc_fft_c2c_f64(const float *in, double *out, int N);
/* nativedsp.cpp:44-81 loop: FFT then 10*log10f(sqrtf((re/N)^2+(im/N)^2)), shifted */
void orc_fft_logmag(const float *interleaved, float *mag, int N);
/* NativeDsp.kt:43-62: window, interleave, FFT+logmag. returns 0 if sizes mismatch */
int orc_windowed_fft_logmag(const float *re, const float *im, int N, int imLen, int magLen, float *mag);

/* ---- FftProcessor ring / peaks / signal strength (FftProcessor.kt) ----- */
typedef struct orc_fftproc orc_fftproc;
orc_fftproc *orc_fftproc_new(int ringRows /*300,400,500*/, int peakHold);
void orc_fftproc_free(orc_fftproc *p);
/* push one dB row; returns index it was written to (the new readIndex) */
int orc_fftproc_push(orc_fftproc *p, const float *mag, int N, long long frequency, int sampleRate);
const float *orc_fftproc_row(const orc_fftproc *p, int idx);
const float *orc_fftproc_peaks(const orc_fftproc *p);
int orc_fftproc_read_index(const orc_fftproc *p);
int orc_fftproc_write_index(const orc_fftproc *p);
int orc_fftproc_rows(const orc_fftproc *p);
/* FftProcessor.kt:143-157; returns 1 and sets *out if the channel covers >=1 bin */
int orc_signal_strength(const float *mag, int N, long long frequency, int sampleRate,
                        long long chanStart, long long chanEnd, float *out);
/* AnalyzerSurface.kt:710-714 for the 1 bin == 1 pixel case: mean of newest L+1 rows,
 * summed newest->oldest in float32 */
void orc_time_average(const orc_fftproc *p, int L, float *avg);
/* Exponential average of `frames` rows (time order), an option the REFERENCE DOES NOT HAVE (its average is the
 * box-car above); this is the definition the GPU option RFA_AVG_EMA is held to: a = a + alpha*(row - a), float32,
 * every operation rounded; init == NULL starts at the first row. */
void orc_ema_rows(const float *rows, long long frames, int N, float alpha, const float *init, float *avg);
/* AnalyzerSurface.kt:599-743 arithmetic (per-pixel mean, time average, colour index) */
void orc_draw_preprocess(const orc_fftproc *p, int width, int fftHeight,
                         long long viewportFrequency, long long viewportSampleRate,
                         float minDB, float maxDB, int L, int colorMapSize,
                         float *timeAverage /*width, NaN where not drawn*/,
                         int *colorIndex /*rows*width, -1 = black*/,
                         float *peaksY /*width or NULL*/);

/* ---- window functions / filter design ---------------------------------- */
float orc_window_value(int kind, double beta, int n, int N);
/* FirFilter.createLowPassTaps; returns ntaps (0 on firdes check failure); taps malloc'd */
int orc_lowpass_taps(float gain, float sampleRate, float cutoff, float transitionWidth,
                     float attenuation, int windowKind, double beta, int maxTaps, float **taps);
int orc_bandpass_taps(float gain, float fs, float lo, float hi, float tw, float att,
                      float **tapsRe, float **tapsIm);

/* ---- FirFilter / ComplexFirFilter -------------------------------------- */
typedef struct orc_fir orc_fir;
orc_fir *orc_fir_new(const float *taps, int ntaps, int decimation);
orc_fir *orc_fir_lowpass(int decimation, float gain, float fs, float cutoff, float tw, float att);
void orc_fir_free(orc_fir *f);
int orc_fir_ntaps(const orc_fir *f);
const float *orc_fir_taps(const orc_fir *f);
int orc_fir_filter(orc_fir *f, const orc_packet *in, orc_packet *out, int offset, int length);
int orc_fir_filter_real(orc_fir *f, const orc_packet *in, orc_packet *out, int offset, int length);

typedef struct orc_cfir orc_cfir;
orc_cfir *orc_cfir_bandpass(int decimation, float gain, float fs, float lo, float hi, float tw, float att);
void orc_cfir_free(orc_cfir *f);
int orc_cfir_ntaps(const orc_cfir *f);
const float *orc_cfir_taps_re(const orc_cfir *f);
const float *orc_cfir_taps_im(const orc_cfir *f);
int orc_cfir_filter(orc_cfir *f, const orc_packet *in, orc_packet *out, int offset, int length);

/* ---- RationalResampler -------------------------------------------------- */
int orc_gcd(int a, int b);
void orc_limit_denominator(int num, int den, int maxDen, int *outNum, int *outDen);
int orc_design_resampler_taps(int interp, int decim, float fractionalBw, int maxTaps, float **taps);
typedef struct orc_resampler orc_resampler;
orc_resampler *orc_resampler_new(int interp, int decim, const float *taps, int ntaps,
                                 float fractionalBw, int maxTaps);
void orc_resampler_free(orc_resampler *r);
int orc_resampler_interp(const orc_resampler *r);
int orc_resampler_decim(const orc_resampler *r);
int orc_resampler_taps_per_phase(const orc_resampler *r);
/* copies the polyphase bank [interp][tapsPerPhase] */
void orc_resampler_bank(const orc_resampler *r, float *bank);
int orc_resampler_resample(orc_resampler *r, const orc_packet *in, orc_packet *out, int offset, int length);

/* ---- Demodulator + AudioSink filters ------------------------------------ */
typedef struct orc_demod orc_demod;
orc_demod *orc_demod_new(int packetSize);
void orc_demod_free(orc_demod *d);
void orc_demod_set_mode(orc_demod *d, int mode);      /* also resets channel width to default */
void orc_demod_set_channel_width(orc_demod *d, int w); /* coerced to mode min/max */
int orc_demod_channel_width(const orc_demod *d);
void orc_demod_set_volume(orc_demod *d, float v);
int orc_mode_quadrature_rate(int mode);
/* one Demodulator.run iteration on a resampled packet: user filter, demod, volume.
 * audio->re receives the result at the quadrature (or SSB: audio) rate */
void orc_demod_process(orc_demod *d, const orc_packet *resampled, orc_packet *audio);
/* stage access for per-stage parity */
void orc_demod_user_filter(orc_demod *d, const orc_packet *in, orc_packet *out);
void orc_demod_fm(orc_demod *d, const orc_packet *in, orc_packet *out, float maxDeviation);
void orc_demod_am(orc_demod *d, const orc_packet *in, orc_packet *out);
void orc_demod_ssb(orc_demod *d, const orc_packet *in, orc_packet *out, int upperBand);
void orc_demod_cw(orc_demod *d, const orc_packet *in, orc_packet *out);

typedef struct orc_audiosink orc_audiosink;
orc_audiosink *orc_audiosink_new(int packetSize, int sampleRate);
void orc_audiosink_free(orc_audiosink *a);
/* AudioSink.applyAudioFilter; returns 0 if the ratio is unsupported */
int orc_audiosink_filter(orc_audiosink *a, const orc_packet *in, orc_packet *out);

/* ---- whole chains over a recording (what Scheduler+threads do, loss-free) */
/* spectrum: frames of N samples -> rows[F][N] dB, peaks[N], avg[N] (newest L+1 rows).
 * ringRows==0 -> linear rows. Returns frames processed. Uses orc_fft_* (restated FFT). */
long long orc_spectrum_run(int fmt, const uint8_t *iq, long long nsamples, int N, int L,
                           float *rows, float *peaks, float *avg);
/* demod chain: returns number of 48 kHz (or pass-through rate) audio samples written */
long long orc_chain_run(int fmt, const uint8_t *iq, long long nsamples, int sampleRate,
                        long long srcFrequency, long long channelFrequency, int mode,
                        int channelWidth, int packetSamples, float volume,
                        float *audio, long long audioCapacity);

/* ---- synthetic IQ generator (tests/bench inputs) -- */
/* All-integer, so a device-side generator can reproduce it bit for bit.
 *   h      = fmix32(seed ^ (u32)n ^ ((u32)(n>>32) * 0x9E3779B9))
 *   noise  = 8-bit: (int8)(h&0xFF)>>noiseShift, (int8)((h>>8)&0xFF)>>noiseShift
 *            16-bit: (int16)(h&0xFFFF)>>noiseShift, (int16)(h>>16)>>noiseShift
 *   comp k : phase = (u32)(n*step) + (u32)(((int64)modK * tab[(u32)(n*modStep)>>20]))
 *            I += (amp*tab[phase>>20] + 8192) >> 14 ; Q uses phase - 2^30 (sine)
 *   tab[j] = lround(16384*cos(2*pi*j/4096))
 * s8: byte = value, u8: byte = value + 128, s16le: little-endian int16. */
typedef struct {
    uint32_t step;
    int32_t amp;
    uint32_t modStep;
    int32_t modK;
} orc_synth_comp;
void orc_synth_iq(int fmt, uint32_t seed, const orc_synth_comp *comps, int ncomp, int noiseShift,
                  long long firstSample, long long nsamples, uint8_t *out);
/* the three-tone spectrum-path signal of SURVEY.md 8(d); returns ncomp (3) */
int orc_synth_default_comps(int fmt, orc_synth_comp *comps);
uint32_t orc_synth_step(double cyclesPerSample);

#ifdef __cplusplus
}
#endif

/* ---- Airspy real -> IQ converter (libairspy/src/main/cpp/libairspy/iqconverter_int16.c:54-208), restated.
 * Integer arithmetic throughout: DC removal with error feedback (:160-186), fs/4 translation (:188-202), the
 * half-band FIR on the even samples (:97-134, kernel = every other tap), a len/4-sample delay on the odd ones
 * (:136-158).  The reference clears only half of its delay line on reset (:94); the restatement starts from zeros. */
typedef struct orc_iqconv orc_iqconv;
orc_iqconv *orc_iqconv_new(const int16_t *hb_kernel, int len);
void orc_iqconv_free(orc_iqconv *c);
void orc_iqconv_reset(orc_iqconv *c);
void orc_iqconv_process(orc_iqconv *c, int16_t *samples, long long len);
/* airspy.c:299-309 convert_samples_int16: (raw - 2048) << 4 */
void orc_airspy_convert_samples(const uint16_t *src, int16_t *dst, long long count);

#endif
