/*
 * rfa_b200.h -- C ABI of librfa_b200.so, the B200 (sm_100a) implementation of
 * RF Analyzer's IQ->spectrum and IQ->audio hot path.
 *
 * Plain pointers and sizes only; every function returns an int status (RFA_OK == 0)
 * and rfa_last_error() describes the last failure on the calling thread.  There is no
 * CPU fallback: without a CUDA device every compute entry point fails with RFA_ERR_CUDA.
 *
 * Each entry point names the reference interface it replaces.  Paths are relative to the
 * reference tree; A/ = app/src/main/java/com/mantz_it/rfanalyzer/.
 *
 * Memory: every data pointer of a call lives in ONE space, given by `mem`:
 *   RFA_MEM_HOST   -- host memory; the library stages H2D/D2H copies on the context's
 *                     stream and the call returns when the results are in the buffers.
 *   RFA_MEM_DEVICE -- device memory of the context's GPU; the call is asynchronous on
 *                     the context's stream (use rfa_ctx_sync or your own stream sync).
 * Sample buffers must be naturally aligned for their element (2 B for 8-bit IQ pairs,
 * 4 B for 16-bit IQ pairs and floats, 8 B for interleaved complex float).
 */
#ifndef RFA_B200_H
#define RFA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFA_VERSION 100

enum { RFA_OK = 0, RFA_ERR_INVALID = 1, RFA_ERR_CUDA = 2, RFA_ERR_UNSUPPORTED = 3, RFA_ERR_NOMEM = 4 };
enum { RFA_MEM_HOST = 0, RFA_MEM_DEVICE = 1 };

/* IQ sample formats = the three IQConverter subclasses
 * (A/source/Signed8BitIQConverter.java, Unsigned8BitIQConverter.java, Signed16BitIQConverter.kt) */
enum { RFA_FMT_S8 = 0, RFA_FMT_U8 = 1, RFA_FMT_S16LE = 2 };

/* FFT windows. BLACKMAN_REF is the only one the reference has
 * (nativedsp/src/main/java/com/mantz_it/nativedsp/NativeDsp.kt:14-21). */
enum { RFA_WIN_BLACKMAN_REF = 0, RFA_WIN_HANN = 1, RFA_WIN_RECT = 2 };

/* tap-design windows of A/dsp/WindowFunctions.kt:44-100 */
enum { RFA_TAPWIN_BLACKMAN = 0, RFA_TAPWIN_HAMMING = 1, RFA_TAPWIN_KAISER = 2 };

/* A/ui/composable/DemodulationTab.kt:90-99 (ordinal order) */
enum { RFA_MODE_OFF = 0, RFA_MODE_AM, RFA_MODE_NFM, RFA_MODE_WFM, RFA_MODE_LSB, RFA_MODE_USB, RFA_MODE_CW };

/* flags for the FIR / resampler / demodulator kernels */
enum {
    RFA_SUM_FMA = 0,   /* fused multiply-add accumulation: fastest, within 1e-6 relative of the reference */
    RFA_SUM_EXACT = 1  /* multiply and add rounded separately, in the reference's tap order: bit-exact */
};

typedef struct rfa_ctx rfa_ctx;
typedef struct rfa_spectrum_plan rfa_spectrum_plan;
typedef struct rfa_fir rfa_fir;
typedef struct rfa_resampler rfa_resampler;
typedef struct rfa_chain rfa_chain;

/* ---- library / context --------------------------------------------------------------- */
int rfa_version(void);
const char *rfa_last_error(void);
/* stream: a cudaStream_t to run on (e.g. torch's current stream), or NULL to create one */
int rfa_ctx_create(int device, void *stream, rfa_ctx **out);
int rfa_ctx_destroy(rfa_ctx *ctx);
int rfa_ctx_sync(rfa_ctx *ctx);
int rfa_ctx_device(const rfa_ctx *ctx);
int rfa_ctx_sm_count(const rfa_ctx *ctx);
void *rfa_ctx_stream(const rfa_ctx *ctx);
/* number of kernels this context has launched (bench.py's gpu_launches) */
long long rfa_ctx_launch_count(const rfa_ctx *ctx);
/* Knobs of a context (csrc/tuning.h lists them: "staged", "pdl", "max_grid", "fs_batch_kib", "fs_tma", "fs_ztma",
 * "fs_pdl", "chunk_kib", "rs_span", "cluster").  The defaults are the product; the knobs exist for A/B timing
 * runs and for tests that compare two code paths.  Nothing in the library reads the environment.  Unknown names
 * (including the experimental-kernel options that only librfa_b200_lab.so carries) return RFA_ERR_UNSUPPORTED. */
int rfa_ctx_set_option(rfa_ctx *ctx, const char *name, long long value);
int rfa_ctx_get_option(rfa_ctx *ctx, const char *name, long long *value);
/* pinned host memory, so RFA_MEM_HOST calls copy at full PCIe rate and overlap */
int rfa_host_alloc(size_t bytes, void **out);
int rfa_host_free(void *p);

/* ---- IQ conversion: IQConverter.fillPacketIntoSamplePacket --------------------------- */
/* A/source/Signed8BitIQConverter.java:80-99, Unsigned8BitIQConverter.java:80-99,
 * Signed16BitIQConverter.kt:89-124.  iq: nsamples interleaved I,Q pairs -> planar re, im.
 * Bit-exact with the reference's look-up tables. */
int rfa_convert(rfa_ctx *ctx, int fmt, const void *iq, long long nsamples, float *re, float *im, int mem);

/* ---- NCO mixer: IQConverter.mixPacketIntoSamplePacket -------------------------------- */
/* Host-side table design = generateMixerLookupTable + calcOptimalCosineLength
 * (A/source/IQConverter.java:64-76, Signed8BitIQConverter.java:54-77,
 * Signed16BitIQConverter.kt:59-87).  mix_frequency = (int)(source - channel frequency).
 * Writes the effective frequency (after the "+= sampleRate" rule), the table length
 * (<= 500) and cos/sin tables of that length (capacity >= 500 floats each). */
int rfa_nco_design(int fmt, int sample_rate, int mix_frequency, int *effective_frequency, int *length,
                   float *cos_table, float *sin_table);
/* re = v(I)*cos - v(Q)*sin, im = v(Q)*cos + v(I)*sin with every product rounded to float
 * first (the reference's 2-D product tables); sample n uses table index
 * (nco_index + n) % nco_length.  cos/sin tables are HOST pointers in both modes. */
int rfa_mix(rfa_ctx *ctx, int fmt, const void *iq, long long nsamples, const float *cos_table,
            const float *sin_table, int nco_length, int nco_index, float *re, float *im, int mem);

/* ---- NativeDsp ------------------------------------------------------------------------ */
/* NativeDsp.makeWindow (NativeDsp.kt:14-21) and variants; host function. */
int rfa_make_window(int window, int n, float *out);
/* Java_com_mantz_1it_nativedsp_NativeDsp_performFFT (nativedsp/src/main/cpp/nativedsp.cpp:19-42):
 * ordered forward complex FFT, interleaved float32, unnormalised; `batch` transforms. */
int rfa_fft_c2c(rfa_ctx *ctx, const float *in, float *out, int n, long long batch, int mem);
/* Java_..._performFFTAndLogMag (nativedsp.cpp:44-81): FFT, 10*log10(|X|/n), fft-shifted. */
int rfa_fft_logmag(rfa_ctx *ctx, const float *in, float *mag, int n, long long batch, int mem);
/* NativeDsp.performWindowedFftAndReturnMag (NativeDsp.kt:43-62): planar re/im in, dB out. */
int rfa_windowed_fft_logmag(rfa_ctx *ctx, const float *re, const float *im, float *mag, int n,
                            long long batch, int window, int mem);

/* ---- fused spectrum path --------------------------------------------------------------
 * Scheduler.kt:266 (fill) + NativeDsp.kt:43-62 + nativedsp.cpp:44-81 +
 * FftProcessor.kt:224-245 (row store, peak hold) + AnalyzerSurface.kt:710-714 (time average)
 * in one pass over the IQ bytes. */
/* How `avg` is formed.  BOXCAR is the reference (AnalyzerSurface.kt:710-714: mean of the newest avg_len+1 rows) and
 * the default.  EMA is an extra option the reference does not have: a_k = a_(k-1) + alpha*(row_k - a_(k-1)) over the
 * frames in time order, float32, each operation rounded (no FMA); the first frame starts it (a_0 = row_0) unless
 * rfa_spectrum_out.avg_accumulate continues from the value already in `avg`.  The kernel walks only as many of the
 * newest frames as carry weight above 2^-40 (and are still stored), so the result equals the full recurrence to
 * far below one float32 ulp; a run of -inf rows (all-zero frames) poisons the average exactly as the recurrence says. */
enum { RFA_AVG_BOXCAR = 0, RFA_AVG_EMA = 1 };

typedef struct {
    int format;    /* RFA_FMT_* */
    int fft_size;  /* power of two, 16 .. 65536 (the app offers 1024 .. 65536, DisplayTab.kt:108-111) */
    int window;    /* RFA_WIN_* */
    int avg_len;   /* L = fftAverageLength, 0 .. 30: the average spans the newest L+1 rows */
    int peak_hold; /* FftProcessor.fftPeakHold */
    int avg_mode;  /* RFA_AVG_BOXCAR (0, the reference) or RFA_AVG_EMA */
    float ema_alpha; /* RFA_AVG_EMA: weight of the newest row, 0 < alpha <= 1 */
} rfa_spectrum_desc;

typedef struct {
    float *rows;          /* dB rows, or NULL to skip the waterfall store */
    long long row0;       /* frame f goes to row (row0 + f*row_step) mod ring_rows           */
    long long row_step;   /* +1: linear [frames][n]; -1: the reference's backwards ring      */
    long long ring_rows;  /* 0 = no wrap (linear); 300/400/500 = FftProcessor.kt:104 rings   */
    long long row_stride; /* floats between rows (>= fft_size)                                */
    long long history_rows; /* ring only: rows already valid before this call (for avg)      */
    float *peaks;         /* [n] running element-wise max of all rows, or NULL               */
    int peaks_accumulate; /* 1: continue from the values in `peaks`; 0: restart at -999999f  */
    float *avg;           /* [n] mean of the newest avg_len+1 rows (or the EMA), or NULL      */
    int avg_accumulate;   /* RFA_AVG_EMA: 1 = continue from the values in `avg`              */
} rfa_spectrum_out;

int rfa_spectrum_plan_create(rfa_ctx *ctx, const rfa_spectrum_desc *desc, rfa_spectrum_plan **out);
int rfa_spectrum_plan_destroy(rfa_spectrum_plan *plan);
int rfa_spectrum_plan_info(const rfa_spectrum_plan *plan, int *fft_size, int *format, int *avg_len);
/* iq: nframes * fft_size samples, frames contiguous and non-overlapping (Scheduler.kt:254-276). */
int rfa_spectrum_process(rfa_spectrum_plan *plan, const void *iq, long long nframes,
                         const rfa_spectrum_out *out, int mem);
/* algorithmic HBM bytes of one process call: nframes*n*(bytes_in + 4) + 8*n */
long long rfa_spectrum_algorithmic_bytes(const rfa_spectrum_plan *plan, long long nframes, int rows_stored);

/* ---- reductions over waterfall rows ---------------------------------------------------- */
/* AnalyzerSurface.kt:683-684,710-714: avg = (sum of rows newest, newest+dir, ... L+1 terms,
 * summed in that order in float32) / (L+1); terms past `valid` count as -9999f. */
int rfa_average_rows(rfa_ctx *ctx, const float *rows, long long newest, long long dir, long long ring_rows,
                     long long row_stride, long long valid, int avg_len, int n, float *avg, int mem_rows,
                     int mem_avg);
/* RFA_AVG_EMA over frames first .. last (inclusive, time order) of device rows laid out as in rfa_spectrum_out
 * (row of frame f = (row0 + f*row_step) mod ring_rows); from_state: continue from `avg` instead of starting at
 * the first row.  avg per mem_avg. */
int rfa_ema_rows(rfa_ctx *ctx, const float *rows, long long row0, long long row_step, long long ring_rows,
                 long long row_stride, long long first, long long last, float alpha, int from_state, int n,
                 float *avg, int mem_avg);
/* FftProcessor.kt:143-157: host helper for the channel's bin range ... */
int rfa_channel_bins(int n, long long frequency, int sample_rate, long long chan_start, long long chan_end,
                     int *bin_start, int *bin_end);
/* ... and the mean dB over [bin_start, bin_end) of nrows rows (device rows, `out` per mem_out). */
int rfa_channel_strength(rfa_ctx *ctx, const float *rows, long long row0, long long row_step,
                         long long ring_rows, long long row_stride, long long nrows, int bin_start,
                         int bin_end, float *out, int mem_out);
/* FftProcessor.kt:199-217: shift `nrows` device rows by `shift` bins, fill with -9999f. */
int rfa_shift_rows(rfa_ctx *ctx, float *rows, long long nrows, long long row_stride, int n, int shift);

/* ---- signal detectors on waterfall rows --------------------------------------
 * ui/MainViewModel.kt: getAverageSignalLevel (:1392-1414), detectSignal (:1416-1461), detectSignalsInFFT
 * (:1463-1550), detectIEMChannelsInFFT (:861-935), detectAirCommSignal / ...AtFrequency (:1151-1250),
 * groupSignals / finalizeGroup (:1552-1607); ScanDetectionMode ui/composable/ScanTab.kt:35-39.
 * Every detector reduces windows of bins of a dB row to peak = windowData.maxOrNull() (NaN-propagating float
 * max) and avg = windowData.average().toFloat() (double sum, one division, one rounding). */
enum { RFA_DETECT_PEAK_ONLY = 0, RFA_DETECT_AVERAGE_ONLY = 1, RFA_DETECT_PEAK_OR_AVERAGE = 2 };
typedef struct rfa_detect_window {
    long long row; /* row index into `rows` (row * row_stride floats from the base) */
    int start;     /* first bin, clamped to 0 */
    int end;       /* last bin INCLUSIVE, clamped to n-1 (sliceArray(windowStart..windowEnd)) */
} rfa_detect_window;
typedef struct rfa_signal { /* DiscoveredSignal: frequency, peakStrength, averageStrength, bandwidth, isGrouped */
    long long frequency;
    float peak;
    float average;
    long long bandwidth;
    int grouped;
} rfa_signal;
/* peak[i], avg[i] of window i over device-resident rows; windows per mem_win, outputs per mem_out.
 * One launch for any number of (row, window) pairs: a batch of rows is scanned without leaving HBM. */
int rfa_detect_windows(rfa_ctx *ctx, const float *rows, long long row_stride, int n, const rfa_detect_window *win,
                       int nwin, float *peak, float *avg, int mem_win, int mem_out);
/* host arithmetic of the detectors, with the JVM's conversions (Long -> Float, float division, toInt()):
 * binIndex = ((freq - (center - fs/2)) / (fs.toFloat() / n)).toInt() */
int rfa_detect_bin(long long center_freq, long long sample_rate, int n, long long freq);
/* windowHalfSize = (half_width_hz / (fs.toFloat() / n)).toInt().coerceAtLeast(min_half) */
int rfa_detect_half_width(long long sample_rate, int n, int half_width_hz, int min_half);
/* bin and clamped window of `freq`; half_width_hz < 0 = a fixed +-min_half bins.  Returns 1 when the bin
 * lies inside the row (`binIndex in currentFFT.indices`), 0 when the reference skips it. */
int rfa_detect_window_at(long long center_freq, long long sample_rate, int n, long long freq, int half_width_hz,
                         int min_half, int *bin, int *start, int *end);
/* detection mode against maxOf(threshold, noiseFloor + noiseFloorMargin): 1 detected, 0 not, -1 bad mode */
int rfa_detect_decide(float peak, float avg, float threshold, float noise_floor, float margin, int mode);
/* the frequency grid of detectSignalsInFFT and its +-2-bin windows on row `row`; writes at most `cap` entries,
 * returns the number of grid points inside the row (call with cap = 0 to size the arrays), -1 on bad input */
long long rfa_scan_grid(long long center_freq, long long sample_rate, long long usable_bandwidth, long long step,
                        long long scan_start, long long scan_end, int n, long long row, long long *freqs,
                        rfa_detect_window *win, long long cap);
/* groupSignals: sort by frequency, merge neighbours closer than step * minimum_gap; `out` holds up to n
 * signals, the count is returned */
long long rfa_group_signals(const rfa_signal *in, long long n, long long step, int minimum_gap, rfa_signal *out);

/* ---- waterfall / FFT-trace preprocessing -------------------------------------
 * AnalyzerSurface.drawPreprocessing (ui/AnalyzerSurface.kt:646-734), arithmetic only: per pixel the mean of
 * the row's bins (:703-713), colour-map index int((avg-minDB)*mapSize/(maxDB-minDB)) clamped (:726-727),
 * black outside the frame (:728-731), the FFT trace = mean over the newest avg_len+1 rows of those means
 * (:716-719), the peak trace y coordinates (:714).  Bit-identical to the JVM (same float32 order).
 * `rows` is the device-resident ring the spectrum plan fills (newest_row = FftProcessorData.readIndex), so a
 * display client reads back `width` pixels per row instead of fft_size floats. */
typedef struct {
    int fft_size;                 /* bins per row */
    long long frequency;          /* FftProcessorData.frequency of the rows */
    int sample_rate;              /* FftProcessorData.sampleRate */
    long long viewport_frequency, viewport_sample_rate;
    int width, fft_height;        /* pixels */
    float min_db, max_db;         /* viewportVerticalScale */
    int avg_len;                  /* fftAverageLength */
    int ring_rows;                /* waterfallBuffer.size */
    long long row_stride;         /* floats between rows, 0 = fft_size */
    int newest_row;               /* ring index of the newest row (currentRowIdx) */
    int first_row, nrows;         /* rowNumber range to render, 0 = newest (dirty rows only, :688-691) */
} rfa_render_desc;
/* argb / color_index: [ring_rows][width] indexed by ring row like colorBuffer (color_index = -1 outside the
 * frame); time_average / peaks_y: [width] (time_average needs first_row = 0 and nrows > avg_len; NaN outside
 * the frame).  Any output may be NULL.  out_mem says where the outputs live, colormap_mem where the map lives. */
int rfa_render_waterfall(rfa_ctx *ctx, const rfa_render_desc *desc, const float *rows, const float *peaks,
                         const uint32_t *colormap, int colormap_size, int colormap_mem, uint32_t *argb,
                         int *color_index, float *time_average, float *peaks_y, int out_mem);
int rfa_fill(rfa_ctx *ctx, float *dst, long long count, float value);

/* ---- filter design (host functions; float/double usage follows the reference) ---------- */
/* WindowFunction.value (A/dsp/WindowFunctions.kt:44-100) */
int rfa_tap_window(int kind, double beta, int n, int N, float *out);
/* FirFilter.createLowPassTaps (A/dsp/FirFilter.kt:182-241).  taps may be NULL to query ntaps. */
int rfa_design_lowpass(float gain, float sample_rate, float cutoff, float transition_width, float attenuation_db,
                       int window, double beta, int max_taps, float *taps, int capacity, int *ntaps);
/* ComplexFirFilter.createBandPass (A/dsp/ComplexFirFilter.java:186-262) */
int rfa_design_bandpass(float gain, float sample_rate, float low_cutoff, float high_cutoff, float transition_width,
                        float attenuation_db, float *taps_re, float *taps_im, int capacity, int *ntaps);
/* RationalResampler.limitDenominator / designResamplerTaps (A/dsp/RationalResampler.kt:183-255) */
int rfa_limit_denominator(int numerator, int denominator, int max_denominator, int *out_num, int *out_den);
int rfa_design_resampler_taps(int interpolation, int decimation, float fractional_bw, int max_taps, float *taps,
                              int capacity, int *ntaps);

/* ---- FirFilter / ComplexFirFilter (A/dsp/FirFilter.kt:34-163, ComplexFirFilter.java:33-170) -
 * Streaming objects: state (delay line, decimationCounter starting at 1) persists across
 * calls exactly like the reference's fields.  taps_im != NULL makes it a ComplexFirFilter.
 * process(): in_im == NULL is filterReal.  Appends nothing itself: out_* receive *n_out
 * samples (at most out_capacity); *consumed follows the reference's return value (less than
 * n only when the output ran full). */
int rfa_fir_create(rfa_ctx *ctx, const float *taps_re, const float *taps_im, int ntaps, int decimation, int flags,
                   rfa_fir **out);
int rfa_fir_destroy(rfa_fir *fir);
int rfa_fir_reset(rfa_fir *fir);
int rfa_fir_process(rfa_fir *fir, const float *in_re, const float *in_im, long long n, float *out_re,
                    float *out_im, long long out_capacity, long long *n_out, long long *consumed, int mem);

/* ---- RationalResampler (A/dsp/RationalResampler.kt:36-156) -------------------------------
 * taps == NULL designs them (designResamplerTaps, Kaiser beta 7).  I/D are reduced by
 * their gcd like the constructor does. */
int rfa_resampler_create(rfa_ctx *ctx, int interpolation, int decimation, const float *taps, int ntaps,
                         float fractional_bw, int max_taps, int flags, rfa_resampler **out);
int rfa_resampler_destroy(rfa_resampler *rs);
int rfa_resampler_info(const rfa_resampler *rs, int *interpolation, int *decimation, int *taps_per_phase);
int rfa_resampler_process(rfa_resampler *rs, const float *in_re, const float *in_im, long long n, float *out_re,
                          float *out_im, long long out_capacity, long long *n_out, long long *consumed, int mem);

/* ---- Demodulator stages (A/analyzer/Demodulator.kt:251-403), one packet per call ---------- */
/* demodulateFM: carry[2] (host, in/out) = last sample of the previous packet; the result is
 * multiplied by `volume` (Demodulator.run :184-187). */
int rfa_demod_fm(rfa_ctx *ctx, const float *re, const float *im, long long n, float *carry,
                 float quadrature_gain, float volume, float *out, int flags, int mem);
/* demodulateAM: |x|^2, minus the packet mean, times 0.75/lastMax; *last_max (host) in/out. */
int rfa_demod_am(rfa_ctx *ctx, const float *re, const float *im, long long n, float *last_max, float volume,
                 float *out, int flags, int mem);
/* the gain control at the end of demodulateSSB / demodulateCW, in place on x */
int rfa_agc(rfa_ctx *ctx, float *x, long long n, float *last_max, float volume, int flags, int mem);
/* mode table: quadrature rate (Demodulator.kt:53-62) and channel widths (DemodulationTab.kt:90-99) */
int rfa_mode_info(int mode, int *quadrature_rate, int *min_width, int *max_width, int *default_width);

/* ---- the whole IQ -> audio chain ------------------------------------------------------------
 * Scheduler.kt:237-244 (mixPacketIntoSamplePacket) -> Resampler.kt:95-113 -> Demodulator.kt:147-187
 * (user filter, demodulator, volume) -> AudioSink.java:182-187 (decimation to 48 kHz), every packet
 * delivered.  State carries across calls; every call but the last must be a whole number of packets.
 * RFA_MEM_HOST calls return with the audio in the caller's buffer.  RFA_MEM_DEVICE calls of RFA_SUM_FMA chains only
 * enqueue work on the context's stream (*n_audio is known on return, the samples are there after rfa_ctx_sync or any
 * later work on that stream); RFA_SUM_EXACT AM / SSB / CW calls synchronise before they return. */
typedef struct {
    int format;                  /* RFA_FMT_* */
    int sample_rate;             /* source sample rate, Hz */
    long long source_frequency;  /* IQSource frequency */
    long long channel_frequency; /* Scheduler.channelFrequency (CW: caller adds the 750 Hz offset, AnalyzerService.kt:439-442) */
    int mode;                    /* RFA_MODE_AM .. RFA_MODE_CW */
    int channel_width;           /* Hz, coerced to the mode's range; 0 = the mode's default */
    int packet_samples;          /* source.packetSize / bytesPerSample (AnalyzerService.kt:324-328) */
    float volume;                /* audioVolumeLevel */
    int flags;                   /* RFA_SUM_FMA / RFA_SUM_EXACT */
} rfa_chain_desc;
int rfa_chain_create(rfa_ctx *ctx, const rfa_chain_desc *desc, rfa_chain **out);
int rfa_chain_destroy(rfa_chain *chain);
int rfa_chain_info(const rfa_chain *chain, int *interpolation, int *decimation, int *taps_per_phase,
                   int *quadrature_rate, int *channel_width, int *nco_length, int *nco_frequency);
/* audio buffer capacity (samples) that rfa_chain_process needs for nsamples inputs */
long long rfa_chain_max_audio(const rfa_chain *chain, long long nsamples);
int rfa_chain_process(rfa_chain *chain, const void *iq, long long nsamples, float *audio, long long capacity,
                      long long *n_audio, int mem);
/* Time-sharding of a long recording (not in the reference, which runs one sequential
 * Scheduler/Demodulator pipeline, Scheduler.kt:140-298): put every counter of the chain -- NCO table
 * index (IQConverter mix index), RationalResampler ctr/delay position, FirFilter.decimationCounter of each
 * filter -- where a run from sample 0 has it at `sample_index` (a packet boundary), empty the delay
 * lines, restart FM carry and AGC maximum.  *audio_index = audio samples produced before that point. */
int rfa_chain_seek(rfa_chain *chain, long long sample_index, long long *audio_index);

/* ---- IQ recordings on disk -------------------------------------------------------
 * IQ_FILE_FORMAT.md: header-less interleaved IQ; the metadata lives in the file NAME
 * "{yyyyMMdd-HHmmss}_{name}_{HACKRF|RTLSDR|AIRSPY|HYDRASDR}_{frequency}_{sampleRate}.iq" (RecordingDao.kt:87-90). */
enum { RFA_FILE_HACKRF = 0, RFA_FILE_RTLSDR = 1, RFA_FILE_AIRSPY = 2, RFA_FILE_HYDRASDR = 3 }; /* FilesourceFileFormat */
typedef struct {
    int file_format;       /* RFA_FILE_* */
    long long frequency;   /* Hz */
    long long sample_rate; /* samples/s */
    int have_format, have_frequency, have_sample_rate; /* set when the name carried the field */
} rfa_recording_info;
/* MainViewModel.setFilesourceUri (ui/MainViewModel.kt:2034-2080): fields the name does not carry keep the values
 * `info` came in with, exactly as the app keeps its current settings. */
int rfa_recording_parse_name(const char *filename, rfa_recording_info *info);
/* Recording.calculateFileName + Long.asStringWithUnit (RecordingDao.kt:87-90, HelperComposables.kt:168-179);
 * `timestamp` is the already formatted "yyyyMMdd-HHmmss" string. */
int rfa_recording_file_name(const char *timestamp, const char *name, int file_format, long long frequency,
                            long long sample_rate, char *out, int capacity);
/* RFA_FILE_* -> RFA_FMT_* (IQ_FILE_FORMAT.md:26-84), -1 if unknown */
int rfa_recording_sample_format(int file_format);

/* FileIQSource (source/FileIQSource.java:64-91,305-369): getPacket() hands out whole packets only, rewinds when
 * `repeat` is set, and -- with pace_sample_rate > 0 -- sleeps so that packets arrive at the hardware's rate. */
typedef struct rfa_file_source rfa_file_source;
int rfa_file_source_open(const char *path, int file_format, long long packet_bytes, int repeat,
                         long long pace_sample_rate, rfa_file_source **out);
/* 1: a packet was written to `packet`; 0: end of file; < 0: error */
int rfa_file_source_get_packet(rfa_file_source *source, void *packet);
long long rfa_file_source_bytes_read(const rfa_file_source *source);
int rfa_file_source_close(rfa_file_source *source);

/* Spectrum pass over frames [first_frame, first_frame + nframes) of a recording (nframes < 0: to its end) with
 * a reader thread and two pinned buffers feeding rfa_spectrum_process: the disk read of chunk i+1 overlaps the
 * H2D copy / kernel / D2H copy of chunk i.  out->rows (host, all rows, or NULL), out->peaks, out->avg as in
 * rfa_spectrum_process with host buffers; chunk_frames <= 0 picks 32 MiB of IQ per chunk. */
int rfa_spectrum_process_file(rfa_spectrum_plan *plan, const char *path, long long first_frame, long long nframes,
                              const rfa_spectrum_out *out, long long chunk_frames, long long *frames_done);

/* ---- Scheduler.run as a batched runtime ------------------------------------------------
 * A/analyzer/Scheduler.kt:140-298: per packet the squelch debounce (:161-165, SQUELCH_DEBOUNCE_COUNT = 50 :52), the
 * recording gate (:199), the demodulator gate and mixPacketIntoSamplePacket (:237-250), the FFT buffer fill (:254-276);
 * FftProcessor.run's channel strength (FftProcessor.kt:143-157) over [channel - width, channel + width]
 * (AnalyzerService.kt:344-353) and squelchSatisfied = strength > squelch (AppStateRepository.kt:318-323).
 * A call carries many packets; the loop is replayed loss-free and synchronously (csrc/scheduler.cu): no packet is
 * dropped for lack of a buffer, a frame completed by packet k has set squelchSatisfied before packet k+1.  The
 * waterfall ring, peak hold and average live on the device (rfa_scheduler_state; rfa_render_waterfall reads them). */
typedef struct rfa_scheduler rfa_scheduler;
typedef struct {
    int format;                  /* RFA_FMT_* of the source's packets */
    int sample_rate;             /* source.sampleRate */
    long long source_frequency;  /* source.frequency */
    int packet_samples;          /* source.packetSize / source.bytesPerSample (Scheduler.kt:92-94) */
    int fft_size, window, avg_len, peak_hold; /* as rfa_spectrum_desc */
    int ring_rows;               /* FftProcessor.kt:104: 300 / 400 / 500 */
    int demodulation_mode;       /* RFA_MODE_OFF: isDemodulationActivated = false; else the Demodulator's mode */
    long long channel_frequency; /* Scheduler.channelFrequency */
    int channel_width;           /* Demodulator.channelWidth (0 = the mode's default) */
    float volume;
    int flags;                   /* RFA_SUM_FMA / RFA_SUM_EXACT for the chain */
    int squelch_enabled;         /* AppStateRepository.squelchEnabled */
    float squelch_db;            /* AppStateRepository.squelch */
    int record_only_when_squelch_satisfied; /* Scheduler.startRecording(onlyWhenSquelchIsSatisfied) */
} rfa_scheduler_desc;
typedef struct {
    long long packets;           /* out: packets consumed */
    long long frames;            /* out: FFT frames delivered to the FftProcessor */
    float *signal_strength;      /* host, >= frames this call can complete, or NULL: averageSignalStrength per frame */
    unsigned char *demod_gate;   /* host [npackets] or NULL: 1 = the packet went to the demodulator */
    unsigned char *record_gate;  /* host [npackets] or NULL: 1 = the packet would be written to the recording */
    float *audio;                /* 48 kHz audio of the delivered packets (same memory space as the packets) */
    long long audio_capacity;    /* >= rfa_chain_max_audio of the call's samples */
    long long n_audio;           /* out */
} rfa_scheduler_io;
int rfa_scheduler_create(rfa_ctx *ctx, const rfa_scheduler_desc *desc, rfa_scheduler **out);
int rfa_scheduler_destroy(rfa_scheduler *sched);
/* packets: npackets * packet_samples samples, back to back (host or device per `mem`) */
int rfa_scheduler_process(rfa_scheduler *sched, const void *packets, long long npackets, rfa_scheduler_io *io, int mem);
/* device pointers of the FftProcessor state and the scheduler's counters; any output may be NULL.
 * newest_row = FftProcessorData.readIndex (-1 before the first frame). */
int rfa_scheduler_state(const rfa_scheduler *sched, float **ring, long long *newest_row, long long *valid_rows,
                        float **peaks, float **avg, int *squelch_satisfied, int *debounce_counter, long long *packets,
                        long long *frames);

/* copies of the ring ([ring_rows][fft_size]), the peak hold and the average into HOST arrays; any may be NULL */
int rfa_scheduler_read(rfa_scheduler *sched, float *ring, float *peaks, float *avg);

/* ---- vendor real -> IQ converter ------------------------------------------------------
 * libairspy/src/main/cpp/libairspy/iqconverter_int16.c: iqconverter_int16_create (:54-78, hb_kernel = the 47-tap
 * HB_KERNEL_INT16 of filters.h:81-132, passed in by the caller exactly as airspy.c:920 does), _reset (:88-95),
 * _process (:204-208 = remove_dc :160-186 + translate_fs_4 :188-202 + fir_interleaved :97-134 + delay_interleaved
 * :136-158), _free (:80-86).  In place on `len` real int16 samples (len % 4 == 0): afterwards samples[2k], samples[2k+1]
 * are I and Q of IQ sample k at half the rate -- the int16 IQ format RFA_FMT_S16LE of the rest of this API.
 * Bit-exact with the reference for every input, across calls (all state is carried like the reference's struct);
 * the delay line starts from zeros (the reference's reset clears only half of it, :94). */
typedef struct rfa_iqconverter rfa_iqconverter;
int rfa_iqconverter_create(rfa_ctx *ctx, const int16_t *hb_kernel, int len, rfa_iqconverter **out);
int rfa_iqconverter_destroy(rfa_iqconverter *cnv);
int rfa_iqconverter_reset(rfa_iqconverter *cnv);
int rfa_iqconverter_process(rfa_iqconverter *cnv, int16_t *samples, long long len, int mem);
/* how the DC blocker's speculative chunks fared since the last reset: chunks run, chunks re-run sequentially, chunks
 * skipped in the blocker's dead zone (csrc/iqconv.cu explains the scheme) */
int rfa_iqconverter_stats(rfa_iqconverter *cnv, long long *chunks, long long *rerun, long long *dead_zone);
/* airspy.c:299-309 convert_samples_int16: raw ADC words -> (raw - 2048) << 4, the converter's input */
int rfa_airspy_convert_samples(rfa_ctx *ctx, const uint16_t *src, int16_t *dst, long long count, int mem);

/* ---- synthetic IQ (benchmark / test input; the reference ships no input fixtures) ------ */
/* All-integer generator of SURVEY.md 8(d): sample n depends on n alone, so any segment of a
 * long recording can be produced in place on any GPU.
 *   h = fmix32(seed ^ (u32)n ^ ((u32)(n>>32) * 0x9E3779B9)); noise = signed low bytes/halves >> noise_shift
 *   component k: phase = (u32)(n*step) + (u32)(mod_k * tab[((u32)(n*mod_step) - 2^30) >> 20]),
 *                I += (amp*tab[phase>>20] + 8192) >> 14, Q likewise with phase - 2^30,
 *   tab[j] = lround(16384*cos(2*pi*j/4096)); u8 adds 128; s16 is little endian. */
typedef struct {
    uint32_t step;     /* tone frequency, cycles/sample * 2^32 */
    int32_t amp;       /* LSB */
    uint32_t mod_step; /* FM: modulating tone, cycles/sample * 2^32 */
    int32_t mod_k;     /* FM: peak phase deviation / (2*pi) * 2^32 / 16384; 0 = plain tone */
} rfa_synth_comp;
int rfa_synth_iq(rfa_ctx *ctx, int fmt, uint32_t seed, const rfa_synth_comp *comps, int ncomp, int noise_shift,
                 long long first_sample, long long nsamples, void *out, int mem);

#ifdef __cplusplus
}
#endif
#endif /* RFA_B200_H */
