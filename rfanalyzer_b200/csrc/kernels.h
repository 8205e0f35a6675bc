// kernels.h -- launchers of the non-spectrum kernels (convert.cu, fir.cu, demod.cu).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace rfa {

cudaError_t convert_launch(int fmt, const void *iq, long long n, float *re, float *im, int num_sms,
                           cudaStream_t st);
cudaError_t mix_launch(int fmt, const void *iq, long long n, float *re, float *im, const float *cosT,
                       const float *sinT, int len, int idx, int num_sms, cudaStream_t st);

// a sample stream as the FIR / resampler kernels see it: index 0 is the first sample of this
// call, negative indices come from the `hist` samples carried over from earlier calls
struct StreamDesc {
    int kind = 3;                       // 0..2: raw IQ codes (FMT_*), converted (and mixed) on load; 3: planar floats
    const float *re = nullptr, *im = nullptr;
    const void *raw = nullptr;
    const float *hist_re = nullptr, *hist_im = nullptr;
    int hist = 0;
    const float *nco_cos = nullptr, *nco_sin = nullptr;  // device tables; nullptr = no mixing
    int nco_len = 1, nco_idx = 0;
};
// hist_new_* / consumed / hist_done: the delay line after the call; *hist_done = true when the resampler kernel wrote it
// itself (tiled and stripe kernels), false when the caller still has to launch history_launch
cudaError_t resample_launch(const StreamDesc &in, const float *bank, int I, int D, int nt, long long rel, int ph0,
                            long long nout, float *out_re, float *out_im, bool exact, cudaStream_t st, int rs_span = 0,
                            float *hist_new_re = nullptr, float *hist_new_im = nullptr, long long consumed = 0,
                            bool *hist_done = nullptr);
cudaError_t fir_launch(const StreamDesc &in, const float *taps_re, const float *taps_im, int ntaps, int dec,
                       long long first, long long nout, bool real_only, float *out_re, float *out_im, bool exact,
                       cudaStream_t st);
cudaError_t history_launch(const StreamDesc &in, long long consumed, float *new_re, float *new_im, cudaStream_t st);
cudaError_t demod_fm_launch(const float *re, const float *im, long long n, float *carry, float gain, float volume,
                            float *out, bool exact, int num_sms, cudaStream_t st);
cudaError_t demod_power_launch(const float *re, const float *im, long long n, float *out, int num_sms,
                               cudaStream_t st);
cudaError_t agc_launch(float *x, const long long *off, int npackets, long long max_packet, bool subtract_mean,
                       float *state, float *scratch, float volume, bool exact, int num_sms, cudaStream_t st);

// the delay lines of the user filter and both decimators slide in one launch
struct ChainStateArgs {
    struct Line {
        const float *in_re, *in_im, *old_re, *old_im;
        float *new_re, *new_im;
        long long n;  // inputs of this call
        int hist;     // 0 = unused
    } line[3];
};
cudaError_t chain_state_launch(const ChainStateArgs &a, cudaStream_t st);

// everything behind the resampler of an FM chain in one launch (chain_fused.cu)
struct FmTailArgs {
    const float *q_re, *q_im;              // quadrature samples of this call (resampler output)
    const float *hist_u_re, *hist_u_im;    // user filter delay line (user_hist samples before index 0)
    int user_hist, user_taps;
    const float *taps_user;
    long long first_u;                     // newest input index of user output 0 (decimationCounter, FirFilter.kt:46)
    long long nu;                          // user filter outputs = demodulator samples of this call
    const float *carry_in;                 // [2] last filtered sample of the previous call
    float *carry_out;                      // [2]
    float gain, volume;
    int ratio;                             // demodulated rate / 48 kHz: 1 (no decimator), 2 (first only), 8 (both)
    const float *taps_a1, *taps_a2;
    int a1_taps, a2_taps, a1_hist, a2_hist;
    const float *hist_a1, *hist_a2;        // delay lines of the decimators (demodulated / first-decimator samples)
    long long first_a1, first_a2, n1, n2;
    float *dem_out, *a1_out;               // intermediates kept for the next call's delay lines (may be NULL when unused)
    float *audio;                          // 48 kHz audio of this call
    // the call's bookkeeping inside the same launch (no chain_state_kernel behind it): one extra CTA slides the user
    // filter's delay line, the CTAs that own the newest demodulated / first-decimator samples write the decimators' next
    // delay lines (NULL: that decimator is not in the chain or saw no input)
    int slide_user;                        // 1: grid has one extra CTA for user_line
    ChainStateArgs::Line user_line;
    float *a1_hist_new, *a2_hist_new;
};
cudaError_t fm_tail_launch(const FmTailArgs &a, bool exact, cudaStream_t st);

// Packet boundaries of a chain call in the demodulated stream, in closed form: the resampler emits output j at input
// position rel + floor((ph + j*D)/I), a decimating filter emits at inputs first + m*dec, so the outputs that exist after n
// inputs are counts of those positions below n -- the same numbers the per-packet loop of rfa_chain_process accumulates
// (it checks them against each other).  off(p) = demodulated samples before packet p; off(npk) = all of the call.
struct PacketMap {
    long long packet_samples, nsamples;
    long long rel;            // resampler state before the call
    int ph, I, D;
    long long first_u;        // user filter (decimation 1)
    long long first_b;        // band-pass (dec_b > 0) behind it
    int dec_b;                // 0: none (AM)
    int npk;
#if defined(__CUDACC__)
    __host__ __device__
#endif
    long long off(long long p) const {
        long long n = p * packet_samples;
        if (n > nsamples) n = nsamples;
        long long q = 0;
        if (n > rel) {
            const long long num = (n - rel) * I - ph;
            q = (num + D - 1) / D;
        }
        long long u = q > first_u ? q - first_u : 0;
        if (dec_b > 0) u = u > first_b ? (u - 1 - first_b) / dec_b + 1 : 0;
        return u;
    }
};
cudaError_t packet_table_launch(const PacketMap &pm, long long *off /* [npk + 1], device */, cudaStream_t st);

// everything behind the resampler of an AM / SSB / CW chain in three launches (chain_agc.cu, RFA_SUM_FMA)
struct AgcTailArgs {
    const float *q_re, *q_im;              // quadrature samples of this call
    const float *hist_u_re, *hist_u_im;    // user filter delay line
    int user_hist, user_taps;
    const float *taps_user;
    long long first_u, nu;                 // as FmTailArgs
    const float *taps_b_re, *taps_b_im;    // complex band-pass (band_taps == 0: AM, x = re^2 + im^2 of the user filter's output)
    int band_taps, band_dec, band_hist;
    long long first_b;
    const float *hist_b_re, *hist_b_im;    // band-pass delay line (user-filter outputs of earlier calls)
    float *u_out_re, *u_out_im;            // [nu] user-filter outputs; only the last band_hist are written (the next delay line)
    long long nx;                          // demodulated samples of this call (band-pass outputs, or nu)
    float *x_out;                          // [nx] demodulated samples before the AGC
    const long long *off;                  // [npk + 1] packet boundaries in the demodulated stream
    int npk;
    unsigned *mx_enc;                      // [npk] per-packet maximum as an ordered key, 0 = cleared
    double *sum;                           // [npk] per-packet sum (AM), cleared
};
bool agc_tail_supported(const AgcTailArgs &a);
cudaError_t agc_tail_launch(const AgcTailArgs &a, cudaStream_t st);
bool agc_scan_is_inline(int npk);  // few enough packets for the apply kernel's CTAs to run the AGC recurrence themselves
cudaError_t agc_scan_enc_launch(const long long *off, int npk, const double *sum, const unsigned *mx_enc, const float *state_in,
                                float *state_out, float *gain, float *mean, cudaStream_t st);
struct AgcApplyArgs {
    const float *x;                        // [nx] demodulated samples
    long long nx;
    const long long *off;
    int npk;
    int scan_inline;                       // 1: gains from mx_enc / sum / state_in in the kernel; 0: from gain / mean
    const unsigned *mx_enc;
    const double *sum;
    const float *state_in;                 // [1] lastMax before this call
    float *state_out;                      // [1] after it (scan_inline)
    const float *gain, *mean;              // [npk] (scan kernel's output)
    unsigned *clear_mx;                    // the packet table of the NEXT call: cleared here
    double *clear_sum;
    int clear_n;
    float volume;
    int ratio;                             // demodulated rate / 48 kHz: 1 or 2
    const float *taps_a1;                  // ratio 2: first audio decimator
    int a1_taps, a1_hist;
    const float *hist_a1;
    float *a1_hist_new;                    // ratio 2: the decimator's delay line after this call
    long long first_a1, n1;
    float *audio;
    int nlines;                            // delay lines slid by extra CTAs (user filter, band-pass)
    ChainStateArgs::Line line[2];
};
cudaError_t agc_apply_launch(const AgcApplyArgs &a, bool subtract_mean, cudaStream_t st);

// waterfall / trace preprocessing (render.cu); viewport scalars are computed by the caller (capi.cu)
struct RenderDesc {
    const float *rows = nullptr;
    long long row_stride = 0;
    int ring_rows = 0, n = 0, newest = 0, first_row = 0, nrows = 0;
    const float *peaks = nullptr;
    int width = 0, start = 0;
    float samples_per_px = 0;
    int first_pixel = 0, last_pixel = 0;
    float min_db = 0, scale = 0, db_width = 0, fft_height = 0;
    const uint32_t *colormap = nullptr;
    int colormap_size = 0;
    uint32_t black = 0xFF000000u;
    int avg_len = 0;
    uint32_t *argb = nullptr;
    int *color_index = nullptr;
    float *row_means = nullptr, *peaks_y = nullptr, *time_average = nullptr;
};
cudaError_t render_launch(const RenderDesc &d, cudaStream_t st);

struct SynthComp {
    unsigned int step;
    int amp;
    unsigned int mod_step;
    int mod_k;
};
void synth_make_table(short *tab /*4096*/);
cudaError_t synth_launch(int fmt, unsigned int seed, const SynthComp *comps, int ncomp, int noise_shift,
                         unsigned long long first, long long nsamples, const short *tab_dev, void *out,
                         int num_sms, cudaStream_t st);

}  // namespace rfa
