// capi_dsp.cu -- the IQ->audio half of the C ABI: filter design, FirFilter / ComplexFirFilter /
// RationalResampler objects with the reference's streaming state, the demodulator stages, and
// the whole mix -> resample -> channel filter -> demodulate -> audio decimation chain.
#include <cmath>
#include <cstring>

#include "capi_core.h"
#include "host_design.h"
#include "kernels.h"

using namespace rfa;

namespace {

// device array with a host-side size (grow-only)
struct DevF {
    Buf b;
    float *p() const { return b.as<float>(); }
    int ensure(size_t n) { return b.ensure((n ? n : 1) * sizeof(float)); }
};

int upload(rfa_ctx *c, Buf &dst, const void *src, size_t bytes) {
    if (int rc = dst.ensure(bytes ? bytes : 4)) return rc;
    if (bytes) RFA_CK(cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return RFA_OK;
}

// copy caller data in (host mode) or pass through (device mode)
int stage_in(rfa_ctx *c, Buf &buf, const void *src, size_t bytes, int mem, const void **dev) {
    if (mem == RFA_MEM_DEVICE || !src) {
        *dev = src;
        return RFA_OK;
    }
    if (int rc = buf.ensure(bytes ? bytes : 4)) return rc;
    if (bytes) RFA_CK(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    *dev = buf.p;
    return RFA_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// streaming state shared by FIR and resampler: the last `hist` samples of the stream
// ------------------------------------------------------------------------------------------
struct History {
    int hist = 0;
    Buf re[2], im[2];
    int cur = 0;
    int init(rfa_ctx *c, int n) {
        hist = n;
        for (int i = 0; i < 2; i++) {
            if (int rc = re[i].ensure((size_t)(n ? n : 1) * sizeof(float))) return rc;
            if (int rc = im[i].ensure((size_t)(n ? n : 1) * sizeof(float))) return rc;
        }
        return reset(c);
    }
    int reset(rfa_ctx *c) {
        cur = 0;
        RFA_CK(cudaMemsetAsync(re[0].p, 0, (size_t)(hist ? hist : 1) * sizeof(float), c->stream));
        RFA_CK(cudaMemsetAsync(im[0].p, 0, (size_t)(hist ? hist : 1) * sizeof(float), c->stream));
        return RFA_OK;
    }
    void attach(StreamDesc &d) const {
        d.hist = hist;
        d.hist_re = re[cur].as<float>();
        d.hist_im = im[cur].as<float>();
    }
    // slide the window over the `consumed` samples of stream `d`
    int advance(rfa_ctx *c, const StreamDesc &d, long long consumed) {
        if (hist == 0 || consumed <= 0) return RFA_OK;
        cudaError_t e = history_launch(d, consumed, re[cur ^ 1].as<float>(), im[cur ^ 1].as<float>(), c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "history kernel");
        c->launches++;
        cur ^= 1;
        return RFA_OK;
    }
    void release() {
        for (int i = 0; i < 2; i++) {
            re[i].release();
            im[i].release();
        }
    }
};

struct rfa_fir {
    rfa_ctx *ctx;
    int ntaps, dec;
    bool cplx, exact;
    Buf taps_re, taps_im;
    History h;
    long long first;  // inputs to skip before the next emitting one (decimationCounter, FirFilter.kt:46)
    Buf s_in[2], s_out[2];
    long long initial_first() const { return dec == 1 ? 1 : dec - 1; }
    // how many outputs n inputs produce, without touching state
    long long count(long long n) const { return n > first ? (n - 1 - first) / dec + 1 : 0; }
};

struct rfa_resampler {
    rfa_ctx *ctx;
    int I, D, nt;
    bool exact;
    Buf bank;
    History h;
    long long rel;  // stream offset of the next output's newest sample, relative to the next input
    int ph;         // its polyphase index ("ctr")
    Buf s_in[2], s_out[2];
    long long count(long long n) const {
        if (n <= rel) return 0;
        const long long num = (n - rel) * I - ph;
        return (num + D - 1) / D;
    }
};

// one FIR call on device-resident data; updates counters and history
static int fir_run(rfa_fir *f, StreamDesc in, long long n, bool real_only, float *out_re, float *out_im,
                   long long capacity, long long *n_out, long long *consumed) {
    rfa_ctx *c = f->ctx;
    long long nout = f->count(n);
    long long cons = n;
    if (nout > capacity) {  // FirFilter.kt:80-84: stop at the sample whose output does not fit
        nout = capacity;
        cons = f->first + nout * f->dec;
    }
    f->h.attach(in);
    if (real_only) in.hist_im = nullptr;
    cudaError_t e = fir_launch(in, f->taps_re.as<float>(), f->cplx ? f->taps_im.as<float>() : nullptr, f->ntaps, f->dec,
                               f->first, nout, real_only, out_re, out_im, f->exact, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "fir kernel");
    if (nout > 0) c->launches++;
    // filterReal never touches the imaginary delay line (FirFilter.kt:121-163); here its
    // history is simply carried as zeros, the object is used for one kind of stream only
    if (int rc = f->h.advance(c, in, cons)) return rc;
    f->first = f->first + nout * f->dec - cons;
    if (n_out) *n_out = nout;
    if (consumed) *consumed = cons;
    return RFA_OK;
}

static int resampler_run(rfa_resampler *r, StreamDesc in, long long n, float *out_re, float *out_im,
                         long long capacity, long long *n_out, long long *consumed) {
    rfa_ctx *c = r->ctx;
    long long nout = r->count(n);
    if (nout > capacity) nout = capacity;
    r->h.attach(in);
    const long long T = (long long)r->ph + nout * r->D;
    const long long kk = r->rel + T / r->I;  // where the next output will be computed
    const long long cons = kk < n ? kk : n;  // RationalResampler.kt:136-149
    bool hist_done = false;  // the tiled / stripe kernels slide the delay line themselves (CTA 0, before its tiles)
    cudaError_t e = resample_launch(in, r->bank.as<float>(), r->I, r->D, r->nt, r->rel, r->ph, nout, out_re, out_im,
                                    r->exact, c->stream, c->tune.rs_span, r->h.re[r->h.cur ^ 1].as<float>(),
                                    r->h.im[r->h.cur ^ 1].as<float>(), cons, &hist_done);
    if (e != cudaSuccess) return cuda_fail(e, "resample kernel");
    if (nout > 0) c->launches++;
    if (hist_done) {
        r->h.cur ^= 1;
    } else if (int rc = r->h.advance(c, in, cons)) {
        return rc;
    }
    r->rel = kk - cons;
    r->ph = (int)(T % r->I);
    if (n_out) *n_out = nout;
    if (consumed) *consumed = cons;
    return RFA_OK;
}

extern "C" {

/* ------------------------------------------------------------------ host-side design ---- */
int rfa_tap_window(int kind, double beta, int n, int N, float *out) {
    RFA_REQUIRE(out && kind >= RFA_TAPWIN_BLACKMAN && kind <= RFA_TAPWIN_KAISER, "rfa_tap_window: bad argument");
    RFA_REQUIRE(kind != RFA_TAPWIN_KAISER || beta >= 0.0, "Kaiser beta must be >= 0");
    *out = design::tap_window(kind, beta, n, N);
    return RFA_OK;
}

int rfa_design_lowpass(float gain, float fs, float cutoff, float tw, float att, int window, double beta,
                       int max_taps, float *taps, int capacity, int *ntaps) {
    RFA_REQUIRE(ntaps != nullptr, "ntaps is NULL");
    std::vector<float> t = design::lowpass_taps(gain, fs, cutoff, tw, att, window, beta, max_taps);
    *ntaps = (int)t.size();
    RFA_REQUIRE(!t.empty(), "createLowPassTaps: firdes check failed (fs=%g cutoff=%g tw=%g)", fs, cutoff, tw);
    if (taps) {
        RFA_REQUIRE(capacity >= (int)t.size(), "tap buffer too small: need %zu", t.size());
        memcpy(taps, t.data(), t.size() * sizeof(float));
    }
    return RFA_OK;
}

int rfa_design_bandpass(float gain, float fs, float lo, float hi, float tw, float att, float *taps_re,
                        float *taps_im, int capacity, int *ntaps) {
    RFA_REQUIRE(ntaps != nullptr, "ntaps is NULL");
    std::vector<float> re, im;
    const bool ok = design::bandpass_taps(gain, fs, lo, hi, tw, att, &re, &im);
    *ntaps = ok ? (int)re.size() : 0;
    RFA_REQUIRE(ok, "createBandPass: firdes check failed (fs=%g lo=%g hi=%g tw=%g)", fs, lo, hi, tw);
    if (taps_re && taps_im) {
        RFA_REQUIRE(capacity >= (int)re.size(), "tap buffer too small: need %zu", re.size());
        memcpy(taps_re, re.data(), re.size() * sizeof(float));
        memcpy(taps_im, im.data(), im.size() * sizeof(float));
    }
    return RFA_OK;
}

int rfa_limit_denominator(int num, int den, int max_den, int *out_num, int *out_den) {
    RFA_REQUIRE(num > 0 && den > 0 && max_den > 0 && out_num && out_den, "rfa_limit_denominator: bad argument");
    design::limit_denominator(num, den, max_den, out_num, out_den);
    return RFA_OK;
}

int rfa_design_resampler_taps(int interp, int decim, float fractional_bw, int max_taps, float *taps, int capacity,
                              int *ntaps) {
    RFA_REQUIRE(interp > 0 && decim > 0 && ntaps, "rfa_design_resampler_taps: bad argument");
    if (fractional_bw <= 0 || fractional_bw >= 0.5f) fractional_bw = 0.4f;
    std::vector<float> t = design::resampler_taps(interp, decim, fractional_bw, max_taps);
    *ntaps = (int)t.size();
    if (taps) {
        RFA_REQUIRE(capacity >= (int)t.size(), "tap buffer too small: need %zu", t.size());
        memcpy(taps, t.data(), t.size() * sizeof(float));
    }
    return RFA_OK;
}

/* ------------------------------------------------------------------ FirFilter objects ---- */
int rfa_fir_create(rfa_ctx *c, const float *taps_re, const float *taps_im, int ntaps, int decimation, int flags,
                   rfa_fir **out) {
    RFA_REQUIRE(c && taps_re && out, "rfa_fir_create: NULL argument");
    RFA_REQUIRE(ntaps >= 1 && ntaps <= 4096, "tap count %d outside 1..4096", ntaps);
    RFA_REQUIRE(decimation >= 1 && decimation <= 1024, "decimation %d outside 1..1024", decimation);
    if (int rc = c->use()) return rc;
    rfa_fir *f = new rfa_fir();
    f->ctx = c;
    f->ntaps = ntaps;
    f->dec = decimation;
    f->cplx = taps_im != nullptr;
    f->exact = (flags & RFA_SUM_EXACT) != 0;
    int rc = upload(c, f->taps_re, taps_re, sizeof(float) * ntaps);
    if (!rc && f->cplx) rc = upload(c, f->taps_im, taps_im, sizeof(float) * ntaps);
    if (!rc) rc = f->h.init(c, ntaps - 1);
    if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = RFA_ERR_CUDA;
    if (rc) {
        delete f;
        return rc;
    }
    f->first = f->initial_first();
    *out = f;
    return RFA_OK;
}

int rfa_fir_destroy(rfa_fir *f) {
    if (!f) return RFA_OK;
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    f->taps_re.release();
    f->taps_im.release();
    f->h.release();
    for (int i = 0; i < 2; i++) {
        f->s_in[i].release();
        f->s_out[i].release();
    }
    delete f;
    return RFA_OK;
}

int rfa_fir_reset(rfa_fir *f) {
    RFA_REQUIRE(f != nullptr, "rfa_fir_reset: NULL");
    if (int rc = f->ctx->use()) return rc;
    f->first = f->initial_first();
    return f->h.reset(f->ctx);
}

int rfa_fir_process(rfa_fir *f, const float *in_re, const float *in_im, long long n, float *out_re, float *out_im,
                    long long out_capacity, long long *n_out, long long *consumed, int mem) {
    RFA_REQUIRE(f && n >= 0 && out_capacity >= 0, "rfa_fir_process: bad argument");
    if (n_out) *n_out = 0;
    if (consumed) *consumed = 0;
    if (n == 0) return RFA_OK;
    RFA_REQUIRE(in_re && out_re, "rfa_fir_process: NULL buffer");
    const bool real_only = in_im == nullptr;
    RFA_REQUIRE(!(real_only && f->cplx), "complex taps need a complex input");
    RFA_REQUIRE(real_only || out_im, "complex output needs out_im");
    rfa_ctx *c = f->ctx;
    if (int rc = c->use()) return rc;
    const void *dre, *dim;
    if (int rc = stage_in(c, f->s_in[0], in_re, n * sizeof(float), mem, &dre)) return rc;
    if (int rc = stage_in(c, f->s_in[1], in_im, n * sizeof(float), mem, &dim)) return rc;
    long long want = f->count(n);
    if (want > out_capacity) want = out_capacity;
    float *ore = out_re, *oim = out_im;
    if (mem == RFA_MEM_HOST) {
        if (int rc = f->s_out[0].ensure((want ? want : 1) * sizeof(float))) return rc;
        if (int rc = f->s_out[1].ensure((want ? want : 1) * sizeof(float))) return rc;
        ore = f->s_out[0].as<float>();
        oim = f->s_out[1].as<float>();
    }
    StreamDesc in;
    in.kind = 3;
    in.re = (const float *)dre;
    in.im = (const float *)dim;
    long long nout = 0, cons = 0;
    if (int rc = fir_run(f, in, n, real_only, ore, oim, out_capacity, &nout, &cons)) return rc;
    if (mem == RFA_MEM_HOST) {
        if (nout) {
            RFA_CK(cudaMemcpyAsync(out_re, ore, nout * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            if (!real_only) RFA_CK(cudaMemcpyAsync(out_im, oim, nout * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        }
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    if (n_out) *n_out = nout;
    if (consumed) *consumed = cons;
    return RFA_OK;
}

/* ------------------------------------------------------------------ RationalResampler ---- */
int rfa_resampler_create(rfa_ctx *c, int interp, int decim, const float *taps, int ntaps, float fractional_bw,
                         int max_taps, int flags, rfa_resampler **out) {
    RFA_REQUIRE(c && out, "rfa_resampler_create: NULL argument");
    RFA_REQUIRE(interp > 0 && decim > 0, "Interpolation and decimation must be > 0");  // RationalResampler.kt:54-55
    if (int rc = c->use()) return rc;
    if (fractional_bw <= 0 || fractional_bw >= 0.5f) fractional_bw = 0.4f;
    const int g = design::gcd(interp, decim);
    interp /= g;
    decim /= g;
    std::vector<float> t;
    if (taps)
        t.assign(taps, taps + ntaps);
    else
        t = design::resampler_taps(interp, decim, fractional_bw, max_taps);
    size_t padded = t.size();
    if (padded % interp) padded += interp - padded % interp;
    const int nt = (int)(padded / interp);
    RFA_REQUIRE(nt >= 1 && nt <= 4096, "taps per phase %d outside 1..4096", nt);
    RFA_REQUIRE((long long)interp * nt <= (1 << 26), "polyphase bank too large");
    std::vector<float> bank((size_t)interp * nt, 0.0f);  // firTaps[phase][i] = taps[i*I + phase], RationalResampler.kt:77-81
    for (int p = 0; p < interp; p++)
        for (int i = 0; i < nt; i++) {
            const size_t src = (size_t)i * interp + p;
            bank[(size_t)p * nt + i] = src < t.size() ? t[src] : 0.0f;
        }
    rfa_resampler *r = new rfa_resampler();
    r->ctx = c;
    r->I = interp;
    r->D = decim;
    r->nt = nt;
    r->exact = (flags & RFA_SUM_EXACT) != 0;
    r->rel = 0;
    r->ph = 0;
    int rc = upload(c, r->bank, bank.data(), bank.size() * sizeof(float));
    if (!rc) rc = r->h.init(c, nt - 1);
    if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = RFA_ERR_CUDA;
    if (rc) {
        delete r;
        return rc;
    }
    *out = r;
    return RFA_OK;
}

int rfa_resampler_destroy(rfa_resampler *r) {
    if (!r) return RFA_OK;
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
    r->bank.release();
    r->h.release();
    for (int i = 0; i < 2; i++) {
        r->s_in[i].release();
        r->s_out[i].release();
    }
    delete r;
    return RFA_OK;
}

int rfa_resampler_info(const rfa_resampler *r, int *interp, int *decim, int *taps_per_phase) {
    RFA_REQUIRE(r != nullptr, "rfa_resampler_info: NULL");
    if (interp) *interp = r->I;
    if (decim) *decim = r->D;
    if (taps_per_phase) *taps_per_phase = r->nt;
    return RFA_OK;
}

int rfa_resampler_process(rfa_resampler *r, const float *in_re, const float *in_im, long long n, float *out_re,
                          float *out_im, long long out_capacity, long long *n_out, long long *consumed, int mem) {
    RFA_REQUIRE(r && n >= 0 && out_capacity >= 0, "rfa_resampler_process: bad argument");
    if (n_out) *n_out = 0;
    if (consumed) *consumed = 0;
    if (n == 0) return RFA_OK;
    RFA_REQUIRE(in_re && in_im && out_re && out_im, "rfa_resampler_process: NULL buffer");
    rfa_ctx *c = r->ctx;
    if (int rc = c->use()) return rc;
    const void *dre, *dim;
    if (int rc = stage_in(c, r->s_in[0], in_re, n * sizeof(float), mem, &dre)) return rc;
    if (int rc = stage_in(c, r->s_in[1], in_im, n * sizeof(float), mem, &dim)) return rc;
    long long want = r->count(n);
    if (want > out_capacity) want = out_capacity;
    float *ore = out_re, *oim = out_im;
    if (mem == RFA_MEM_HOST) {
        if (int rc = r->s_out[0].ensure((want ? want : 1) * sizeof(float))) return rc;
        if (int rc = r->s_out[1].ensure((want ? want : 1) * sizeof(float))) return rc;
        ore = r->s_out[0].as<float>();
        oim = r->s_out[1].as<float>();
    }
    StreamDesc in;
    in.kind = 3;
    in.re = (const float *)dre;
    in.im = (const float *)dim;
    long long nout = 0, cons = 0;
    if (int rc = resampler_run(r, in, n, ore, oim, out_capacity, &nout, &cons)) return rc;
    if (mem == RFA_MEM_HOST) {
        if (nout) {
            RFA_CK(cudaMemcpyAsync(out_re, ore, nout * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            RFA_CK(cudaMemcpyAsync(out_im, oim, nout * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        }
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    if (n_out) *n_out = nout;
    if (consumed) *consumed = cons;
    return RFA_OK;
}

}  // extern "C"

/* ---------------------------------------------------------------------- whole chain ---- */
static const int kAudioRate = 48000;  // Demodulator.kt:49
static int quadrature_rate(int mode) {  // Demodulator.kt:53-62
    if (mode == RFA_MODE_WFM) return 8 * kAudioRate;
    if (mode == RFA_MODE_CW) return kAudioRate;
    return 2 * kAudioRate;
}
static const int kMinCw[] = {0, 3000, 3000, 30000, 1500, 1500, 150};  // DemodulationTab.kt:91-97
static const int kMaxCw[] = {50000, 15000, 15000, 150000, 5000, 5000, 800};
static const int kDefCw[] = {0, 8000, 10000, 100000, 2800, 2800, 300};

struct rfa_chain {
    rfa_ctx *ctx;
    rfa_chain_desc d;
    int quad_rate, channel_width;
    bool exact;
    // NCO (generateMixerLookupTable)
    Buf nco;  // cos[512] then sin[512]
    int nco_len = 0, nco_idx = 0, nco_freq = 0;
    rfa_resampler *rs = nullptr;
    rfa_fir *user = nullptr, *band = nullptr, *audio1 = nullptr, *audio2 = nullptr;
    Buf fm_carry, agc_state, agc_scratch, seg;
    Buf agc_tab;                 // fused AGC tail: double sum[2][agc_tab_npk], unsigned max key[2][agc_tab_npk], used in turn
    long long agc_tab_npk = 0, agc_tab_used[2] = {0, 0};
    int agc_tab_cur = 0;
    int agc_cur = 0;             // which of the two agc_state floats holds lastMax
    DevF q_re, q_im, u_re, u_im, b_re, b_im, dem, a1, a2;
    Buf s_iq, s_audio;
    // rfa_chain_process advances the streaming state stage by stage; a failure in the middle of a call leaves
    // the stages out of step, so the chain refuses further packets until rfa_chain_seek re-positions it
    bool poisoned = false;
    int carry_cur = 0;  // which half of fm_carry holds the previous call's last filtered sample
};

extern "C" {

int rfa_mode_info(int mode, int *quad_rate, int *min_width, int *max_width, int *default_width) {
    RFA_REQUIRE(mode >= RFA_MODE_OFF && mode <= RFA_MODE_CW, "unknown demodulation mode %d", mode);
    if (quad_rate) *quad_rate = quadrature_rate(mode);
    if (min_width) *min_width = kMinCw[mode];
    if (max_width) *max_width = kMaxCw[mode];
    if (default_width) *default_width = kDefCw[mode];
    return RFA_OK;
}

int rfa_chain_destroy(rfa_chain *ch) {
    if (!ch) return RFA_OK;
    cudaSetDevice(ch->ctx->device);
    cudaStreamSynchronize(ch->ctx->stream);
    rfa_resampler_destroy(ch->rs);
    rfa_fir_destroy(ch->user);
    rfa_fir_destroy(ch->band);
    rfa_fir_destroy(ch->audio1);
    rfa_fir_destroy(ch->audio2);
    for (Buf *b : {&ch->nco, &ch->fm_carry, &ch->agc_state, &ch->agc_scratch, &ch->seg, &ch->s_iq, &ch->s_audio, &ch->agc_tab}) b->release();
    for (DevF *f : {&ch->q_re, &ch->q_im, &ch->u_re, &ch->u_im, &ch->b_re, &ch->b_im, &ch->dem, &ch->a1, &ch->a2}) f->b.release();
    delete ch;
    return RFA_OK;
}

int rfa_chain_create(rfa_ctx *c, const rfa_chain_desc *d, rfa_chain **out) {
    RFA_REQUIRE(c && d && out, "rfa_chain_create: NULL argument");
    *out = nullptr;
    RFA_REQUIRE(d->format >= RFA_FMT_S8 && d->format <= RFA_FMT_S16LE, "unknown sample format %d", d->format);
    RFA_REQUIRE(d->mode >= RFA_MODE_AM && d->mode <= RFA_MODE_CW, "demodulation mode %d is not a demodulator", d->mode);
    RFA_REQUIRE(d->sample_rate > 0 && d->packet_samples > 0, "sample rate and packet size must be positive");
    if (int rc = c->use()) return rc;
    const int quad = quadrature_rate(d->mode);
    RFA_REQUIRE(d->sample_rate >= quad, "input rate %d below the quadrature rate %d: the reference only downsamples "
                                        "(Resampler.kt:115-116)", d->sample_rate, quad);
    rfa_chain *ch = new rfa_chain();
    ch->ctx = c;
    ch->d = *d;
    ch->quad_rate = quad;
    ch->exact = (d->flags & RFA_SUM_EXACT) != 0;
    int cw = d->channel_width > 0 ? d->channel_width : kDefCw[d->mode];  // Demodulator.kt:75-76
    cw = cw < kMinCw[d->mode] ? kMinCw[d->mode] : (cw > kMaxCw[d->mode] ? kMaxCw[d->mode] : cw);
    ch->channel_width = cw;
    int rc = RFA_OK;
    do {
        // mixer: Scheduler.kt:243 -> mixPacketIntoSamplePacket(packet, buffer, channelFrequency)
        std::vector<float> cs, sn;
        design::nco_tables(d->format, d->sample_rate, (int)(d->source_frequency - d->channel_frequency), &ch->nco_freq, &cs, &sn);
        if (cs.empty() || cs.size() > 500) {
            set_error("mixer table of %zu entries is unusable", cs.size());
            rc = RFA_ERR_INVALID;
            break;
        }
        ch->nco_len = (int)cs.size();
        std::vector<float> both(1024, 0.0f);
        memcpy(both.data(), cs.data(), cs.size() * sizeof(float));
        memcpy(both.data() + 512, sn.data(), sn.size() * sizeof(float));
        if ((rc = upload(c, ch->nco, both.data(), both.size() * sizeof(float)))) break;
        RFA_CK(cudaStreamSynchronize(c->stream));
        // resampler: Resampler.kt:99-102
        int I, D;
        design::limit_denominator(quad, d->sample_rate, 10000, &I, &D);
        if ((rc = rfa_resampler_create(c, I, D, nullptr, 0, 0.4f, 500, d->flags, &ch->rs))) break;
        // user filter: Demodulator.kt:219-226
        {
            std::vector<float> t = design::lowpass_taps(1.0f, (float)quad, (float)cw, quad * 0.10f, 60.0f, 0, 0.0, 0);
            if (t.empty()) {
                set_error("user filter design failed (width %d at %d Hz)", cw, quad);
                rc = RFA_ERR_INVALID;
                break;
            }
            if ((rc = rfa_fir_create(c, t.data(), nullptr, (int)t.size(), 1, d->flags, &ch->user))) break;
        }
        if (d->mode == RFA_MODE_LSB || d->mode == RFA_MODE_USB || d->mode == RFA_MODE_CW) {
            std::vector<float> re, im;
            bool ok;
            int dec;
            if (d->mode == RFA_MODE_CW) {  // Demodulator.kt:372-380
                ok = design::bandpass_taps(1.0f, (float)quad, 750 - cw / 2.0f, 750 + cw / 2.0f, quad * 0.01f, 40.0f, &re, &im);
                dec = 1;
            } else {  // Demodulator.kt:325-333
                const bool upper = d->mode == RFA_MODE_USB;
                ok = design::bandpass_taps(1.0f, (float)quad, upper ? 200.0f : -(float)cw, upper ? (float)cw : -200.0f,
                                           quad * 0.01f, 40.0f, &re, &im);
                dec = 2;
            }
            if (!ok) {
                set_error("band-pass design failed");
                rc = RFA_ERR_INVALID;
                break;
            }
            if ((rc = rfa_fir_create(c, re.data(), im.data(), (int)re.size(), dec, d->flags, &ch->band))) break;
        }
        // audio decimators: AudioSink.java:94-96
        {
            std::vector<float> t1 = design::lowpass_taps(1.0f, 1.0f, 0.1f, 0.15f, 30.0f, 0, 0.0, 0);
            std::vector<float> t2 = design::lowpass_taps(1.0f, 1.0f, 0.1f, 0.1f, 30.0f, 0, 0.0, 0);
            if ((rc = rfa_fir_create(c, t1.data(), nullptr, (int)t1.size(), 2, d->flags, &ch->audio1))) break;
            if ((rc = rfa_fir_create(c, t2.data(), nullptr, (int)t2.size(), 4, d->flags, &ch->audio2))) break;
        }
        if ((rc = ch->fm_carry.ensure(4 * sizeof(float)))) break;
        if ((rc = ch->agc_state.ensure(2 * sizeof(float)))) break;
        RFA_CK(cudaMemsetAsync(ch->fm_carry.p, 0, 4 * sizeof(float), c->stream));
        RFA_CK(cudaMemsetAsync(ch->agc_state.p, 0, 2 * sizeof(float), c->stream));
    } while (0);
    if (rc) {
        rfa_chain_destroy(ch);
        return rc;
    }
    *out = ch;
    return RFA_OK;
}

int rfa_chain_info(const rfa_chain *ch, int *interp, int *decim, int *taps_per_phase, int *quad_rate,
                   int *channel_width, int *nco_length, int *nco_frequency) {
    RFA_REQUIRE(ch != nullptr, "rfa_chain_info: NULL");
    if (interp) *interp = ch->rs->I;
    if (decim) *decim = ch->rs->D;
    if (taps_per_phase) *taps_per_phase = ch->rs->nt;
    if (quad_rate) *quad_rate = ch->quad_rate;
    if (channel_width) *channel_width = ch->channel_width;
    if (nco_length) *nco_length = ch->nco_len;
    if (nco_frequency) *nco_frequency = ch->nco_freq;
    return RFA_OK;
}

long long rfa_chain_max_audio(const rfa_chain *ch, long long nsamples) {
    if (!ch) return 0;
    return (long long)((double)nsamples * kAudioRate / ch->d.sample_rate) + 64;
}

// Position a chain at input sample `sample_index` of a recording: every counter (NCO phase, resampler
// phase, the decimation counters of the filters behind it) takes the value a sequential run from
// sample 0 would have there, every delay line is emptied, the FM carry and the AGC maximum restart.
// A rank of a time-sharded run seeks to a packet boundary one warm-up halo before its segment,
// processes the halo (which refills the delay lines with real samples) and discards its audio.
// *audio_index = number of audio samples the sequential run has produced before sample_index.
int rfa_chain_seek(rfa_chain *ch, long long sample_index, long long *audio_index) {
    RFA_REQUIRE(ch != nullptr, "rfa_chain_seek: NULL");
    RFA_REQUIRE(sample_index >= 0 && sample_index % ch->d.packet_samples == 0,
                "seek position %lld is not a packet boundary (%d samples)", sample_index, ch->d.packet_samples);
    rfa_ctx *c = ch->ctx;
    if (int rc = c->use()) return rc;
    rfa_fir *firs[4] = {ch->user, ch->band, ch->audio1, ch->audio2};
    for (rfa_fir *f : firs)
        if (f) {
            f->first = f->initial_first();
            if (int rc = f->h.reset(c)) return rc;
        }
    ch->rs->rel = 0;
    ch->rs->ph = 0;
    if (int rc = ch->rs->h.reset(c)) return rc;
    RFA_CK(cudaMemsetAsync(ch->fm_carry.p, 0, 4 * sizeof(float), c->stream));
    ch->carry_cur = 0;
    RFA_CK(cudaMemsetAsync(ch->agc_state.p, 0, 2 * sizeof(float), c->stream));
    ch->agc_cur = 0;
    if (ch->agc_tab.p) RFA_CK(cudaMemsetAsync(ch->agc_tab.p, 0, ch->agc_tab.cap, c->stream));  // a failed call may have left maxima behind
    ch->agc_tab_used[0] = ch->agc_tab_used[1] = 0;
    // counters: the same arithmetic rfa_chain_process applies, for sample_index inputs in one step
    const long long n = sample_index;
    ch->nco_idx = (int)(n % ch->nco_len);
    long long produced = 0;
    if (n > 0) {
        rfa_resampler *r = ch->rs;
        const long long nq = r->count(n);
        const long long T = (long long)r->ph + nq * r->D, kk = r->rel + T / r->I;
        r->rel = kk - n;
        r->ph = (int)(T % r->I);
        auto advance = [](rfa_fir *f, long long nin) {
            const long long nout = f->count(nin);
            f->first = f->first + nout * f->dec - nin;
            return nout;
        };
        long long ndem = advance(ch->user, nq);
        int dem_rate = ch->quad_rate;
        if (ch->band) {
            ndem = advance(ch->band, ndem);
            dem_rate = ch->quad_rate / ch->band->dec;
        }
        produced = ndem;
        if (dem_rate > kAudioRate) {
            const int ratio = dem_rate / kAudioRate;
            if (ratio == 8 || ratio == 2) {
                produced = advance(ch->audio1, ndem);
                if (ratio == 8) produced = advance(ch->audio2, produced);
            }
        }
    }
    ch->poisoned = false;
    if (audio_index) *audio_index = produced;
    return RFA_OK;
}

int rfa_chain_process(rfa_chain *ch, const void *iq, long long nsamples, float *audio, long long capacity,
                      long long *n_audio, int mem) {
    RFA_REQUIRE(ch && n_audio, "rfa_chain_process: NULL argument");
    *n_audio = 0;
    RFA_REQUIRE(nsamples >= 0, "negative sample count");
    if (nsamples == 0) return RFA_OK;
    RFA_REQUIRE(iq && audio, "rfa_chain_process: NULL buffer");
    RFA_REQUIRE(capacity >= rfa_chain_max_audio(ch, nsamples), "audio buffer too small: need rfa_chain_max_audio()");
    RFA_REQUIRE(!ch->poisoned, "an earlier rfa_chain_process failed half way: call rfa_chain_seek to re-position the chain");
    rfa_ctx *c = ch->ctx;
    if (int rc = c->use()) return rc;
    struct Guard {  // set until the call has run to its end
        rfa_chain *ch;
        bool ok = false;
        ~Guard() { ch->poisoned = !ok; }
    } guard{ch};
    const int P = ch->d.packet_samples, mode = ch->d.mode;
    const int bps = ch->d.format == RFA_FMT_S16LE ? 4 : 2;
    const long long npk = (nsamples + P - 1) / P;
    RFA_REQUIRE(npk < (1 << 24), "too many packets in one call");
    const void *d_iq;
    if (int rc = stage_in(c, ch->s_iq, iq, (size_t)nsamples * bps, mem, &d_iq)) return rc;
    if (mem == RFA_MEM_DEVICE) RFA_REQUIRE(((uintptr_t)iq % (bps == 4 ? 4 : 2)) == 0, "iq pointer misaligned");

    // ---- packet boundaries in every domain (what the reference's per-packet loops see) -------
    std::vector<long long> q_off(npk + 1, 0), u_off(npk + 1, 0), b_off(npk + 1, 0);
    {
        rfa_resampler r = *ch->rs;  // counters only; buffers are not touched
        rfa_fir u = *ch->user;
        rfa_fir b{};
        if (ch->band) b = *ch->band;
        for (long long p = 0; p < npk; p++) {
            const long long n = (p == npk - 1) ? nsamples - p * P : P;
            const long long nq = r.count(n);
            const long long T = (long long)r.ph + nq * r.D, kk = r.rel + T / r.I;
            r.rel = kk - n;  // downsampling: the packet is always consumed in full
            r.ph = (int)(T % r.I);
            q_off[p + 1] = q_off[p] + nq;
            const long long nu = u.count(nq);
            u.first = u.first + nu * u.dec - nq;
            u_off[p + 1] = u_off[p] + nu;
            if (ch->band) {
                const long long nb = b.count(nu);
                b.first = b.first + nb * b.dec - nu;
                b_off[p + 1] = b_off[p] + nb;
            }
        }
    }
    const long long nq = q_off[npk], nu = u_off[npk], nb = b_off[npk];

    // ---- AM / SSB / CW, RFA_SUM_FMA: the packet boundaries of the demodulated stream go to the device in closed form,
    // by a one-thread-per-packet kernel in FRONT of the resampler (the kernels behind it then follow each other with no
    // copy in between).  The closed form must reproduce the per-packet loop above, boundary for boundary.
    bool packet_table_ready = false;
    if (!ch->exact && mode != RFA_MODE_NFM && mode != RFA_MODE_WFM && ch->user->dec == 1) {
        const bool am = mode == RFA_MODE_AM;
        PacketMap pm{};
        pm.packet_samples = P;
        pm.nsamples = nsamples;
        pm.rel = ch->rs->rel;
        pm.ph = ch->rs->ph;
        pm.I = ch->rs->I;
        pm.D = ch->rs->D;
        pm.first_u = ch->user->first;
        pm.first_b = am ? 0 : ch->band->first;
        pm.dec_b = am ? 0 : ch->band->dec;
        pm.npk = (int)npk;
        const std::vector<long long> &want = am ? u_off : b_off;
        for (long long p = 0; p <= npk; p++)
            RFA_REQUIRE(pm.off(p) == want[(size_t)p], "internal: packet boundary %lld in closed form (%lld) differs from the "
                        "per-packet count (%lld)", p, pm.off(p), want[(size_t)p]);
        if (int rc = ch->seg.ensure((size_t)(npk + 1) * sizeof(long long))) return rc;
        cudaError_t e = packet_table_launch(pm, ch->seg.as<long long>(), c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "packet table kernel");
        c->launches++;
        packet_table_ready = true;
    }

    // ---- K2+K4: convert, mix and resample to the quadrature rate ------------------------------
    if (int rc = ch->q_re.ensure(nq)) return rc;
    if (int rc = ch->q_im.ensure(nq)) return rc;
    {
        StreamDesc in;
        in.kind = ch->d.format;
        in.raw = d_iq;
        in.nco_cos = ch->nco.as<float>();
        in.nco_sin = ch->nco.as<float>() + 512;
        in.nco_len = ch->nco_len;
        in.nco_idx = ch->nco_idx;
        long long got = 0, cons = 0;
        if (int rc = resampler_run(ch->rs, in, nsamples, ch->q_re.p(), ch->q_im.p(), nq, &got, &cons)) return rc;
        ch->nco_idx = (int)((ch->nco_idx + nsamples) % ch->nco_len);
    }
    // ---- FM modes: user filter, discriminator, volume and audio decimators in ONE launch (chain_fused.cu) --------
    if (mode == RFA_MODE_NFM || mode == RFA_MODE_WFM) {
        const int ratio = ch->quad_rate / kAudioRate;  // 8 (wFM) or 2 (nFM): both decimators / the first one only
        rfa_fir *u = ch->user, *f1 = ch->audio1, *f2 = ch->audio2;
        const long long n1 = ratio >= 2 ? f1->count(nu) : 0, n2 = ratio == 8 ? f2->count(n1) : 0;
        const long long nfinal = ratio == 8 ? n2 : (ratio == 2 ? n1 : nu);
        RFA_REQUIRE(nfinal <= capacity, "internal: audio count %lld exceeds capacity %lld", nfinal, capacity);
        if (int rc = ch->dem.ensure(nu)) return rc;
        if (int rc = ch->a1.ensure(n1)) return rc;
        float *d_audio = audio;
        if (mem == RFA_MEM_HOST) {
            if (int rc = ch->a2.ensure(nfinal)) return rc;
            d_audio = ch->a2.p();
        }
        const float max_dev = ch->channel_width * (mode == RFA_MODE_NFM ? 0.75f : 0.85f);  // Demodulator.kt:175-176
        FmTailArgs fa{};
        fa.q_re = ch->q_re.p();
        fa.q_im = ch->q_im.p();
        fa.hist_u_re = u->h.re[u->h.cur].as<float>();
        fa.hist_u_im = u->h.im[u->h.cur].as<float>();
        fa.user_hist = u->h.hist;
        fa.user_taps = u->ntaps;
        fa.taps_user = u->taps_re.as<float>();
        fa.first_u = u->first;
        fa.nu = nu;
        fa.carry_in = ch->fm_carry.as<float>() + 2 * ch->carry_cur;
        fa.carry_out = ch->fm_carry.as<float>() + 2 * (ch->carry_cur ^ 1);
        fa.gain = ch->quad_rate / (float)(2 * 3.14159265358979323846 * (double)max_dev);  // :256
        fa.volume = ch->d.volume;
        fa.ratio = ratio;
        fa.taps_a1 = f1->taps_re.as<float>();
        fa.taps_a2 = f2->taps_re.as<float>();
        fa.a1_taps = f1->ntaps;
        fa.a2_taps = f2->ntaps;
        fa.a1_hist = f1->h.hist;
        fa.a2_hist = f2->h.hist;
        fa.hist_a1 = f1->h.re[f1->h.cur].as<float>();
        fa.hist_a2 = f2->h.re[f2->h.cur].as<float>();
        fa.first_a1 = f1->first;
        fa.first_a2 = f2->first;
        fa.n1 = n1;
        fa.n2 = n2;
        fa.audio = d_audio;
        // the delay lines slide inside the same launch (an extra CTA for the user filter, the owning CTAs for the
        // decimators); nothing else needs the demodulated stream or the first decimator's output in global memory
        const bool fused_state = nu > 0;
        if (fused_state) {
            fa.dem_out = nullptr;
            fa.a1_out = nullptr;
            if (nq > 0) {
                fa.slide_user = 1;
                ChainStateArgs::Line &l = fa.user_line;
                l.in_re = ch->q_re.p();
                l.in_im = ch->q_im.p();
                l.old_re = u->h.re[u->h.cur].as<float>();
                l.old_im = u->h.im[u->h.cur].as<float>();
                l.new_re = u->h.re[u->h.cur ^ 1].as<float>();
                l.new_im = u->h.im[u->h.cur ^ 1].as<float>();
                l.n = nq;
                l.hist = u->h.hist;
            }
            if (ratio >= 2) fa.a1_hist_new = f1->h.re[f1->h.cur ^ 1].as<float>();
            if (ratio == 8 && n1 > 0) fa.a2_hist_new = f2->h.re[f2->h.cur ^ 1].as<float>();
        } else {
            fa.dem_out = ch->dem.p();
            fa.a1_out = ch->a1.p();
        }
        cudaError_t e = fm_tail_launch(fa, ch->exact, c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "fused FM kernel");
        if (!fused_state) {  // a call too short for a single demodulated sample: only delay lines move
            ChainStateArgs sa{};
            auto line = [&](int i, rfa_fir *f, const float *in_re, const float *in_im, long long n) {
                ChainStateArgs::Line &l = sa.line[i];
                l.in_re = in_re;
                l.in_im = in_im;
                l.old_re = f->h.re[f->h.cur].as<float>();
                l.old_im = in_im ? f->h.im[f->h.cur].as<float>() : nullptr;
                l.new_re = f->h.re[f->h.cur ^ 1].as<float>();
                l.new_im = in_im ? f->h.im[f->h.cur ^ 1].as<float>() : nullptr;
                l.n = n;
                l.hist = n > 0 ? f->h.hist : 0;
            };
            line(0, u, ch->q_re.p(), ch->q_im.p(), nq);
            line(1, f1, ch->dem.p(), nullptr, ratio >= 2 ? nu : 0);
            line(2, f2, ch->a1.p(), nullptr, ratio == 8 ? n1 : 0);
            e = chain_state_launch(sa, c->stream);
            if (e != cudaSuccess) return cuda_fail(e, "chain state kernel");
        }
        c->launches += 1;
        if (nq > 0) u->h.cur ^= 1;
        u->first = u->first + nu * u->dec - nq;
        if (nu > 0) ch->carry_cur ^= 1;
        if (ratio >= 2) {
            if (nu > 0) f1->h.cur ^= 1;
            f1->first = f1->first + n1 * f1->dec - nu;
        }
        if (ratio == 8) {
            if (n1 > 0) f2->h.cur ^= 1;
            f2->first = f2->first + n2 * f2->dec - n1;
        }
        if (mem == RFA_MEM_HOST) {
            if (nfinal) RFA_CK(cudaMemcpyAsync(audio, d_audio, nfinal * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            RFA_CK(cudaStreamSynchronize(c->stream));
        }
        // device mode: nothing of this call lives on the host, the call stays asynchronous on the context's stream
        *n_audio = nfinal;
        guard.ok = true;
        return RFA_OK;
    }
    // ---- AM / SSB / CW, RFA_SUM_FMA: user filter, band-pass / power, AGC and audio decimator in three launches
    // (chain_agc.cu) plus the delay-line kernel; nothing of the call lives on the host, so a device-buffer call stays
    // asynchronous on the context's stream like the FM modes.  RFA_SUM_EXACT keeps the separate kernels below.
    if (!ch->exact) {
        const bool am = mode == RFA_MODE_AM;
        rfa_fir *u = ch->user, *b = ch->band, *f1 = ch->audio1;
        const long long nx = am ? nu : nb;
        const int dem_rate = am ? ch->quad_rate : ch->quad_rate / b->dec;  // ComplexFirFilter.java:167
        const int ratio = dem_rate / kAudioRate;
        AgcTailArgs ta{};
        ta.q_re = ch->q_re.p();
        ta.q_im = ch->q_im.p();
        ta.hist_u_re = u->h.re[u->h.cur].as<float>();
        ta.hist_u_im = u->h.im[u->h.cur].as<float>();
        ta.user_hist = u->h.hist;
        ta.user_taps = u->ntaps;
        ta.taps_user = u->taps_re.as<float>();
        ta.first_u = u->first;
        ta.nu = nu;
        ta.nx = nx;
        ta.npk = (int)npk;
        if (!am) {
            ta.taps_b_re = b->taps_re.as<float>();
            ta.taps_b_im = b->taps_im.as<float>();
            ta.band_taps = b->ntaps;
            ta.band_dec = b->dec;
            ta.band_hist = b->h.hist;
            ta.first_b = b->first;
            ta.hist_b_re = b->h.re[b->h.cur].as<float>();
            ta.hist_b_im = b->h.im[b->h.cur].as<float>();
        }
        if (packet_table_ready && (ratio == 1 || ratio == 2) && (ratio == 1 || f1->ntaps <= 9) && agc_tail_supported(ta)) {
            const long long n1 = ratio == 2 ? f1->count(nx) : 0;
            const long long nfinal = ratio == 2 ? n1 : nx;
            RFA_REQUIRE(nfinal <= capacity, "internal: audio count %lld exceeds capacity %lld", nfinal, capacity);
            DevF &xbuf = am ? ch->dem : ch->b_re;
            if (int rc = xbuf.ensure(nx)) return rc;
            if (!am) {
                if (int rc = ch->u_re.ensure(nu)) return rc;
                if (int rc = ch->u_im.ensure(nu)) return rc;
            }
            if (int rc = ch->agc_scratch.ensure(4 * (size_t)npk * sizeof(float))) return rc;
            {  // two per-packet tables (sum, maximum key) used in turn: all zero when (re)allocated, and the apply
               // kernel of a call clears the one the next call fills
                const size_t cap0 = ch->agc_tab.cap;
                if (int rc = ch->agc_tab.ensure(2 * (size_t)npk * (sizeof(double) + sizeof(unsigned)))) return rc;
                if (ch->agc_tab.cap != cap0) {
                    RFA_CK(cudaMemsetAsync(ch->agc_tab.p, 0, ch->agc_tab.cap, c->stream));
                    ch->agc_tab_npk = (long long)(ch->agc_tab.cap / (2 * (sizeof(double) + sizeof(unsigned))));
                    ch->agc_tab_used[0] = ch->agc_tab_used[1] = 0;
                }
            }
            const int tcur = ch->agc_tab_cur;
            double *sum = ch->agc_tab.as<double>() + (size_t)tcur * ch->agc_tab_npk;
            double *sum_next = ch->agc_tab.as<double>() + (size_t)(tcur ^ 1) * ch->agc_tab_npk;
            unsigned *mx_base = reinterpret_cast<unsigned *>(ch->agc_tab.as<double>() + 2 * (size_t)ch->agc_tab_npk);
            unsigned *mx_enc = mx_base + (size_t)tcur * ch->agc_tab_npk, *mx_next = mx_base + (size_t)(tcur ^ 1) * ch->agc_tab_npk;
            float *d_audio = audio;
            if (mem == RFA_MEM_HOST) {
                if (int rc = ch->a2.ensure(nfinal)) return rc;
                d_audio = ch->a2.p();
            }
            ta.u_out_re = ch->u_re.p();
            ta.u_out_im = ch->u_im.p();
            ta.x_out = xbuf.p();
            ta.off = ch->seg.as<long long>();
            ta.mx_enc = mx_enc;
            ta.sum = sum;
            cudaError_t e = agc_tail_launch(ta, c->stream);
            if (e != cudaSuccess) return cuda_fail(e, "fused AGC tail kernel");
            const float *state_in = ch->agc_state.as<float>() + ch->agc_cur;
            float *state_out = ch->agc_state.as<float>() + (ch->agc_cur ^ 1);
            float *gain = ch->agc_scratch.as<float>() + 2 * (size_t)npk, *mean = gain + npk;
            const bool inl = agc_scan_is_inline((int)npk);
            if (!inl) {
                e = agc_scan_enc_launch(ta.off, (int)npk, sum, mx_enc, state_in, state_out, gain, mean, c->stream);
                if (e != cudaSuccess) return cuda_fail(e, "AGC scan kernel");
            }
            AgcApplyArgs aa{};
            aa.x = xbuf.p();
            aa.nx = nx;
            aa.off = ta.off;
            aa.npk = (int)npk;
            aa.scan_inline = inl ? 1 : 0;
            aa.mx_enc = mx_enc;
            aa.sum = sum;
            aa.state_in = state_in;
            aa.state_out = state_out;
            aa.gain = gain;
            aa.mean = mean;
            aa.clear_mx = mx_next;
            aa.clear_sum = sum_next;
            aa.clear_n = (int)ch->agc_tab_used[tcur ^ 1];
            aa.volume = ch->d.volume;
            aa.ratio = ratio;
            aa.taps_a1 = f1->taps_re.as<float>();
            aa.a1_taps = f1->ntaps;
            aa.a1_hist = f1->h.hist;
            aa.hist_a1 = f1->h.re[f1->h.cur].as<float>();
            aa.a1_hist_new = f1->h.re[f1->h.cur ^ 1].as<float>();
            aa.first_a1 = f1->first;
            aa.n1 = n1;
            aa.audio = d_audio;
            auto line = [&](rfa_fir *f, const float *in_re, const float *in_im, long long n) {
                ChainStateArgs::Line &l = aa.line[aa.nlines++];
                l.in_re = in_re;
                l.in_im = in_im;
                l.old_re = f->h.re[f->h.cur].as<float>();
                l.old_im = f->h.im[f->h.cur].as<float>();
                l.new_re = f->h.re[f->h.cur ^ 1].as<float>();
                l.new_im = f->h.im[f->h.cur ^ 1].as<float>();
                l.n = n;
                l.hist = f->h.hist;
            };
            if (nq > 0) line(u, ch->q_re.p(), ch->q_im.p(), nq);
            if (!am && nu > 0) line(b, ch->u_re.p(), ch->u_im.p(), nu);
            e = agc_apply_launch(aa, am, c->stream);
            if (e != cudaSuccess) return cuda_fail(e, "AGC apply kernel");
            c->launches += inl ? 2 : 3;
            ch->agc_tab_used[tcur] = npk;
            ch->agc_tab_cur ^= 1;
            ch->agc_cur ^= 1;
            if (nq > 0) u->h.cur ^= 1;
            u->first = u->first + nu * u->dec - nq;
            if (!am) {
                if (nu > 0) b->h.cur ^= 1;
                b->first = b->first + nb * b->dec - nu;
            }
            if (ratio == 2) {
                f1->h.cur ^= 1;  // nx > 0 here
                f1->first = f1->first + n1 * f1->dec - nx;
            }
            if (mem == RFA_MEM_HOST) {
                if (nfinal) RFA_CK(cudaMemcpyAsync(audio, d_audio, nfinal * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
                RFA_CK(cudaStreamSynchronize(c->stream));
            }
            *n_audio = nfinal;
            guard.ok = true;
            return RFA_OK;
        }
    }
    // ---- K5: user (channel) filter ---------------------------------------------------------------
    if (int rc = ch->u_re.ensure(nu)) return rc;
    if (int rc = ch->u_im.ensure(nu)) return rc;
    {
        StreamDesc in;
        in.re = ch->q_re.p();
        in.im = ch->q_im.p();
        if (int rc = fir_run(ch->user, in, nq, false, ch->u_re.p(), ch->u_im.p(), nu, nullptr, nullptr)) return rc;
    }
    // ---- K6: demodulate ----------------------------------------------------------------------------
    const float volume = ch->d.volume;
    const float *dem = nullptr;  // real audio-band signal after demodulation
    long long ndem = 0;
    int dem_rate = ch->quad_rate;
    auto upload_segments = [&](const std::vector<long long> &off) -> int {
        return upload(c, ch->seg, off.data(), off.size() * sizeof(long long));
    };
    auto max_segment = [&](const std::vector<long long> &off) {
        long long m = 0;
        for (long long p = 0; p < npk; p++) m = off[p + 1] - off[p] > m ? off[p + 1] - off[p] : m;
        return m;
    };
    if (mode == RFA_MODE_AM) {
        if (int rc = ch->dem.ensure(nu)) return rc;
        if (int rc = ch->agc_scratch.ensure(4 * (size_t)npk * sizeof(float))) return rc;
        if (int rc = upload_segments(u_off)) return rc;
        cudaError_t e = demod_power_launch(ch->u_re.p(), ch->u_im.p(), nu, ch->dem.p(), c->num_sms, c->stream);
        if (e == cudaSuccess)
            e = agc_launch(ch->dem.p(), ch->seg.as<long long>(), (int)npk, max_segment(u_off), true,
                           ch->agc_state.as<float>() + ch->agc_cur, ch->agc_scratch.as<float>(), volume, ch->exact, c->num_sms, c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "am kernels");
        c->launches += 4;
        dem = ch->dem.p();
        ndem = nu;
    } else {  // LSB / USB / CW: complex band-pass, real part, AGC
        if (int rc = ch->b_re.ensure(nb)) return rc;
        if (int rc = ch->b_im.ensure(nb)) return rc;
        if (int rc = ch->agc_scratch.ensure(4 * (size_t)npk * sizeof(float))) return rc;
        StreamDesc in;
        in.re = ch->u_re.p();
        in.im = ch->u_im.p();
        if (int rc = fir_run(ch->band, in, nu, false, ch->b_re.p(), ch->b_im.p(), nb, nullptr, nullptr)) return rc;
        if (int rc = upload_segments(b_off)) return rc;
        cudaError_t e = agc_launch(ch->b_re.p(), ch->seg.as<long long>(), (int)npk, max_segment(b_off), false,
                                   ch->agc_state.as<float>() + ch->agc_cur, ch->agc_scratch.as<float>(), volume, ch->exact, c->num_sms,
                                   c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "agc kernels");
        c->launches += 3;
        dem = ch->b_re.p();
        ndem = nb;
        dem_rate = ch->quad_rate / ch->band->dec;  // ComplexFirFilter.java:167
    }
    // ---- K7: audio decimation (AudioSink.java:182-187, :215-237) -------------------------------------
    const float *final_ptr = dem;
    long long nfinal = ndem;
    if (dem_rate > kAudioRate) {
        const int ratio = dem_rate / kAudioRate;
        if (ratio == 8 || ratio == 2) {
            const long long n1 = ch->audio1->count(ndem);
            if (int rc = ch->a1.ensure(n1)) return rc;
            StreamDesc in;
            in.re = dem;
            if (int rc = fir_run(ch->audio1, in, ndem, true, ch->a1.p(), nullptr, n1, nullptr, nullptr)) return rc;
            final_ptr = ch->a1.p();
            nfinal = n1;
            if (ratio == 8) {
                const long long n2 = ch->audio2->count(n1);
                if (int rc = ch->a2.ensure(n2)) return rc;
                StreamDesc in2;
                in2.re = ch->a1.p();
                if (int rc = fir_run(ch->audio2, in2, n1, true, ch->a2.p(), nullptr, n2, nullptr, nullptr)) return rc;
                final_ptr = ch->a2.p();
                nfinal = n2;
            }
        }  // other ratios: AudioSink logs "not supported" and plays the unfiltered packet
    }
    RFA_REQUIRE(nfinal <= capacity, "internal: audio count %lld exceeds capacity %lld", nfinal, capacity);
    if (nfinal)
        RFA_CK(cudaMemcpyAsync(audio, final_ptr, nfinal * sizeof(float),
                               mem == RFA_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
    // the segment table was staged from a host vector that dies with this call
    RFA_CK(cudaStreamSynchronize(c->stream));
    *n_audio = nfinal;
    guard.ok = true;
    return RFA_OK;
}

/* ------------------------------------------------------------- single demodulator stages ---- */
int rfa_demod_fm(rfa_ctx *c, const float *re, const float *im, long long n, float *carry, float quadrature_gain,
                 float volume, float *out, int flags, int mem) {
    RFA_REQUIRE(c && carry && n >= 0, "rfa_demod_fm: bad argument");
    if (n == 0) return RFA_OK;
    RFA_REQUIRE(re && im && out, "rfa_demod_fm: NULL buffer");
    if (int rc = c->use()) return rc;
    const void *dre, *dim;
    if (int rc = stage_in(c, c->stage[0], re, n * sizeof(float), mem, &dre)) return rc;
    if (int rc = stage_in(c, c->stage[1], im, n * sizeof(float), mem, &dim)) return rc;
    if (int rc = c->stage[5].ensure(2 * sizeof(float))) return rc;
    RFA_CK(cudaMemcpyAsync(c->stage[5].p, carry, 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    float *dout = out;
    if (mem == RFA_MEM_HOST) {
        if (int rc = c->stage[2].ensure(n * sizeof(float))) return rc;
        dout = c->stage[2].as<float>();
    }
    cudaError_t e = demod_fm_launch((const float *)dre, (const float *)dim, n, c->stage[5].as<float>(), quadrature_gain,
                                    volume, dout, (flags & RFA_SUM_EXACT) != 0, c->num_sms, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "fm kernel");
    c->launches += 2;
    RFA_CK(cudaMemcpyAsync(carry, c->stage[5].p, 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (mem == RFA_MEM_HOST) RFA_CK(cudaMemcpyAsync(out, dout, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    RFA_CK(cudaStreamSynchronize(c->stream));
    return RFA_OK;
}

// one packet of AM (re/im in) or of the SSB/CW gain control (x in place, re == NULL)
static int agc_packet(rfa_ctx *c, const float *re, const float *im, float *x, long long n, float *last_max,
                      float volume, bool am, int flags, int mem) {
    if (int rc = c->use()) return rc;
    float *dx = x;
    if (mem == RFA_MEM_HOST) {
        if (int rc = c->stage[2].ensure(n * sizeof(float))) return rc;
        dx = c->stage[2].as<float>();
        if (!am) RFA_CK(cudaMemcpyAsync(dx, x, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    if (am) {
        const void *dre, *dim;
        if (int rc = stage_in(c, c->stage[0], re, n * sizeof(float), mem, &dre)) return rc;
        if (int rc = stage_in(c, c->stage[1], im, n * sizeof(float), mem, &dim)) return rc;
        cudaError_t e = demod_power_launch((const float *)dre, (const float *)dim, n, dx, c->num_sms, c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "power kernel");
        c->launches++;
    }
    if (int rc = c->stage[5].ensure(8 * sizeof(float) + 2 * sizeof(long long))) return rc;
    long long off[2] = {0, n};
    char *base = (char *)c->stage[5].p;
    RFA_CK(cudaMemcpyAsync(base, off, sizeof(off), cudaMemcpyHostToDevice, c->stream));
    float *state = (float *)(base + sizeof(off));
    RFA_CK(cudaMemcpyAsync(state, last_max, sizeof(float), cudaMemcpyHostToDevice, c->stream));
    cudaError_t e = agc_launch(dx, (const long long *)base, 1, n, am, state, state + 1, volume,
                               (flags & RFA_SUM_EXACT) != 0, c->num_sms, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "agc kernels");
    c->launches += 3;
    RFA_CK(cudaMemcpyAsync(last_max, state, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (mem == RFA_MEM_HOST) RFA_CK(cudaMemcpyAsync(x, dx, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    RFA_CK(cudaStreamSynchronize(c->stream));
    return RFA_OK;
}

int rfa_demod_am(rfa_ctx *c, const float *re, const float *im, long long n, float *last_max, float volume,
                 float *out, int flags, int mem) {
    RFA_REQUIRE(c && last_max && n >= 0, "rfa_demod_am: bad argument");
    if (n == 0) return RFA_OK;
    RFA_REQUIRE(re && im && out, "rfa_demod_am: NULL buffer");
    return agc_packet(c, re, im, out, n, last_max, volume, true, flags, mem);
}

int rfa_agc(rfa_ctx *c, float *x, long long n, float *last_max, float volume, int flags, int mem) {
    RFA_REQUIRE(c && last_max && n >= 0, "rfa_agc: bad argument");
    if (n == 0) return RFA_OK;
    RFA_REQUIRE(x != nullptr, "rfa_agc: NULL buffer");
    return agc_packet(c, nullptr, nullptr, x, n, last_max, volume, false, flags, mem);
}

}  // extern "C"
