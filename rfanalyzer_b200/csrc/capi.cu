// capi.cu -- context, conversion, NativeDsp-level and fused-spectrum entry points of the
// C ABI (include/rfa_b200.h).  Thin: argument checks, table caches, staging for host
// buffers, then a kernel launch on the context's stream.
#include <cstring>

#include "capi_core.h"
#include "host_design.h"
#include "kernels.h"
#include "rfa_tables.h"
#include "spectrum_launch.h"

namespace rfa {

static thread_local std::string g_error;

void set_error(const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return RFA_ERR_CUDA;
}

int Buf::ensure(size_t bytes) {
    if (bytes <= cap) return RFA_OK;
    release();
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = pinned ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        cap = 0;
        set_error("out of %s memory allocating %zu bytes (%s)", pinned ? "pinned host" : "device", want,
                  cudaGetErrorString(e));
        cudaGetLastError();
        return RFA_ERR_NOMEM;
    }
    cap = want;
    return RFA_OK;
}

void Buf::release() {
    if (p) {
        if (pinned)
            cudaFreeHost(p);
        else
            cudaFree(p);
    }
    p = nullptr;
    cap = 0;
}

static int fmt_bytes(int fmt) { return fmt == RFA_FMT_S16LE ? 4 : 2; }
static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

}  // namespace rfa

using namespace rfa;

int rfa_ctx::use() {
    RFA_CK(cudaSetDevice(device));
    return RFA_OK;
}

static const int kColumnTwiddles = 1 << 24;

int rfa_ctx::get_twiddles(int n, const cf **out) {
    auto it = twiddles.find(n);
    if (it == twiddles.end()) {
        // n > 0: per-pass Stockham tables of an n-point transform; n < 0: plain exp(-2*pi*i*t/|n|);
        // n = kColumnTwiddles + N: the four-step column twiddles [N/256][256], W_N^(n2 k1) at row k1 (cluster path)
        std::vector<cf> host;
        if (n > kColumnTwiddles) {
            const int N = n - kColumnTwiddles, n1 = N / 256;
            std::vector<cf> w((size_t)N);
            make_twiddles(N, w.data());
            host.resize((size_t)N);
            for (int k1 = 0; k1 < n1; k1++)
                for (int n2 = 0; n2 < 256; n2++) host[(size_t)k1 * 256 + n2] = w[(size_t)(n2 * k1) & (size_t)(N - 1)];
        } else if (n > 0) {
            host = make_pass_twiddles(n);
        } else {
            host.resize((size_t)-n);
            make_twiddles(-n, host.data());
        }
        cf *dev = nullptr;
        RFA_CK(cudaMalloc(&dev, sizeof(cf) * host.size()));
        RFA_CK(cudaMemcpyAsync(dev, host.data(), sizeof(cf) * host.size(), cudaMemcpyHostToDevice, stream));
        RFA_CK(cudaStreamSynchronize(stream));
        it = twiddles.emplace(n, dev).first;
    }
    *out = it->second;
    return RFA_OK;
}

int rfa_ctx::get_window(int kind, int n, const float **out) {
    auto key = std::make_pair(kind, n);
    auto it = windows.find(key);
    if (it == windows.end()) {
        std::vector<float> host((size_t)n);
        make_window(kind, n, host.data());
        float *dev = nullptr;
        RFA_CK(cudaMalloc(&dev, sizeof(float) * (size_t)n));
        RFA_CK(cudaMemcpyAsync(dev, host.data(), sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, stream));
        RFA_CK(cudaStreamSynchronize(stream));
        it = windows.emplace(key, dev).first;
    }
    *out = it->second;
    return RFA_OK;
}

extern "C" {

int rfa_version(void) { return RFA_VERSION; }
const char *rfa_last_error(void) { return g_error.c_str(); }

int rfa_ctx_create(int device, void *stream, rfa_ctx **out) {
    RFA_REQUIRE(out != nullptr, "rfa_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); librfa_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return RFA_ERR_CUDA;
    }
    RFA_REQUIRE(device >= 0 && device < count, "rfa_ctx_create: device %d out of range (0..%d)", device, count - 1);
    RFA_CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    RFA_CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
        return RFA_ERR_UNSUPPORTED;
    }
    rfa_ctx *c = new rfa_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    // any failure below releases what was created so far (rfa_ctx_destroy copes with half-built contexts)
    auto fail = [&](cudaError_t e, const char *what) {
        rfa_ctx_destroy(c);
        return cuda_fail(e, what);
    };
    cudaError_t e2;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        if ((e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e2, "cudaStreamCreate");
        c->own_stream = true;
    }
    if ((e2 = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking)) != cudaSuccess) return fail(e2, "cudaStreamCreate");
    if ((e2 = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking)) != cudaSuccess) return fail(e2, "cudaStreamCreate");
    for (int i = 0; i < 2; i++) {
        if ((e2 = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e2, "cudaEventCreate");
        if ((e2 = cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e2, "cudaEventCreate");
        if ((e2 = cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e2, "cudaEventCreate");
    }
    *out = c;
    return RFA_OK;
}

int rfa_ctx_destroy(rfa_ctx *c) {
    if (!c) return RFA_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto &kv : c->twiddles) cudaFree(kv.second);
    for (auto &kv : c->windows) cudaFree(kv.second);
    if (c->synth_table) cudaFree(c->synth_table);
    for (auto &b : c->stage) b.release();
    for (int i = 0; i < 2; i++) {
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
    }
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return RFA_OK;
}

int rfa_ctx_sync(rfa_ctx *c) {
    RFA_REQUIRE(c != nullptr, "rfa_ctx_sync: ctx is NULL");
    RFA_CK(cudaStreamSynchronize(c->stream));
    return RFA_OK;
}
int rfa_ctx_device(const rfa_ctx *c) { return c ? c->device : -1; }
int rfa_ctx_sm_count(const rfa_ctx *c) { return c ? c->num_sms : 0; }
void *rfa_ctx_stream(const rfa_ctx *c) { return c ? (void *)c->stream : nullptr; }
long long rfa_ctx_launch_count(const rfa_ctx *c) { return c ? c->launches : 0; }

// the context's knobs (tuning.h): name -> field
static long long *option_slot(rfa::Tuning &t, const char *name, long long *scratch, int **as_int) {
    *as_int = nullptr;
    struct IntOpt { const char *name; int rfa::Tuning::*field; };
    static const IntOpt ints[] = {{"staged", &rfa::Tuning::staged},     {"pdl", &rfa::Tuning::pdl},
                                  {"max_grid", &rfa::Tuning::max_grid}, {"fs_tma", &rfa::Tuning::fs_tma},
                                  {"fs_ztma", &rfa::Tuning::fs_ztma},   {"fs_pdl", &rfa::Tuning::fs_pdl},
                                  {"chunk_kib", &rfa::Tuning::chunk_kib}, {"rs_span", &rfa::Tuning::rs_span},
                                  {"cluster", &rfa::Tuning::cluster},
#ifdef RFA_LAB
                                  {"kernel", &rfa::Tuning::kernel},     {"fourstep", &rfa::Tuning::fourstep},
                                  {"fs_fused", &rfa::Tuning::fs_fused},
#endif
    };
    for (const IntOpt &o : ints)
        if (!strcmp(name, o.name)) {
            *as_int = &(t.*(o.field));
            return scratch;
        }
    if (!strcmp(name, "fs_batch_kib")) return &t.fs_batch_kib;
#ifdef RFA_LAB
    if (!strcmp(name, "fs_ring_kib")) return &t.fs_ring_kib;
#endif
    return nullptr;
}

int rfa_ctx_set_option(rfa_ctx *c, const char *name, long long value) {
    RFA_REQUIRE(c && name, "rfa_ctx_set_option: NULL argument");
    long long scratch = 0;
    int *as_int = nullptr;
    long long *slot = option_slot(c->tune, name, &scratch, &as_int);
    if (!slot) {
        set_error("unknown option '%s' (lab-only options need librfa_b200_lab.so)", name);
        return RFA_ERR_UNSUPPORTED;
    }
    RFA_REQUIRE(value >= 0 && value <= 0x7FFFFFFFLL, "option '%s': value %lld out of range", name, value);
    RFA_REQUIRE(strcmp(name, "fs_batch_kib") || value >= 1, "fs_batch_kib must be at least 1");
    RFA_REQUIRE(strcmp(name, "chunk_kib") || value >= 1, "chunk_kib must be at least 1");
    if (as_int)
        *as_int = (int)value;
    else
        *slot = value;
    return RFA_OK;
}

int rfa_ctx_get_option(rfa_ctx *c, const char *name, long long *value) {
    RFA_REQUIRE(c && name && value, "rfa_ctx_get_option: NULL argument");
    long long scratch = 0;
    int *as_int = nullptr;
    long long *slot = option_slot(c->tune, name, &scratch, &as_int);
    if (!slot) {
        set_error("unknown option '%s' (lab-only options need librfa_b200_lab.so)", name);
        return RFA_ERR_UNSUPPORTED;
    }
    *value = as_int ? *as_int : *slot;
    return RFA_OK;
}

int rfa_host_alloc(size_t bytes, void **out) {
    RFA_REQUIRE(out != nullptr, "rfa_host_alloc: out is NULL");
    RFA_CK(cudaMallocHost(out, bytes ? bytes : 1));
    return RFA_OK;
}
int rfa_host_free(void *p) {
    if (p) RFA_CK(cudaFreeHost(p));
    return RFA_OK;
}

/* ---------------------------------------------------------------- convert / mix ---- */
static int convert_or_mix(rfa_ctx *c, int fmt, const void *iq, long long n, const float *cosT, const float *sinT,
                          int ncoLen, int ncoIdx, float *re, float *im, int mem, bool mix) {
    RFA_REQUIRE(c != nullptr, "ctx is NULL");
    RFA_REQUIRE(fmt >= RFA_FMT_S8 && fmt <= RFA_FMT_S16LE, "unknown sample format %d", fmt);
    RFA_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return RFA_OK;
    RFA_REQUIRE(iq && re && im, "NULL buffer");
    if (int rc = c->use()) return rc;
    const size_t in_bytes = (size_t)n * fmt_bytes(fmt), out_bytes = (size_t)n * sizeof(float);
    const float *dc = nullptr, *ds = nullptr;
    if (mix) {
        RFA_REQUIRE(cosT && sinT, "NULL NCO table");
        RFA_REQUIRE(ncoLen >= 1 && ncoLen <= 500, "NCO table length %d outside 1..500", ncoLen);
        RFA_REQUIRE(ncoIdx >= 0 && ncoIdx < ncoLen, "NCO index %d outside table", ncoIdx);
        if (int rc = c->stage[3].ensure(sizeof(float) * 1024)) return rc;
        float *t = c->stage[3].as<float>();
        RFA_CK(cudaMemcpyAsync(t, cosT, sizeof(float) * ncoLen, cudaMemcpyHostToDevice, c->stream));
        RFA_CK(cudaMemcpyAsync(t + 512, sinT, sizeof(float) * ncoLen, cudaMemcpyHostToDevice, c->stream));
        dc = t;
        ds = t + 512;
    }
    const void *din = iq;
    float *dre = re, *dim = im;
    if (mem == RFA_MEM_HOST) {
        if (int rc = c->stage[0].ensure(in_bytes)) return rc;
        if (int rc = c->stage[1].ensure(out_bytes)) return rc;
        if (int rc = c->stage[2].ensure(out_bytes)) return rc;
        RFA_CK(cudaMemcpyAsync(c->stage[0].p, iq, in_bytes, cudaMemcpyHostToDevice, c->stream));
        din = c->stage[0].p;
        dre = c->stage[1].as<float>();
        dim = c->stage[2].as<float>();
    } else {
        RFA_REQUIRE(((uintptr_t)iq % (fmt == RFA_FMT_S16LE ? 4 : 2)) == 0, "iq pointer misaligned for its format");
    }
    cudaError_t e = mix ? mix_launch(fmt, din, n, dre, dim, dc, ds, ncoLen, ncoIdx, c->num_sms, c->stream)
                        : convert_launch(fmt, din, n, dre, dim, c->num_sms, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, mix ? "mix kernel" : "convert kernel");
    c->launches++;
    if (mem == RFA_MEM_HOST) {
        RFA_CK(cudaMemcpyAsync(re, dre, out_bytes, cudaMemcpyDeviceToHost, c->stream));
        RFA_CK(cudaMemcpyAsync(im, dim, out_bytes, cudaMemcpyDeviceToHost, c->stream));
        RFA_CK(cudaStreamSynchronize(c->stream));
    } else if (mix) {
        // the NCO tables were staged from host memory the caller may reuse right away
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    return RFA_OK;
}

int rfa_convert(rfa_ctx *c, int fmt, const void *iq, long long n, float *re, float *im, int mem) {
    return convert_or_mix(c, fmt, iq, n, nullptr, nullptr, 0, 0, re, im, mem, false);
}

int rfa_nco_design(int fmt, int sample_rate, int mix_frequency, int *effective_frequency, int *length,
                   float *cos_table, float *sin_table) {
    RFA_REQUIRE(fmt >= RFA_FMT_S8 && fmt <= RFA_FMT_S16LE, "unknown sample format %d", fmt);
    RFA_REQUIRE(sample_rate > 0, "sample rate must be positive");
    RFA_REQUIRE(length && cos_table && sin_table, "NULL output");
    std::vector<float> c, s;
    int eff = 0;
    design::nco_tables(fmt, sample_rate, mix_frequency, &eff, &c, &s);
    RFA_REQUIRE(c.size() <= 500, "NCO table of %zu entries exceeds the reference's 500", c.size());
    if (effective_frequency) *effective_frequency = eff;
    *length = (int)c.size();
    memcpy(cos_table, c.data(), sizeof(float) * c.size());
    memcpy(sin_table, s.data(), sizeof(float) * s.size());
    return RFA_OK;
}

int rfa_mix(rfa_ctx *c, int fmt, const void *iq, long long n, const float *cosT, const float *sinT, int ncoLen,
            int ncoIdx, float *re, float *im, int mem) {
    return convert_or_mix(c, fmt, iq, n, cosT, sinT, ncoLen, ncoIdx, re, im, mem, true);
}

/* ---------------------------------------------------------------- NativeDsp level ---- */
int rfa_make_window(int window, int n, float *out) {
    RFA_REQUIRE(window >= RFA_WIN_BLACKMAN_REF && window <= RFA_WIN_RECT, "unknown window %d", window);
    RFA_REQUIRE(n > 0 && out, "bad window request");
    make_window(window, n, out);
    return RFA_OK;
}

// in_kind: FMT_CF32 (interleaved) or FMT_PF32 (planar); out_kind: OUT_DB / OUT_CPLX
static int fft_generic(rfa_ctx *c, int in_kind, int out_kind, const float *in_a, const float *in_b, float *out,
                       int n, long long batch, int window, int mem) {
    RFA_REQUIRE(c != nullptr, "ctx is NULL");
    RFA_REQUIRE(is_pow2(n) && n >= 16 && n <= 65536,
                "FFT size %d unsupported: need a power of two in 16..65536", n);
    RFA_REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return RFA_OK;
    RFA_REQUIRE(in_a && out && (in_kind != FMT_PF32 || in_b), "NULL buffer");
    if (int rc = c->use()) return rc;
    const size_t total = (size_t)batch * (size_t)n;
    const size_t in_floats = in_kind == FMT_CF32 ? 2 * total : total;
    const size_t out_floats = out_kind == OUT_CPLX ? 2 * total : total;
    SpectrumLaunch L{};
    L.N = n;
    L.in_fmt = in_kind;
    L.out_kind = out_kind;
    L.stream = c->stream;
    L.num_sms = c->num_sms;
    L.tune = c->tune;
    const int nl = n > 16384 ? 16384 : n;
    if (int rc = c->get_twiddles(nl, &L.p.tw)) return rc;
    if (n > 16384)
        if (int rc = c->get_twiddles(-n, &L.p.twN)) return rc;
    if (window >= 0)
        if (int rc = c->get_window(window, n, &L.p.win)) return rc;
    const float *da = in_a, *db = in_b;
    float *dout = out;
    if (mem == RFA_MEM_HOST) {
        if (int rc = c->stage[0].ensure(in_floats * sizeof(float))) return rc;
        if (int rc = c->stage[1].ensure(out_floats * sizeof(float))) return rc;
        RFA_CK(cudaMemcpyAsync(c->stage[0].p, in_a, in_floats * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        da = c->stage[0].as<float>();
        if (in_kind == FMT_PF32) {
            if (int rc = c->stage[2].ensure(total * sizeof(float))) return rc;
            RFA_CK(cudaMemcpyAsync(c->stage[2].p, in_b, total * sizeof(float), cudaMemcpyHostToDevice, c->stream));
            db = c->stage[2].as<float>();
        }
        dout = c->stage[1].as<float>();
    }
    L.p.in = da;
    L.p.in_im = db;
    L.p.rows = dout;
    L.p.row0 = 0;
    L.p.row_step = 1;
    L.p.ring_rows = 0;
    L.p.row_stride = out_kind == OUT_CPLX ? 2LL * n : n;
    L.p.nframes = batch;
    L.p.store_from = 0;
    L.p.inv_n2 = -3.0102999566398120f * log2f((float)n);  // dB bias, see logmag_db
    // chunk counters of the kernel's work distribution (stage[7] is reserved for them)
    if (!c->stage[7].p) {
        if (int rc = c->stage[7].ensure(TICKET_WORDS * sizeof(unsigned int))) return rc;
        RFA_CK(cudaMemsetAsync(c->stage[7].p, 0, TICKET_WORDS * sizeof(unsigned int), c->stream));
    }
    L.p.ticket = c->stage[7].as<unsigned int>();
    cudaError_t e = spectrum_launch(L);
    if (e != cudaSuccess) return cuda_fail(e, "spectrum kernel");
    c->launches++;
    if (mem == RFA_MEM_HOST) {
        RFA_CK(cudaMemcpyAsync(out, dout, out_floats * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    return RFA_OK;
}

int rfa_fft_c2c(rfa_ctx *c, const float *in, float *out, int n, long long batch, int mem) {
    return fft_generic(c, FMT_CF32, OUT_CPLX, in, nullptr, out, n, batch, -1, mem);
}
int rfa_fft_logmag(rfa_ctx *c, const float *in, float *mag, int n, long long batch, int mem) {
    return fft_generic(c, FMT_CF32, OUT_DB, in, nullptr, mag, n, batch, -1, mem);
}
int rfa_windowed_fft_logmag(rfa_ctx *c, const float *re, const float *im, float *mag, int n, long long batch,
                            int window, int mem) {
    RFA_REQUIRE(window >= RFA_WIN_BLACKMAN_REF && window <= RFA_WIN_RECT, "unknown window %d", window);
    return fft_generic(c, FMT_PF32, OUT_DB, re, im, mag, n, batch, window, mem);
}

/* ---------------------------------------------------------------- fused spectrum ---- */
}  // extern "C"

struct rfa_spectrum_plan {
    rfa_ctx *ctx;
    rfa_spectrum_desc d;
    long long ema_frames = 0;  // RFA_AVG_EMA: newest frames whose weight (1-alpha)^k is above 2^-40
    const cf *tw = nullptr, *twN = nullptr;
    const float *win = nullptr;
    Buf ticket;     // finished-tail-row counters of the fused kernel
    Buf zbuf;       // four-step intermediate (N >= 32768)
    Buf fsync;      // four-step fused launch: per-frame hand-over counters
    Buf tail;       // (L+1) rows when the caller stores no rows
    Buf dpeaks;     // device running peaks (host mode)
    Buf davg;       // device average (host mode)
    Buf din[2], drows[2];  // pipelined host mode
};

// avg request of one launch (device pointer + which rows hold data)
struct AvgReq {
    float *avg = nullptr;
    long long valid = 0;
    bool accumulate = false;  // RFA_AVG_EMA: continue from the values in avg
};

// one device-resident batch = ONE kernel launch (plus a fill when the peaks restart).
// `peaks` / `aq.avg` are device pointers or NULL.
static int spectrum_device(rfa_spectrum_plan *pl, const void *iq, long long nframes, float *rows, long long row0,
                           long long row_step, long long ring_rows, long long row_stride, long long store_from,
                           float *peaks, bool accumulate, AvgReq aq) {
    rfa_ctx *c = pl->ctx;
    const int n = pl->d.fft_size;
    SpectrumLaunch L{};
    L.N = n;
    L.in_fmt = pl->d.format;
    L.out_kind = OUT_DB;
    L.stream = c->stream;
    L.num_sms = c->num_sms;
    L.tune = c->tune;
    L.p.in = iq;
    L.p.win = pl->win;
    L.p.tw = pl->tw;
    L.p.twN = pl->twN;
    L.p.rows = rows;
    L.p.row0 = row0;
    L.p.row_step = row_step;
    L.p.ring_rows = ring_rows;
    L.p.row_stride = row_stride;
    L.p.nframes = nframes;
    L.p.store_from = store_from;
    L.p.inv_n2 = -3.0102999566398120f * log2f((float)n);  // dB bias, see logmag_db
    const bool ema = pl->d.avg_mode == RFA_AVG_EMA && aq.avg != nullptr;
    L.p.peaks = peaks;
    L.p.avg = ema ? nullptr : aq.avg;  // the exponential average is a row pass of its own after the transform
    L.p.avg_newest = row0 + (nframes - 1) * row_step;
    L.p.avg_dir = -row_step;
    L.p.avg_valid = aq.valid;
    L.p.avg_len = pl->d.avg_len;
    L.p.ticket = pl->ticket.as<unsigned int>();
    if (peaks && !accumulate) {  // FftProcessor.kt:236: "no peak" is -999999f
        fill_f32(peaks, (size_t)n, -999999.0f, c->stream);
        RFA_CK(cudaGetLastError());
        c->launches++;
    }
    if (fourstep_supported(n, pl->d.format, OUT_DB, c->tune)) {
        // N = 32768 / 65536: column transforms, twiddle, row transforms through an intermediate buffer
        // (fourstep_kernel.cuh); the time average is a row reduction afterwards.  Measured on B200 the fixed cost
        // of a launch pair (persistent-CTA prologue and tail) outweighs keeping the intermediate inside L2:
        // 2^24 samples take 92 / 99 us as one 128 MiB batch, 109 / 119 us as four 32 MiB batches.
        FourStepLaunch fs{};
        if (int rc = c->get_twiddles(n / 256, &fs.tw_n1)) return rc;
        if (int rc = c->get_twiddles(256, &fs.tw_256)) return rc;
        fs.tw_n = pl->twN;
        // first choice: thread-block clusters, the intermediate in distributed shared memory (one launch, no Z buffer)
        cudaError_t e4 = cudaErrorNotSupported;
        if (c->tune.cluster) {
            if (int rc = c->get_twiddles(kColumnTwiddles + n, &fs.tz)) return rc;
            e4 = fourstep_cluster_launch(L, fs);
            if (e4 == cudaSuccess) c->launches += fourstep_launches(n, nframes, 0);
        }
        if (e4 == cudaErrorNotSupported) {
        long long want = c->tune.fs_batch_kib << 10;  // knob "fs_batch_kib"; default: 128 MiB per batch
        const long long all = nframes * (long long)n * (long long)sizeof(cf);
        if (want > all) want = all;
        if (want < (long long)n * (long long)sizeof(cf)) want = (long long)n * (long long)sizeof(cf);
        if (int rc = pl->zbuf.ensure((size_t)want)) return rc;
        if (int rc = pl->fsync.ensure(2 * (size_t)nframes * sizeof(unsigned int))) return rc;
        fs.sync = pl->fsync.as<unsigned int>();
        fs.z = pl->zbuf.as<cf>();
        fs.z_bytes = want;
        e4 = fourstep_launch(L, fs);
        if (e4 == cudaSuccess) c->launches += fourstep_launches(n, nframes, fs.z_bytes);
        }
        if (e4 != cudaSuccess) return cuda_fail(e4, "four-step spectrum kernels");
        if (aq.avg && !ema) {
            average_rows(rows, L.p.avg_newest, L.p.avg_dir, ring_rows, row_stride, aq.valid, pl->d.avg_len, n, aq.avg, c->stream);
            RFA_CK(cudaGetLastError());
            c->launches++;
        }
    } else {
        cudaError_t e = spectrum_launch(L);
        if (e != cudaSuccess) return cuda_fail(e, "spectrum kernel");
        c->launches++;
    }
    if (ema) {
        // frames of this call that are stored and still carry weight, oldest first
        long long first = nframes - pl->ema_frames;
        if (first < store_from) first = store_from;
        if (first < 0) first = 0;
        ema_rows(rows, row0, row_step, ring_rows, row_stride, first, nframes - 1, pl->d.ema_alpha,
                 aq.accumulate && first == 0, n, aq.avg, c->stream);
        RFA_CK(cudaGetLastError());
        c->launches++;
    }
    return RFA_OK;
}

extern "C" {

int rfa_spectrum_plan_create(rfa_ctx *c, const rfa_spectrum_desc *d, rfa_spectrum_plan **out) {
    RFA_REQUIRE(c && d && out, "rfa_spectrum_plan_create: NULL argument");
    *out = nullptr;
    RFA_REQUIRE(d->format >= RFA_FMT_S8 && d->format <= RFA_FMT_S16LE, "unknown sample format %d", d->format);
    RFA_REQUIRE(is_pow2(d->fft_size) && d->fft_size >= 16 && d->fft_size <= 65536,
                "FFT size %d unsupported: need a power of two in 16..65536", d->fft_size);
    RFA_REQUIRE(d->window >= RFA_WIN_BLACKMAN_REF && d->window <= RFA_WIN_RECT, "unknown window %d", d->window);
    RFA_REQUIRE(d->avg_len >= 0 && d->avg_len <= 30, "avg_len %d outside 0..30 (DisplayTab.kt:202-212)", d->avg_len);
    RFA_REQUIRE(d->avg_mode == RFA_AVG_BOXCAR || d->avg_mode == RFA_AVG_EMA, "unknown averaging mode %d", d->avg_mode);
    RFA_REQUIRE(d->avg_mode != RFA_AVG_EMA || (d->ema_alpha > 0.0f && d->ema_alpha <= 1.0f),
                "ema_alpha %g outside (0, 1]", (double)d->ema_alpha);
    if (int rc = c->use()) return rc;
    rfa_spectrum_plan *pl = new rfa_spectrum_plan();
    pl->ctx = c;
    pl->d = *d;
    if (d->avg_mode == RFA_AVG_EMA) {
        // (1-alpha)^K < 2^-40: older frames cannot move a float32 result
        const double keep = 1.0 - (double)d->ema_alpha;
        pl->ema_frames = keep <= 0.0 ? 1 : (long long)ceil(40.0 * log(2.0) / -log(keep)) + 1;
    }
    const int n = d->fft_size, nl = n > 16384 ? 16384 : n;
    int rc = c->get_twiddles(nl, &pl->tw);
    if (!rc && (n > 16384 || n == 4096)) rc = c->get_twiddles(-n, &pl->twN);  // 4096: the 64 x 64 kernel's table
    if (!rc) rc = c->get_window(d->window, n, &pl->win);
    if (!rc) rc = pl->ticket.ensure(TICKET_WORDS * sizeof(unsigned int));
    if (!rc && cudaMemsetAsync(pl->ticket.p, 0, TICKET_WORDS * sizeof(unsigned int), c->stream) != cudaSuccess) rc = RFA_ERR_CUDA;
    if (rc) {
        delete pl;
        return rc;
    }
    *out = pl;
    return RFA_OK;
}

int rfa_spectrum_plan_info(const rfa_spectrum_plan *pl, int *fft_size, int *format, int *avg_len) {
    RFA_REQUIRE(pl != nullptr, "rfa_spectrum_plan_info: NULL");
    if (fft_size) *fft_size = pl->d.fft_size;
    if (format) *format = pl->d.format;
    if (avg_len) *avg_len = pl->d.avg_len;
    return RFA_OK;
}

int rfa_spectrum_plan_destroy(rfa_spectrum_plan *pl) {
    if (!pl) return RFA_OK;
    cudaSetDevice(pl->ctx->device);
    cudaStreamSynchronize(pl->ctx->stream);
    pl->tail.release();
    pl->ticket.release();
    pl->zbuf.release();
    pl->fsync.release();
    pl->dpeaks.release();
    pl->davg.release();
    for (int i = 0; i < 2; i++) {
        pl->din[i].release();
        pl->drows[i].release();
    }
    delete pl;
    return RFA_OK;
}

long long rfa_spectrum_algorithmic_bytes(const rfa_spectrum_plan *pl, long long nframes, int rows_stored) {
    if (!pl) return 0;
    const long long n = pl->d.fft_size;
    return nframes * n * (fmt_bytes(pl->d.format) + (rows_stored ? 4 : 0)) + 2 * n * 4;
}

int rfa_spectrum_process(rfa_spectrum_plan *pl, const void *iq, long long nframes, const rfa_spectrum_out *o,
                         int mem) {
    RFA_REQUIRE(pl && o, "rfa_spectrum_process: NULL argument");
    RFA_REQUIRE(nframes >= 0, "negative frame count");
    rfa_ctx *c = pl->ctx;
    const int n = pl->d.fft_size, L = pl->d.avg_len;
    const int bps = fmt_bytes(pl->d.format);
    if (nframes == 0) return RFA_OK;
    RFA_REQUIRE(iq != nullptr, "iq is NULL");
    RFA_REQUIRE(!o->rows || o->row_stride >= n, "row_stride %lld smaller than the FFT size", o->row_stride);
    RFA_REQUIRE(!o->rows || o->row_step == 1 || o->row_step == -1, "row_step must be +1 or -1");
    RFA_REQUIRE(o->ring_rows >= 0, "negative ring_rows");
    RFA_REQUIRE(!o->rows || o->ring_rows > 0 || o->row_step == 1, "a linear row buffer needs row_step == +1");
    if (int rc = c->use()) return rc;
    const bool want_peaks = o->peaks != nullptr && pl->d.peak_hold;
    const bool want_avg = o->avg != nullptr;

    const bool ema = pl->d.avg_mode == RFA_AVG_EMA;
    if (mem == RFA_MEM_DEVICE) {
        RFA_REQUIRE(((uintptr_t)iq % (bps == 4 ? 4 : 2)) == 0, "iq pointer misaligned for its format");
        float *rows = o->rows;
        long long row0 = o->row0, step = o->row_step, ring = o->ring_rows, stride = o->row_stride;
        long long store_from = 0;
        if (!rows) {
            if (!want_avg && !want_peaks) return RFA_OK;  // nothing observable is requested
            // keep just the rows the average needs: the newest L+1, or the exponential average's window
            long long keep = L + 1;
            if (ema && want_avg) {
                const long long cap = (256LL << 20) / ((long long)n * (long long)sizeof(float));  // at most 256 MiB of rows
                keep = pl->ema_frames < nframes ? pl->ema_frames : nframes;
                if (keep > cap && cap >= 1) {
                    // a window longer than the scratch ring: walk the call in pieces, the average carried between them
                    rfa_spectrum_out piece = *o;
                    for (long long f0 = 0; f0 < nframes; f0 += cap) {
                        const long long nf = nframes - f0 < cap ? nframes - f0 : cap;
                        piece.peaks_accumulate = o->peaks_accumulate || f0 > 0;
                        piece.avg_accumulate = o->avg_accumulate || f0 > 0;
                        if (int rc = rfa_spectrum_process(pl, (const char *)iq + (size_t)f0 * n * bps, nf, &piece, mem)) return rc;
                    }
                    return RFA_OK;
                }
            }
            if (int rc = pl->tail.ensure((size_t)keep * n * sizeof(float))) return rc;
            rows = pl->tail.as<float>();
            row0 = 0;
            step = 1;
            ring = keep;
            stride = n;
            store_from = want_avg ? nframes - keep : nframes;
        }
        // more frames than ring rows: the older ones would be overwritten by newer frames of this
        // very call (FftProcessor.kt:224-229 runs sequentially), so only the newest ring_rows
        // frames are stored -- the frames are transformed concurrently, order must not matter
        if (ring > 0 && nframes > ring && store_from < nframes - ring) store_from = nframes - ring;
        AvgReq aq;
        if (want_avg) {
            aq.avg = o->avg;
            aq.valid = nframes;
            aq.accumulate = o->avg_accumulate != 0;
            if (o->rows && o->ring_rows > 0) {
                aq.valid = o->history_rows + nframes;
                if (aq.valid > o->ring_rows) aq.valid = o->ring_rows;
            }
        }
        return spectrum_device(pl, iq, nframes, rows, row0, step, ring, stride, store_from,
                               want_peaks ? o->peaks : nullptr, o->peaks_accumulate != 0, aq);
    }

    // ---- host buffers: chunked H2D -> kernel -> D2H pipeline over three streams ----
    RFA_REQUIRE(mem == RFA_MEM_HOST, "unknown memory space %d", mem);
    RFA_REQUIRE(!o->rows || o->ring_rows == 0, "ring-buffer rows must be device resident");
    const size_t frame_in = (size_t)n * bps, frame_out = (size_t)n * sizeof(float);
    // IQ bytes per chunk: small enough that the first copy in and the last copy out (which overlap with
    // nothing) are short, large enough that the per-chunk launch and copy set-up costs stay hidden
    const size_t chunk_bytes = (size_t)c->tune.chunk_kib << 10;  // knob "chunk_kib", default 8 MiB
    long long cf_frames = (long long)(chunk_bytes / frame_in);
    if (cf_frames < L + 1) cf_frames = L + 1;
    if (cf_frames > nframes) cf_frames = nframes;
    // chunk schedule: a SHORT first chunk (an eighth), so that the copy out -- the engine that bounds this path: 4 bytes
    // of rows leave per 2 bytes of IQ -- starts after 15 us instead of 90
    std::vector<long long> sizes;
    {
        long long left = nframes;
        long long first = cf_frames / 8;
        if (first < L + 1) first = L + 1;
#ifdef RFA_NO_RAMP  // timing build: round 2's earlier schedule (one short chunk, whole chunks, the remainder)
        if (left >= first + 2 * cf_frames) {
            sizes.push_back(first);
            left -= first;
        }
        while (left >= 2 * cf_frames) {
            sizes.push_back(cf_frames);
            left -= cf_frames;
        }
        sizes.push_back(left);
        left = 0;
#endif
        if (left >= first + 2 * cf_frames) {
            // ... and a ramp behind it: while a chunk's rows leave (4 bytes per 2 bytes of IQ), the next chunk's IQ must
            // arrive and be transformed, or the copy-out engine idles -- after an eighth-size chunk a full-size one takes
            // four times as long to come in as the small one's rows take to leave.  Growth by 3/2 keeps it fed.
            long long sz = first;
            while (sz < cf_frames && left >= sz + 2 * cf_frames) {
                sizes.push_back(sz);
                left -= sz;
                sz += (sz + 1) / 2;
            }
        }
        // the rest in equal chunks of at most cf_frames (a long last chunk would starve the copy-out engine while it comes in)
        long long nrest = left > 0 ? (left + cf_frames - 1) / cf_frames : 0;
        while (nrest > 1 && left / nrest < L + 1) nrest--;
        for (long long i = 0; i < nrest; i++) {
            const long long take = left / (nrest - i);
            sizes.push_back(take);
            left -= take;
        }
        if (sizes.empty()) sizes.push_back(left);
    }
    const long long nchunks = (long long)sizes.size();
    long long max_chunk = 0;
    for (long long v : sizes) max_chunk = v > max_chunk ? v : max_chunk;
    const bool store_rows = o->rows != nullptr;
    for (int b = 0; b < 2; b++) {
        if (int rc = pl->din[b].ensure(max_chunk * frame_in)) return rc;
        if (store_rows || want_avg)
            if (int rc = pl->drows[b].ensure(max_chunk * frame_out)) return rc;
    }
    if (want_avg) {
        if (int rc = pl->davg.ensure(frame_out)) return rc;
        if (ema && o->avg_accumulate)
            RFA_CK(cudaMemcpyAsync(pl->davg.p, o->avg, frame_out, cudaMemcpyHostToDevice, c->stream));
    }
    float *dpeaks = nullptr;
    if (want_peaks) {
        if (int rc = pl->dpeaks.ensure(frame_out)) return rc;
        dpeaks = pl->dpeaks.as<float>();
        if (o->peaks_accumulate)
            RFA_CK(cudaMemcpyAsync(dpeaks, o->peaks, frame_out, cudaMemcpyHostToDevice, c->stream));
    }
    // no early return may leave copies into / out of the caller's buffers in flight
    struct DrainOnExit {
        rfa_ctx *c;
        ~DrainOnExit() {
            cudaStreamSynchronize(c->s_in);
            cudaStreamSynchronize(c->s_out);
            cudaStreamSynchronize(c->stream);
        }
    } drain{c};
    // make the side streams start after whatever is already queued on the context stream
    RFA_CK(cudaEventRecord(c->ev_k[0], c->stream));
    RFA_CK(cudaEventRecord(c->ev_k[1], c->stream));
    RFA_CK(cudaEventRecord(c->ev_out[0], c->stream));
    RFA_CK(cudaEventRecord(c->ev_out[1], c->stream));
    long long done = 0;
    for (long long i = 0; i < nchunks; i++) {
        const int b = (int)(i & 1);
        const long long frames = sizes[(size_t)i];
        RFA_CK(cudaStreamWaitEvent(c->s_in, c->ev_k[b], 0));  // kernel that last read din[b] is done
        RFA_CK(cudaMemcpyAsync(pl->din[b].p, (const char *)iq + (size_t)done * frame_in, (size_t)frames * frame_in,
                               cudaMemcpyHostToDevice, c->s_in));
        RFA_CK(cudaEventRecord(c->ev_in[b], c->s_in));
        RFA_CK(cudaStreamWaitEvent(c->stream, c->ev_in[b], 0));
        RFA_CK(cudaStreamWaitEvent(c->stream, c->ev_out[b], 0));  // drows[b] has been copied out
        float *drows = (store_rows || want_avg) ? pl->drows[b].as<float>() : nullptr;
        long long store_from = store_rows || (ema && want_avg) ? 0 : (want_avg && i == nchunks - 1 ? frames - (L + 1) : frames);
        if (drows || dpeaks) {
            float *rows_arg = drows ? drows : pl->din[b].as<float>();  // never written when store_from == frames
            AvgReq aq;
            if (want_avg && (ema || i == nchunks - 1)) {  // the exponential average walks every chunk, carried in davg
                aq.avg = pl->davg.as<float>();
                aq.valid = nframes < frames ? nframes : frames;
                aq.accumulate = o->avg_accumulate != 0 || i > 0;
            }
            if (int rc = spectrum_device(pl, pl->din[b].p, frames, rows_arg, 0, 1, 0, n, store_from, dpeaks,
                                         o->peaks_accumulate != 0 || i > 0, aq))
                return rc;
        }
        RFA_CK(cudaEventRecord(c->ev_k[b], c->stream));
        if (store_rows) {
            RFA_CK(cudaStreamWaitEvent(c->s_out, c->ev_k[b], 0));
            if (o->row_stride == n) {
                RFA_CK(cudaMemcpyAsync(o->rows + (size_t)(o->row0 + done) * n, drows, (size_t)frames * frame_out,
                                       cudaMemcpyDeviceToHost, c->s_out));
            } else {
                RFA_CK(cudaMemcpy2DAsync(o->rows + (size_t)(o->row0 + done) * o->row_stride,
                                         (size_t)o->row_stride * sizeof(float), drows, frame_out, frame_out,
                                         (size_t)frames, cudaMemcpyDeviceToHost, c->s_out));
            }
            RFA_CK(cudaEventRecord(c->ev_out[b], c->s_out));
        }
        done += frames;
    }
    if (want_avg) RFA_CK(cudaMemcpyAsync(o->avg, pl->davg.p, frame_out, cudaMemcpyDeviceToHost, c->stream));
    if (want_peaks) RFA_CK(cudaMemcpyAsync(o->peaks, dpeaks, frame_out, cudaMemcpyDeviceToHost, c->stream));
    RFA_CK(cudaStreamSynchronize(c->s_out));
    RFA_CK(cudaStreamSynchronize(c->stream));
    return RFA_OK;
}

/* ---------------------------------------------------------------- row reductions ---- */
int rfa_average_rows(rfa_ctx *c, const float *rows, long long newest, long long dir, long long ring_rows,
                     long long row_stride, long long valid, int avg_len, int n, float *avg, int mem_rows,
                     int mem_avg) {
    RFA_REQUIRE(c && rows && avg, "rfa_average_rows: NULL argument");
    RFA_REQUIRE(mem_rows == RFA_MEM_DEVICE, "rows must be device resident");
    RFA_REQUIRE(avg_len >= 0 && avg_len <= 30 && n > 0, "bad averaging request");
    if (int rc = c->use()) return rc;
    float *davg = avg;
    if (mem_avg == RFA_MEM_HOST) {
        if (int rc = c->stage[4].ensure((size_t)n * sizeof(float))) return rc;
        davg = c->stage[4].as<float>();
    }
    average_rows(rows, newest, dir, ring_rows, row_stride, valid, avg_len, n, davg, c->stream);
    RFA_CK(cudaGetLastError());
    c->launches++;
    if (mem_avg == RFA_MEM_HOST) {
        RFA_CK(cudaMemcpyAsync(avg, davg, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    return RFA_OK;
}

int rfa_ema_rows(rfa_ctx *c, const float *rows, long long row0, long long row_step, long long ring_rows,
                 long long row_stride, long long first, long long last, float alpha, int from_state, int n, float *avg,
                 int mem_avg) {
    RFA_REQUIRE(c && rows && avg, "rfa_ema_rows: NULL argument");
    RFA_REQUIRE(n > 0 && alpha > 0.0f && alpha <= 1.0f && first >= 0 && last >= first, "bad averaging request");
    if (int rc = c->use()) return rc;
    float *davg = avg;
    if (mem_avg == RFA_MEM_HOST) {
        if (int rc = c->stage[4].ensure((size_t)n * sizeof(float))) return rc;
        davg = c->stage[4].as<float>();
        if (from_state) RFA_CK(cudaMemcpyAsync(davg, avg, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    ema_rows(rows, row0, row_step, ring_rows, row_stride, first, last, alpha, from_state != 0, n, davg, c->stream);
    RFA_CK(cudaGetLastError());
    c->launches++;
    if (mem_avg == RFA_MEM_HOST) {
        RFA_CK(cudaMemcpyAsync(avg, davg, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    return RFA_OK;
}

int rfa_channel_bins(int n, long long frequency, int sample_rate, long long chan_start, long long chan_end,
                     int *bin_start, int *bin_end) {
    RFA_REQUIRE(n > 0 && sample_rate > 0 && bin_start && bin_end, "rfa_channel_bins: bad argument");
    design::channel_bins(n, frequency, sample_rate, chan_start, chan_end, bin_start, bin_end);
    return RFA_OK;
}

// Java's (int) cast of a double: NaN -> 0, saturating
static int java_int(double v) {
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return -2147483647 - 1;
    return (int)v;
}

int rfa_render_waterfall(rfa_ctx *c, const rfa_render_desc *d, const float *rows, const float *peaks,
                         const uint32_t *colormap, int colormap_size, int colormap_mem, uint32_t *argb,
                         int *color_index, float *time_average, float *peaks_y, int out_mem) {
    RFA_REQUIRE(c && d && rows, "rfa_render_waterfall: NULL argument");
    RFA_REQUIRE(d->fft_size > 0 && d->sample_rate > 0 && d->width > 0 && d->ring_rows > 0, "bad render geometry");
    RFA_REQUIRE(d->first_row >= 0 && d->nrows >= 0 && d->first_row + d->nrows <= d->ring_rows,
                "row range %d+%d outside the ring of %d rows", d->first_row, d->nrows, d->ring_rows);
    RFA_REQUIRE(d->avg_len >= 0 && d->avg_len <= 30, "avg_len %d outside 0..30", d->avg_len);
    RFA_REQUIRE(colormap_size > 0 && (colormap || !argb), "a colour map is needed for ARGB output");
    RFA_REQUIRE(!time_average || (d->first_row == 0 && d->nrows > d->avg_len),
                "the FFT trace needs the newest avg_len+1 rows in the rendered range");
    if (d->nrows == 0) return RFA_OK;
    if (int rc = c->use()) return rc;
    // viewport arithmetic, AnalyzerSurface.kt:647-676 (Float / Double exactly as written there)
    const int fftSize = d->fft_size, width = d->width;
    const float samplesPerHz = (float)fftSize / (float)d->sample_rate;
    const long long frequencyDiff = d->viewport_frequency - d->frequency;
    const long long sampleRateDiff = d->viewport_sample_rate - (long long)d->sample_rate;
    const int start = java_int(((double)frequencyDiff - sampleRateDiff / 2.0) * samplesPerHz);
    const int end = fftSize + java_int(((double)frequencyDiff + sampleRateDiff / 2.0) * samplesPerHz);
    const float samplesPerPx = (float)(end - start) / (float)width;
    const float dbDiff = d->max_db - d->min_db;
    RenderDesc r;
    r.rows = rows;
    r.row_stride = d->row_stride > 0 ? d->row_stride : fftSize;
    r.ring_rows = d->ring_rows;
    r.n = fftSize;
    r.newest = d->newest_row;
    r.first_row = d->first_row;
    r.nrows = d->nrows;
    r.peaks = peaks;
    r.width = width;
    r.start = start;
    r.samples_per_px = samplesPerPx;
    r.first_pixel = start >= 0 ? 0 : java_int((double)((float)(start * -1) / samplesPerPx));
    r.last_pixel = end >= fftSize ? java_int((double)((float)(fftSize - start) / samplesPerPx))
                                  : java_int((double)((float)(end - start) / samplesPerPx));
    r.min_db = d->min_db;
    r.db_width = (float)d->fft_height / dbDiff;
    r.scale = (float)colormap_size / dbDiff;
    r.fft_height = (float)d->fft_height;
    r.colormap_size = colormap_size;
    r.avg_len = d->avg_len;
    if (colormap) {
        if (colormap_mem == RFA_MEM_HOST) {
            if (int rc = c->render[0].ensure((size_t)colormap_size * sizeof(uint32_t))) return rc;
            RFA_CK(cudaMemcpyAsync(c->render[0].p, colormap, (size_t)colormap_size * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice, c->stream));
            r.colormap = c->render[0].as<uint32_t>();
        } else {
            r.colormap = colormap;
        }
    }
    const size_t px = (size_t)d->ring_rows * width;
    const bool host = out_mem == RFA_MEM_HOST;
    if (time_average) {
        if (int rc = c->render[1].ensure((size_t)(d->avg_len + 1) * width * sizeof(float))) return rc;
        r.row_means = c->render[1].as<float>();
    }
    auto out_buf = [&](int slot, void *user, size_t bytes, void **dev) -> int {
        *dev = user;
        if (user && host) {
            if (int rc = c->render[slot].ensure(bytes)) return rc;
            *dev = c->render[slot].p;
        }
        return RFA_OK;
    };
    void *dv;
    if (int rc = out_buf(2, argb, px * sizeof(uint32_t), &dv)) return rc;
    r.argb = (uint32_t *)dv;
    if (int rc = out_buf(3, color_index, px * sizeof(int), &dv)) return rc;
    r.color_index = (int *)dv;
    if (int rc = out_buf(4, time_average, (size_t)width * sizeof(float), &dv)) return rc;
    r.time_average = (float *)dv;
    if (int rc = out_buf(5, peaks_y, (size_t)width * sizeof(float), &dv)) return rc;
    r.peaks_y = (float *)dv;
    if (host) {  // rows outside the rendered range keep what the caller has: bring them along
        if (argb) RFA_CK(cudaMemcpyAsync(r.argb, argb, px * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        if (color_index) RFA_CK(cudaMemcpyAsync(r.color_index, color_index, px * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    cudaError_t e = render_launch(r, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "render kernels");
    c->launches += time_average ? 2 : 1;
    if (host) {
        if (argb) RFA_CK(cudaMemcpyAsync(argb, r.argb, px * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        if (color_index) RFA_CK(cudaMemcpyAsync(color_index, r.color_index, px * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        if (time_average) RFA_CK(cudaMemcpyAsync(time_average, r.time_average, (size_t)width * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        if (peaks_y) RFA_CK(cudaMemcpyAsync(peaks_y, r.peaks_y, (size_t)width * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    return RFA_OK;
}

int rfa_channel_strength(rfa_ctx *c, const float *rows, long long row0, long long row_step, long long ring_rows,
                         long long row_stride, long long nrows, int bin_start, int bin_end, float *out,
                         int mem_out) {
    RFA_REQUIRE(c && rows && out, "rfa_channel_strength: NULL argument");
    RFA_REQUIRE(bin_end > bin_start && bin_start >= 0, "empty channel bin range");
    if (nrows <= 0) return RFA_OK;
    if (int rc = c->use()) return rc;
    float *dout = out;
    if (mem_out == RFA_MEM_HOST) {
        if (int rc = c->stage[4].ensure((size_t)nrows * sizeof(float))) return rc;
        dout = c->stage[4].as<float>();
    }
    channel_strength(rows, row0, row_step, ring_rows, row_stride, nrows, bin_start, bin_end, dout, c->stream);
    RFA_CK(cudaGetLastError());
    c->launches++;
    if (mem_out == RFA_MEM_HOST) {
        RFA_CK(cudaMemcpyAsync(out, dout, (size_t)nrows * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    return RFA_OK;
}

int rfa_shift_rows(rfa_ctx *c, float *rows, long long nrows, long long row_stride, int n, int shift) {
    RFA_REQUIRE(c && rows, "rfa_shift_rows: NULL argument");
    RFA_REQUIRE(n > 0 && n <= 65536 && row_stride >= n, "bad row geometry");
    if (int rc = c->use()) return rc;
    shift_rows(rows, nrows, row_stride, n, shift, c->stream);
    RFA_CK(cudaGetLastError());
    c->launches++;
    return RFA_OK;
}

int rfa_synth_iq(rfa_ctx *c, int fmt, uint32_t seed, const rfa_synth_comp *comps, int ncomp, int noise_shift,
                 long long first_sample, long long nsamples, void *out, int mem) {
    RFA_REQUIRE(c && out, "rfa_synth_iq: NULL argument");
    RFA_REQUIRE(fmt >= RFA_FMT_S8 && fmt <= RFA_FMT_S16LE, "unknown sample format %d", fmt);
    RFA_REQUIRE(ncomp >= 0 && ncomp <= 8 && (ncomp == 0 || comps), "0..8 components");
    RFA_REQUIRE(noise_shift >= 0 && noise_shift < 16 && first_sample >= 0 && nsamples >= 0, "bad generator range");
    if (nsamples == 0) return RFA_OK;
    if (int rc = c->use()) return rc;
    if (!c->synth_table) {
        short host[4096];
        synth_make_table(host);
        RFA_CK(cudaMalloc(&c->synth_table, sizeof(host)));
        RFA_CK(cudaMemcpyAsync(c->synth_table, host, sizeof(host), cudaMemcpyHostToDevice, c->stream));
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    SynthComp cc[8];
    for (int i = 0; i < ncomp; i++) cc[i] = SynthComp{comps[i].step, comps[i].amp, comps[i].mod_step, comps[i].mod_k};
    const size_t bytes = (size_t)nsamples * fmt_bytes(fmt);
    void *dout = out;
    if (mem == RFA_MEM_HOST) {
        if (int rc = c->stage[0].ensure(bytes)) return rc;
        dout = c->stage[0].p;
    } else {
        RFA_REQUIRE(((uintptr_t)out & 3) == 0, "device output must be 4-byte aligned");
    }
    cudaError_t e = synth_launch(fmt, seed, cc, ncomp, noise_shift, (unsigned long long)first_sample, nsamples,
                                 c->synth_table, dout, c->num_sms, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "synth kernel");
    c->launches++;
    if (mem == RFA_MEM_HOST) {
        RFA_CK(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, c->stream));
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    return RFA_OK;
}

int rfa_fill(rfa_ctx *c, float *dst, long long count, float value) {
    RFA_REQUIRE(c && dst && count >= 0, "rfa_fill: bad argument");
    if (int rc = c->use()) return rc;
    fill_f32(dst, (size_t)count, value, c->stream);
    RFA_CK(cudaGetLastError());
    c->launches++;
    return RFA_OK;
}

}  // extern "C"
