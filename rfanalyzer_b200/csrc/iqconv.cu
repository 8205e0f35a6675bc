// iqconv.cu -- the Airspy / HydraSDR real -> IQ converter on the GPU:
// libairspy/src/main/cpp/libairspy/iqconverter_int16.c:54-208, bit-exact.
//
//   remove_dc        (:160-186)  y = (x - x') + ((e + 32100*y') >> 15), e = low 15 bits: a first-order DC blocker with
//                                error feedback, all int16 / int32 with the reference's wrap-around
//   translate_fs_4   (:188-202)  samples 0..3 of every group of four: -s, (-s) >> 1, s, s >> 1
//   fir_interleaved  (:97-134)   even samples: half-band FIR (every other tap of the kernel), int32 sum >> 15
//   delay_interleaved(:136-158)  odd samples: delayed by len/4 samples
//
// The FIR, the delay and the translation are data parallel.  The DC blocker is a NON-LINEAR recurrence (the
// truncating shift feeds back), so no scan reproduces it bit for bit; its state, though, is one integer
// T = 2^15*y + e with T' = T - 668*floor(T / 2^15) + 2^15*w, and two runs over the same input whose states differ by a
// multiple of 668 approach each other monotonically and MERGE (the difference shrinks by 668 whenever the floors
// differ, about every 49 samples per e-fold).  T mod 668 is known in closed form from the input alone
// (T == T0 + 2^15*(x - x0) mod 668 while nothing wraps), so:
//
//   dc_speculate_kernel : one thread per 512-sample chunk starts 2048 samples early from y = 0 and the right residue,
//                         runs the exact integer recurrence, keeps the chunk's outputs and its start / end states
//   dc_verify_kernel    : chunk c is exact if its start state equals the end state of chunk c-1 (chunk 0 starts from the
//                         carried state); all comparisons in parallel
//   dc_repair_kernel    : one thread re-runs, in order, only the chunks that failed the comparison (int16 wrap-around
//                         of y or w shifts the residue; exactly constant input parks the state in the blocker's dead
//                         zone where nothing merges -- those chunks are recognised and skipped in O(1))
//
// so the result is the sequential one for EVERY input, and typical signals never reach the repair kernel.
#include <cuda_runtime.h>
#include <stdint.h>

#include "capi_core.h"

using namespace rfa;

namespace {

constexpr int kChunk = 512;    // samples per speculative chunk
constexpr int kWarm = 2048;    // warm-up samples before a chunk (merge probability: see DESIGN.md)
constexpr int kMaxTaps = 64;   // cnv->len = len/2 + 1 <= 64 (the Airspy kernel has 24)

struct DcState {
    int x, y, e;  // old_x, old_y, old_e
};

// One sample of remove_dc (iqconverter_int16.c:172-183).  |u| < 2^31 and u >> 15 always fits an int16 (|old_y| <= 2^15,
// 0 <= old_e < 2^15), so the reference's `s = u >> 15` truncation never bites and `old_e = u - (s << 15)` is the low
// 15 bits of u; the two int16 wrap-arounds that do bite (w and y) are kept.
__host__ __device__ __forceinline__ int dc_step(int x, DcState &s) {
    const int w = (int)(short)(x - s.x);
    const int u = s.e + s.y * 32100;
    const int y = (int)(short)(w + (u >> 15));
    s.e = u & 0x7FFF;
    s.x = x;
    s.y = y;
    return y;
}

struct ChunkInfo {
    DcState start, end;  // state when the chunk proper begins / after its last sample
    int constant;        // every sample of the chunk equals the sample before it (w == 0 throughout)
    int pad;
};

__device__ __forceinline__ int pos_mod(long long v, int m) {
    int r = (int)(v % m);
    return r < 0 ? r + m : r;
}

// samples: this call's input (len), state: carried DcState at sample 0
__global__ void __launch_bounds__(128) dc_speculate_kernel(const short *__restrict__ x, long long len, const DcState *__restrict__ state,
                                                           short *__restrict__ y, ChunkInfo *__restrict__ info, int nchunks, int vec) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const long long begin = (long long)c * kChunk;
    const long long end = begin + kChunk < len ? begin + kChunk : len;
    const DcState s0 = *state;
    DcState s;
    long long n = begin - kWarm;
    if (n <= 0) {  // the warm-up reaches the start of the call: run from the carried (exact) state
        n = 0;
        s = s0;
    } else {
        // guessed state: y = 0, e = the residue T must have (mod 668) when nothing wrapped since the call began
        const long long t0 = (long long)s0.y * 32768 + s0.e;
        const int xp = x[n - 1];
        s.x = xp;
        s.y = 0;
        s.e = pos_mod(t0 + 32768LL * ((long long)xp - s0.x), 668);
    }
    // Eight samples per 128-bit access (chunk and warm-up boundaries are multiples of 8 samples; the buffers are
    // 16-byte aligned or `vec` is false) -- one 2-byte access per step would cost a memory wavefront per lane -- and
    // four such loads in flight ahead of the dependent chain: two warps per scheduler hide no latency by themselves.
    ChunkInfo ci;
    ci.start = s;
    int moved = 0;
    if (vec) {
        const long long stop = end - ((end - n) & 7);  // whole groups of eight from n
        constexpr int DEPTH = 4;
        uint4 q[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; d++)
            if (n + 8 * d < stop) q[d] = __ldg(reinterpret_cast<const uint4 *>(x + n + 8 * d));
        while (n < stop) {
#pragma unroll
            for (int d = 0; d < DEPTH; d++) {
                if (n < stop) {
                    const uint4 v = q[d];
                    if (n + 8 * DEPTH < stop) q[d] = __ldg(reinterpret_cast<const uint4 *>(x + n + 8 * DEPTH));
                    const unsigned int wd[4] = {v.x, v.y, v.z, v.w};
                    if (n == begin) ci.start = s;
                    if (n < begin) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            dc_step((int)(short)(wd[k] & 0xFFFFu), s);
                            dc_step((int)(short)(wd[k] >> 16), s);
                        }
                    } else {
                        unsigned int o[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int a = (int)(short)(wd[k] & 0xFFFFu), b = (int)(short)(wd[k] >> 16);
                            moved |= (a ^ s.x) | (b ^ a);
                            const unsigned int ya = (unsigned int)dc_step(a, s) & 0xFFFFu;
                            const unsigned int yb = (unsigned int)dc_step(b, s) & 0xFFFFu;
                            o[k] = ya | (yb << 16);
                        }
                        *reinterpret_cast<uint4 *>(y + n) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                    n += 8;
                }
            }
        }
    }
    for (; n < end; n++) {  // unaligned buffers, and the last few samples of a call
        if (n == begin) ci.start = s;
        const int xv = x[n];
        if (n >= begin) {
            moved |= xv ^ s.x;
            y[n] = (short)dc_step(xv, s);
        } else {
            dc_step(xv, s);
        }
    }
    ci.end = s;
    ci.constant = moved == 0;
    ci.pad = 0;
    info[c] = ci;
}

__device__ __forceinline__ bool same(const DcState &a, const DcState &b) { return a.x == b.x && a.y == b.y && a.e == b.e; }

// flags[c] = 1 when chunk c did not start from the end state of chunk c-1; counters[0] += number of such chunks
__global__ void dc_verify_kernel(const ChunkInfo *__restrict__ info, int nchunks, unsigned char *__restrict__ flags,
                                 unsigned int *__restrict__ counters) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const bool bad = c > 0 && !same(info[c].start, info[c - 1].end);
    flags[c] = bad ? 1 : 0;
    if (bad) atomicAdd(counters, 1u);
}

// one thread: walk the flagged chunks in order, re-running each from the exact state; a repair whose end state differs
// from the speculative one invalidates the next chunk as well.  counters[1] += chunks re-run, [2] += chunks skipped
// through the dead-zone shortcut.  Leaves the exact end state of the call in *state.
__global__ void dc_repair_kernel(const short *__restrict__ x, long long len, DcState *__restrict__ state, short *__restrict__ y,
                                 ChunkInfo *__restrict__ info, int nchunks, unsigned char *__restrict__ flags,
                                 unsigned int *__restrict__ counters) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (counters[0] != 0) {
        int c = 1;
        while (c < nchunks) {
            if (!flags[c]) {
                c++;
                continue;
            }
            DcState s = info[c - 1].end;  // exact: every earlier chunk has been verified or repaired
            const long long begin = (long long)c * kChunk;
            const long long end = begin + kChunk < len ? begin + kChunk : len;
            const long long t = (long long)s.y * 32768 + s.e;
            if (info[c].constant && s.y == 0 && t >= 0 && t < 32768 && s.x == (int)x[begin]) {
                // dead zone of the blocker: constant input, y = 0: the state does not move and every output is 0.
                // The speculative run wrote zeros too when it entered the chunk in the dead zone; otherwise clear them.
                const DcState &sp = info[c].start;
                const long long ts = (long long)sp.y * 32768 + sp.e;
                if (!(sp.y == 0 && ts >= 0 && ts < 32768))
                    for (long long n = begin; n < end; n++) y[n] = 0;
                counters[2]++;
            } else {
                for (long long n = begin; n < end; n++) y[n] = (short)dc_step(x[n], s);
                counters[1]++;
            }
            const bool merged = same(s, info[c].end);
            info[c].end = s;
            if (!merged && c + 1 < nchunks) flags[c + 1] = 1;  // its start state was compared with a state that was not exact
            flags[c] = 0;
            c++;
        }
        counters[0] = 0;
    }
    *state = info[nchunks - 1].end;
}

struct FirParams {
    int taps;                 // cnv->len
    int delay;                // cnv->len >> 1
    int kernel[kMaxTaps];     // hb_kernel[2*i]
};

// translated even / odd sample k of this call (k >= 0) or of the history (k < 0)
__device__ __forceinline__ int z_even(const short *__restrict__ y, const short *__restrict__ hist_e, int nhist, long long k) {
    if (k < 0) return hist_e[nhist + k];
    const int v = y[2 * k];
    return (k & 1) ? v : (int)(short)(-v);
}
__device__ __forceinline__ int z_odd(const short *__restrict__ y, const short *__restrict__ hist_o, int nhist, long long k) {
    if (k < 0) return hist_o[nhist + k];
    const int v = y[2 * k + 1];
    return (k & 1) ? (int)(short)(v >> 1) : (int)(short)((-v) >> 1);
}

// one thread per output pair: out[2k] = FIR over the even samples, out[2k+1] = the odd sample `delay` pairs ago
__global__ void __launch_bounds__(256) translate_fir_delay_kernel(const short *__restrict__ y, long long pairs, const FirParams fp,
                                                                  const short *__restrict__ hist_e, const short *__restrict__ hist_o,
                                                                  short *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < pairs; k += stride) {
        unsigned int acc = 0;
#pragma unroll 4
        for (int j = 0; j < fp.taps; j++) acc += (unsigned int)(fp.kernel[j] * z_even(y, hist_e, kMaxTaps, k - j));
        const int i_out = (int)(short)((int)acc >> 15);
        const int q_out = z_odd(y, hist_o, kMaxTaps, k - fp.delay);
        reinterpret_cast<unsigned int *>(out)[k] = ((unsigned int)(unsigned short)q_out << 16) | (unsigned int)(unsigned short)i_out;
    }
}

// The same for a kernel of exactly TAPS taps (24 = the Airspy half-band), 1024 outputs per CTA: the translated even
// samples of the tile (plus TAPS-1 of history) are staged once in shared memory, the taps are compile-time indices
// into the parameter block, the sum wraps modulo 2^32 like the reference's int32 accumulator.
template <int TAPS>
__global__ void __launch_bounds__(256) translate_fir_delay_tiled_kernel(const short *__restrict__ y, long long pairs, const FirParams fp,
                                                                        const short *__restrict__ hist_e, const short *__restrict__ hist_o,
                                                                        short *__restrict__ out) {
    constexpr int TILE = 1024;
    __shared__ int ze[TILE + TAPS - 1];
    for (long long tile0 = (long long)blockIdx.x * TILE; tile0 < pairs; tile0 += (long long)gridDim.x * TILE) {
        for (int i = threadIdx.x; i < TILE + TAPS - 1; i += 256) {
            const long long k = tile0 - (TAPS - 1) + i;
            ze[i] = k < pairs ? z_even(y, hist_e, kMaxTaps, k) : 0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < TILE / 256; r++) {
            const int t = threadIdx.x + r * 256;
            const long long k = tile0 + t;
            if (k < pairs) {
                unsigned int acc = 0;
#pragma unroll
                for (int j = 0; j < TAPS; j++) acc += (unsigned int)(fp.kernel[j] * ze[t + TAPS - 1 - j]);
                const int i_out = (int)(short)((int)acc >> 15);
                const int q_out = z_odd(y, hist_o, kMaxTaps, k - fp.delay);
                reinterpret_cast<unsigned int *>(out)[k] = ((unsigned int)(unsigned short)q_out << 16) | (unsigned int)(unsigned short)i_out;
            }
        }
        __syncthreads();
    }
}

// the newest kMaxTaps translated even / odd samples become the next call's history
__global__ void history_kernel(const short *__restrict__ y, long long pairs, const short *__restrict__ old_e,
                               const short *__restrict__ old_o, short *__restrict__ new_e, short *__restrict__ new_o) {
    const int i = threadIdx.x;
    if (i >= kMaxTaps) return;
    const long long k = pairs - kMaxTaps + i;
    new_e[i] = (short)(k < -(long long)kMaxTaps ? 0 : (k < 0 ? old_e[kMaxTaps + k] : z_even(y, old_e, kMaxTaps, k)));
    new_o[i] = (short)(k < -(long long)kMaxTaps ? 0 : (k < 0 ? old_o[kMaxTaps + k] : z_odd(y, old_o, kMaxTaps, k)));
}

}  // namespace

struct rfa_iqconverter {
    rfa_ctx *ctx;
    FirParams fp;
    Buf state;       // DcState
    Buf hist[2][2];  // [buffer][even / odd] kMaxTaps shorts
    int cur = 0;
    Buf ybuf, info, flags, counters, stage;
    long long chunks_total = 0;
};

extern "C" {

int rfa_iqconverter_reset(rfa_iqconverter *cv) {
    RFA_REQUIRE(cv != nullptr, "rfa_iqconverter_reset: NULL");
    rfa_ctx *c = cv->ctx;
    if (int rc = c->use()) return rc;
    RFA_CK(cudaMemsetAsync(cv->state.p, 0, sizeof(DcState), c->stream));
    for (int b = 0; b < 2; b++)
        for (int k = 0; k < 2; k++) RFA_CK(cudaMemsetAsync(cv->hist[b][k].p, 0, kMaxTaps * sizeof(short), c->stream));
    RFA_CK(cudaMemsetAsync(cv->counters.p, 0, 4 * sizeof(unsigned int), c->stream));
    cv->cur = 0;
    cv->chunks_total = 0;
    return RFA_OK;
}

int rfa_iqconverter_destroy(rfa_iqconverter *cv) {
    if (!cv) return RFA_OK;
    cudaSetDevice(cv->ctx->device);
    cudaStreamSynchronize(cv->ctx->stream);
    cv->state.release();
    for (int b = 0; b < 2; b++)
        for (int k = 0; k < 2; k++) cv->hist[b][k].release();
    for (Buf *b : {&cv->ybuf, &cv->info, &cv->flags, &cv->counters, &cv->stage}) b->release();
    delete cv;
    return RFA_OK;
}

int rfa_iqconverter_create(rfa_ctx *c, const int16_t *hb_kernel, int len, rfa_iqconverter **out) {
    RFA_REQUIRE(c && hb_kernel && out, "rfa_iqconverter_create: NULL argument");
    *out = nullptr;
    RFA_REQUIRE(len >= 1 && len / 2 + 1 <= kMaxTaps, "half-band kernel of %d taps unsupported (at most %d)", len, 2 * kMaxTaps - 1);
    if (int rc = c->use()) return rc;
    rfa_iqconverter *cv = new rfa_iqconverter();
    cv->ctx = c;
    cv->fp.taps = len / 2 + 1;       // iqconverter_int16.c:60
    cv->fp.delay = cv->fp.taps >> 1;  // :144
    for (int i = 0; i < kMaxTaps; i++) cv->fp.kernel[i] = i < cv->fp.taps ? hb_kernel[i * 2] : 0;  // :72-75
    int rc = cv->state.ensure(sizeof(DcState));
    for (int b = 0; b < 2 && !rc; b++)
        for (int k = 0; k < 2 && !rc; k++) rc = cv->hist[b][k].ensure(kMaxTaps * sizeof(short));
    if (!rc) rc = cv->counters.ensure(4 * sizeof(unsigned int));
    if (!rc) rc = rfa_iqconverter_reset(cv);
    if (rc) {
        rfa_iqconverter_destroy(cv);
        return rc;
    }
    *out = cv;
    return RFA_OK;
}

int rfa_iqconverter_process(rfa_iqconverter *cv, int16_t *samples, long long len, int mem) {
    RFA_REQUIRE(cv != nullptr && len >= 0, "rfa_iqconverter_process: bad argument");
    if (len == 0) return RFA_OK;
    RFA_REQUIRE(samples != nullptr, "samples is NULL");
    RFA_REQUIRE(len % 4 == 0, "sample count %lld is not a multiple of 4 (translate_fs_4 walks groups of four)", len);
    RFA_REQUIRE(len / kChunk < 0x7FFFFF00LL, "too many samples in one call");
    rfa_ctx *c = cv->ctx;
    if (int rc = c->use()) return rc;
    short *dev = samples;
    if (mem == RFA_MEM_HOST) {
        if (int rc = cv->stage.ensure((size_t)len * sizeof(short))) return rc;
        dev = cv->stage.as<short>();
        RFA_CK(cudaMemcpyAsync(dev, samples, (size_t)len * sizeof(short), cudaMemcpyHostToDevice, c->stream));
    } else {
        RFA_REQUIRE(((uintptr_t)samples & 3) == 0, "device samples must be 4-byte aligned");
    }
    const int nchunks = (int)((len + kChunk - 1) / kChunk);
    if (int rc = cv->ybuf.ensure((size_t)len * sizeof(short))) return rc;
    if (int rc = cv->info.ensure((size_t)nchunks * sizeof(ChunkInfo))) return rc;
    if (int rc = cv->flags.ensure((size_t)nchunks)) return rc;
    short *y = cv->ybuf.as<short>();
    DcState *st = cv->state.as<DcState>();
    unsigned int *cnt = cv->counters.as<unsigned int>();
    const int vec = ((uintptr_t)dev & 15) == 0 && ((uintptr_t)y & 15) == 0;
    dc_speculate_kernel<<<(nchunks + 127) / 128, 128, 0, c->stream>>>(dev, len, st, y, cv->info.as<ChunkInfo>(), nchunks, vec);
    dc_verify_kernel<<<(nchunks + 255) / 256, 256, 0, c->stream>>>(cv->info.as<ChunkInfo>(), nchunks, cv->flags.as<unsigned char>(), cnt);
    dc_repair_kernel<<<1, 32, 0, c->stream>>>(dev, len, st, y, cv->info.as<ChunkInfo>(), nchunks, cv->flags.as<unsigned char>(), cnt);
    const long long pairs = len / 2;
    long long blocks = (pairs + 255) / 256;
    if (blocks > (long long)c->num_sms * 8) blocks = (long long)c->num_sms * 8;
    const int cur = cv->cur;
    if (cv->fp.taps == 24) {
        long long tiles = (pairs + 1023) / 1024;
        if (tiles > (long long)c->num_sms * 8) tiles = (long long)c->num_sms * 8;
        translate_fir_delay_tiled_kernel<24><<<(unsigned)tiles, 256, 0, c->stream>>>(y, pairs, cv->fp, cv->hist[cur][0].as<short>(),
                                                                                    cv->hist[cur][1].as<short>(), dev);
    } else {
        translate_fir_delay_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(y, pairs, cv->fp, cv->hist[cur][0].as<short>(),
                                                                            cv->hist[cur][1].as<short>(), dev);
    }
    history_kernel<<<1, kMaxTaps, 0, c->stream>>>(y, pairs, cv->hist[cur][0].as<short>(), cv->hist[cur][1].as<short>(),
                                                  cv->hist[cur ^ 1][0].as<short>(), cv->hist[cur ^ 1][1].as<short>());
    RFA_CK(cudaGetLastError());
    cv->cur ^= 1;
    cv->chunks_total += nchunks;
    c->launches += 5;
    if (mem == RFA_MEM_HOST) {
        RFA_CK(cudaMemcpyAsync(samples, dev, (size_t)len * sizeof(short), cudaMemcpyDeviceToHost, c->stream));
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    return RFA_OK;
}

int rfa_iqconverter_stats(rfa_iqconverter *cv, long long *chunks, long long *rerun, long long *dead_zone) {
    RFA_REQUIRE(cv != nullptr, "rfa_iqconverter_stats: NULL");
    rfa_ctx *c = cv->ctx;
    if (int rc = c->use()) return rc;
    unsigned int h[4] = {0, 0, 0, 0};
    RFA_CK(cudaMemcpyAsync(h, cv->counters.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    RFA_CK(cudaStreamSynchronize(c->stream));
    if (chunks) *chunks = cv->chunks_total;
    if (rerun) *rerun = h[1];
    if (dead_zone) *dead_zone = h[2];
    return RFA_OK;
}

/* airspy.c:299-309 convert_samples_int16: raw 12-bit ADC words -> (raw - 2048) << 4 */
__global__ void airspy_convert_kernel(const unsigned short *__restrict__ src, short *__restrict__ dst, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = (short)(((int)src[i] - 2048) << 4);
}

int rfa_airspy_convert_samples(rfa_ctx *c, const uint16_t *src, int16_t *dst, long long count, int mem) {
    RFA_REQUIRE(c && count >= 0, "rfa_airspy_convert_samples: bad argument");
    if (count == 0) return RFA_OK;
    RFA_REQUIRE(src && dst, "NULL buffer");
    if (int rc = c->use()) return rc;
    const unsigned short *ds = src;
    short *dd = dst;
    if (mem == RFA_MEM_HOST) {
        if (int rc = c->stage[0].ensure((size_t)count * 2)) return rc;
        if (int rc = c->stage[1].ensure((size_t)count * 2)) return rc;
        RFA_CK(cudaMemcpyAsync(c->stage[0].p, src, (size_t)count * 2, cudaMemcpyHostToDevice, c->stream));
        ds = c->stage[0].as<unsigned short>();
        dd = c->stage[1].as<short>();
    }
    long long blocks = (count + 255) / 256;
    if (blocks > (long long)c->num_sms * 8) blocks = (long long)c->num_sms * 8;
    airspy_convert_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(ds, dd, count);
    RFA_CK(cudaGetLastError());
    c->launches++;
    if (mem == RFA_MEM_HOST) {
        RFA_CK(cudaMemcpyAsync(dst, dd, (size_t)count * 2, cudaMemcpyDeviceToHost, c->stream));
        RFA_CK(cudaStreamSynchronize(c->stream));
    }
    return RFA_OK;
}

}  // extern "C"
